"""Summarise `ncu -i X.ncu-rep --page raw --csv` exports (made on the GPU box: the reports themselves exceed gpurun's return
limit) into one small JSON under profiles/.  Usage: python scripts/ncu_csv_summary.py out.json raw1.csv [raw2.csv ...]"""
import csv
import json
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__inst_executed_pipe_tensor.sum",
        "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct"]
MULT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(out, files):
    res = []
    for f in files:
        rows = list(csv.reader(open(f)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = {"export": f, "kernel": r[hdr.index("Kernel Name")].strip()}
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    d[k] = {"value": r[i], "unit": units[i]}
            rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
            if rd and wr:
                d["dram_traffic_bytes"] = float(rd["value"].replace(",", "")) * MULT.get(rd["unit"], 1) + float(wr["value"].replace(",", "")) * MULT.get(wr["unit"], 1)
            res.append(d)
    json.dump({"command": "ncu --set full --clock-control none --profile-from-start off -k regex:<name> --launch-skip S --launch-count C "
                          "python scripts/one_pair.py f16x3; ncu -i X.ncu-rep --page raw --csv (scripts/gpu_ncu_r2.sh)", "launches": res},
              open(out, "w"), indent=1)
    for d in res:
        t = d.get("gpu__time_duration.sum", {})
        print("%-60s %10s %-4s dram %8.1f MB  tensor %5s%%  dram-thr %5s%%  L2-thr %5s%%" % (
            d["kernel"][:60], t.get("value"), t.get("unit"), d.get("dram_traffic_bytes", 0) / 1e6,
            d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", {}).get("value"),
            d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", {}).get("value"),
            d.get("lts__throughput.avg.pct_of_peak_sustained_elapsed", {}).get("value")))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
