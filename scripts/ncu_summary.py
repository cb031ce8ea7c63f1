"""Summarise an .ncu-rep (read here, no GPU needed) into a small JSON under profiles/.
Usage: python scripts/ncu_summary.py gpurun_out/prof_corr_r1.ncu-rep profiles/r1_corr_ncu.json"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * m.get(unit, 1)


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].strip()}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = {"value": r[i], "unit": units[i]}
        rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
        if rd and wr:
            d["dram_traffic_bytes"] = to_bytes(rd["value"], rd["unit"]) + to_bytes(wr["value"], wr["unit"])
        res.append(d)
    json.dump({"source": rep, "command": "ncu --set full --clock-control none --import-source on", "launches": res}, open(out, "w"), indent=1)
    for d in res:
        print(d["kernel"][:60], d.get("gpu__time_duration.sum"), "traffic MB", d.get("dram_traffic_bytes", 0) / 1e6)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
