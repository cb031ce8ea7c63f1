#!/bin/bash
# End-of-round evidence on one B200 (outputs small enough for gpurun's return limit).
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 500 python -m pytest tests -m gpu -q > $OUT/r2_pytest_gpu_final.log 2>&1; tail -n 3 $OUT/r2_pytest_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2_smoke.log 2>&1; tail -n 3 $OUT/r2_smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 > $OUT/r2_bench_line_N1.json 2> $OUT/r2_bench_line_N1.err
python -c "
import json;d=json.load(open('$OUT/r2_bench_line_N1.json'));print('N1', round(d['value'],1), round(d['e2e']['value'],1), d['roofline']['frac'], d['cpu_baseline']['value'], d['parity']['match_symdiff'], d['parity']['within_north_star'], d['clocks'])"
timeout 100 python scripts/profile_step.py f16x3 --seq > $OUT/r2_step_profile_f16x3.txt 2>/dev/null; head -n 4 $OUT/r2_step_profile_f16x3.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/r2_launches_f16x3.csv python scripts/one_pair.py f16x3 > $OUT/r2_ncu_launches.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size,launch__registers_per_thread
timeout 300 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file $OUT/r2_metrics_f16x3.csv python scripts/one_pair.py f16x3 > $OUT/r2_ncu_metrics.log 2>&1
full() {
    timeout 300 ncu --set full --clock-control none --profile-from-start off -k regex:"$2" --launch-skip $3 --launch-count $4 -f -o /tmp/r2_$1 python scripts/one_pair.py f16x3 > $OUT/r2_ncu_full_$1.log 2>&1
    ncu -i /tmp/r2_$1.ncu-rep --page raw --csv > $OUT/r2_full_$1_raw.csv 2>/dev/null
    ncu -i /tmp/r2_$1.ncu-rep --page details --csv > $OUT/r2_full_$1_details.csv 2>/dev/null
    rm -f /tmp/r2_$1.ncu-rep
}
full corr "tc_corr_pipe" 0 1
full stem "stem7_split" 0 1
# tc_split_kernel launches of a pair, in order: layer1 = 0..8 (5 = block 1 conv3 + residual, double-buffered staging),
# layer3 block 1 = 24 (c1), 25 (3x3 halo reuse, 128-channel tiles), 26 (conv3 + residual, wide tile)
full res2 "tc_split_kernel" 5 1
full halo128 "tc_split_kernel" 25 2
wc -l $OUT/r2_launches_f16x3.csv $OUT/r2_metrics_f16x3.csv; du -sh $OUT
