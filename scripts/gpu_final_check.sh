#!/bin/bash
# End-of-round verification on one B200 (run under gpurun): the whole GPU test suite with the shipped defaults, the
# per-engine parity report, smoke(), the default bench line and an ncu launch list of one eager pair.
mkdir -p gpurun_out
S=gpurun_out/final_summary.txt
: > $S
t0=$(date +%s)
timeout 480 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$? t=$(( $(date +%s) - t0 ))s: $(tail -1 gpurun_out/pytest_gpu.log)" >> $S
timeout 90 python scripts/engine_parity_report.py > gpurun_out/engine_parity.log 2>&1; echo "engine_parity rc=$? t=$(( $(date +%s) - t0 ))s" >> $S
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? t=$(( $(date +%s) - t0 ))s" >> $S
timeout 240 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench_default rc=$? t=$(( $(date +%s) - t0 ))s" >> $S
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:rf:: -c 1200 --csv \
    --log-file gpurun_out/launches_f16.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --lanes 1 > gpurun_out/ncu_launches.log 2>&1
echo "ncu_launches rc=$? t=$(( $(date +%s) - t0 ))s" >> $S
cat $S
