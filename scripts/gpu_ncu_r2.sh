#!/bin/bash
# Round-2 ncu evidence, sized to fit gpurun's 64 MiB return limit: reports stay on the box, only CSV exports come back.
#   1. launch list of one steady-state pair (gpu__time_duration per launch, every kernel)
#   2. light metric pass over every library kernel of the pair (DRAM bytes, L2 bytes, tensor pipe, throughputs)
#   3. `--set full` of one instance of each hot kernel, exported with --page raw / --page details
set -u
OUT=gpurun_out
E=${1:-f16x3}
mkdir -p $OUT
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/r2_launches_$E.csv python scripts/one_pair.py $E > $OUT/r2_ncu_launches.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size,launch__registers_per_thread
timeout 400 ncu --metrics $M --clock-control none --profile-from-start off --csv \
    --log-file $OUT/r2_metrics_$E.csv python scripts/one_pair.py $E > $OUT/r2_ncu_metrics.log 2>&1
full() {   # name regex skip count
    timeout 300 ncu --set full --clock-control none --profile-from-start off -k regex:"$2" --launch-skip $3 --launch-count $4 \
        -f -o /tmp/r2_$1 python scripts/one_pair.py $E > $OUT/r2_ncu_full_$1.log 2>&1
    ncu -i /tmp/r2_$1.ncu-rep --page raw --csv > $OUT/r2_full_$1_raw.csv 2>/dev/null
    ncu -i /tmp/r2_$1.ncu-rep --page details --csv > $OUT/r2_full_$1_details.csv 2>/dev/null
    rm -f /tmp/r2_$1.ncu-rep
}
full corr "tc_corr_pipe" 0 1
full stem "stem7_split" 0 1
# tc_split_kernel launches of a pair, in order: 0-14 FeatureExtractor(target), 15-56 ResNet trunk, 57-71 FeatureExtractor(source), 72-79 heads
full tap128_halo "tc_split_kernel" 42 2       # trunk l3.1 c1 (1x1 1024->256, BN 128) and l3.1 c2 (3x3 halo)
full res2 "tc_split_kernel" 18 1              # trunk l1.0 c3 + residual (64->256, double-buffered staging)
full tap64 "tc_split_kernel" 29 1             # trunk l2.1 c1 (1x1 512->128)
full cneigh "corr_neigh7" 0 1
full small "maxpool_split|l2norm_split|mutual_cols|ransac_kernel|grid_sample_kernel|compose_fine" 0 8
ls -la $OUT | tail -n 20
du -sh $OUT
