set -u
OUT=gpurun_out; E=f16x3
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size,launch__registers_per_thread
timeout 400 ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file $OUT/r2_metrics_$E.csv python scripts/one_pair.py $E > $OUT/r2_ncu_metrics.log 2>&1
full() {
    timeout 300 ncu --set full --clock-control none --profile-from-start off -k regex:"$2" --launch-skip $3 --launch-count $4 -f -o /tmp/r2_$1 python scripts/one_pair.py $E > $OUT/r2_ncu_full_$1.log 2>&1
    ncu -i /tmp/r2_$1.ncu-rep --page raw --csv > $OUT/r2_full_$1_raw.csv 2>/dev/null
    ncu -i /tmp/r2_$1.ncu-rep --page details --csv > $OUT/r2_full_$1_details.csv 2>/dev/null
    rm -f /tmp/r2_$1.ncu-rep
}
full tap128_halo "tc_split_kernel" 42 2
full res2 "tc_split_kernel" 18 1
full tap64 "tc_split_kernel" 29 1
rm -f $OUT/r2_full_halo_* $OUT/r2_full_tap_* $OUT/r2_ncu_full_halo.log $OUT/r2_ncu_full_tap.log
wc -l $OUT/r2_metrics_$E.csv; du -sh $OUT
