set -u
TRN="python bench.py --mode train --steps 1 --warmup 3 --no-cpu-baseline"
$TRN > gpurun_out/plain_st.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_kernel -s 36 -c 1 -f -o /tmp/src_st $TRN > gpurun_out/ncu_src_st.log 2>&1
ncu -i /tmp/src_st.ncu-rep --page source --csv --print-source sass > gpurun_out/src_fwd_stats_n128_b.csv 2>/dev/null
ncu -i /tmp/src_st.ncu-rep --page raw --csv > gpurun_out/src_fwd_stats_n128_b_raw.csv 2>/dev/null
