set -u
TRN="python bench.py --mode train --model single --steps 1 --warmup 3 --no-cpu-baseline"
$TRN > gpurun_out/plain_gram.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:patch_gram_kernel -s 3 -c 1 -f -o /tmp/src_gram $TRN > gpurun_out/ncu_src_gram.log 2>&1
ncu -i /tmp/src_gram.ncu-rep --page source --csv --print-source sass > gpurun_out/src_gram2.csv 2>/dev/null
ncu -i /tmp/src_gram.ncu-rep --page raw --csv > gpurun_out/src_gram2_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:first_wgrad_codes_kernel -s 3 -c 1 -f -o /tmp/src_fw2 $TRN > gpurun_out/ncu_src_fw2.log 2>&1
ncu -i /tmp/src_fw2.ncu-rep --page source --csv --print-source sass > gpurun_out/src_fw2.csv 2>/dev/null
ncu -i /tmp/src_fw2.ncu-rep --page raw --csv > gpurun_out/src_fw2_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
