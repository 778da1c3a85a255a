#!/bin/bash
# Launch list of the bench command plus ncu --set full of the kernels that changed after the main capture (row driver, policy pass).
out=gpurun_out/${1:-r02k}; mkdir -p $out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $out/ncu_bench.log 2>&1
timeout 300 $NCU -k 'regex:rbergomi_rows_kernel|rows_price_kernel' -s 2 -c 2 -o $out/prof_rows python tools/rows_throughput.py 4096 > $out/ncu_rows.log 2>&1
timeout 300 $NCU -k regex:lsm_policy_kernel -c 1 -o $out/prof_policy python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-configs > $out/ncu_policy.log 2>&1
python tools/summarize_profiles.py ${1:-r02k} $out/launches.csv "row driver=$out/prof_rows.ncu-rep" "policy pass (lsm_policy_kernel, 2^26 x 253)=$out/prof_policy.ncu-rep" --out $out > $out/summarize.log 2>&1
rm -f $out/*.ncu-rep
ls -la $out | tail -8
