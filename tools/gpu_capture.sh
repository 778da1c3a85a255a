#!/bin/bash
# One gpurun call: tests, bench, launch list and the ncu --set full captures that profiles/ summarises.  The reports are
# summarised ON the box (gpurun brings back at most 64 MiB) and only the generator's report travels.
#   gpurun --timeout 2400 -- 'bash tools/gpu_capture.sh <tag> [skip-tests]'
tag=${1:-r02a}; out=gpurun_out/$tag; mkdir -p $out
NCU="ncu --set full --clock-control none --import-source on"
LIB=montecarlooptionspricer_b200/libmcp_b200.so
if [ "$2" != "skip-tests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc $?" >> $out/pytest.log
  tail -30 $out/pytest.log
fi
timeout 900 python bench.py --steps 5 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "bench rc $?" >> $out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > $out/ncu_bench.log 2>&1
timeout 300 $NCU -k regex:n256pair -c 1 -o $out/prof_gen python tools/one_price.py 26 1 > $out/ncu_gen.log 2>&1
timeout 300 $NCU -k regex:lsm_sweep_tma_kernel -s 100 -c 1 -o $out/prof_sweep python tools/one_price.py 26 1 > $out/ncu_sweep.log 2>&1
timeout 300 $NCU -k regex:lsm_sweep_tma64 -s 100 -c 1 -o $out/prof_sweep64 python tools/one_lsm.py 26 f64 > $out/ncu_sweep64.log 2>&1
timeout 300 $NCU -k regex:lsm_multi_kernel -s 300 -c 1 -o $out/prof_multi python tools/one_surface.py 1 22 > $out/ncu_multi.log 2>&1
timeout 300 $NCU -k 'regex:rbergomi_rows_kernel|rows_price_kernel' -s 2 -c 2 -o $out/prof_rows python tools/rows_throughput.py 4096 > $out/ncu_rows.log 2>&1
timeout 300 $NCU -k 'regex:dual_nested_kernel|gbm_paths_kernel' -c 3 -o $out/prof_dual python tools/dual_bench.py 16 1000 > $out/ncu_dual.log 2>&1
MCP_SWEEP_IMPL=4 timeout 300 $NCU -k regex:lsm_persist -c 1 -o $out/prof_persist python tools/one_lsm.py 23 f32 > $out/ncu_persist.log 2>&1
python tools/summarize_profiles.py $tag $out/launches.csv generator=$out/prof_gen.ncu-rep "LSM sweep (fp32 carry)=$out/prof_sweep.ncu-rep" \
    "LSM sweep (fp64 carry, parity mode)=$out/prof_sweep64.ncu-rep" "strike ladder (lsm_multi_kernel)=$out/prof_multi.ncu-rep" \
    "row driver=$out/prof_rows.ncu-rep" "nested duality / GBM generator=$out/prof_dual.ncu-rep" \
    "persistent sweep, 2^23 paths=$out/prof_persist.ncu-rep" --out $out > $out/summarize.log 2>&1
python tools/ncu_by_line.py $out/prof_gen.ncu-rep $LIB n256pair_kernelILb0 "" 60 > $out/gen_by_line.txt 2>&1
python tools/ncu_by_line.py $out/prof_multi.ncu-rep $LIB lsm_multi_kernelILi3 "" 40 > $out/multi_by_line.txt 2>&1
python tools/ncu_by_line.py $out/prof_persist.ncu-rep $LIB lsm_persist_kernelILi3ELb0 "" 40 > $out/persist_by_line.txt 2>&1
ncu -i $out/prof_gen.ncu-rep --page source --csv > $out/gen_source.csv 2>/dev/null
python tools/sweep_timing.py 23 26 > $out/sweep_timing.log 2>&1
for f in $out/*.ncu-rep; do [ "$f" != "$out/prof_gen.ncu-rep" ] && rm -f $f; done
du -sh $out; ls -la $out
