#!/bin/bash
# ncu --set full of one full-size strike-ladder launch (16 strikes x 2^22 paths), summarised on the box.
out=gpurun_out/${1:-r02d}; mkdir -p $out
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:lsm_multi_kernel -s 300 -c 1 -o $out/prof_multi python tools/one_surface.py 1 22 > $out/ncu_multi.log 2>&1
python tools/summarize_profiles.py ${1:-r02d}_multi none "strike ladder (lsm_multi_kernel, 16 strikes x 2^22 paths)=$out/prof_multi.ncu-rep" --out $out > $out/summarize.log 2>&1
python tools/ncu_by_line.py $out/prof_multi.ncu-rep montecarlooptionspricer_b200/libmcp_b200.so lsm_multi_kernelILi3 "" 50 > $out/multi_by_line.txt 2>&1
rm -f $out/*.ncu-rep
tail -3 $out/ncu_multi.log
