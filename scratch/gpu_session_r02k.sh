#!/bin/bash
# final bench lines after the prep-reuse change (tests ran in the previous call)
mkdir -p gpurun_out
( time python bench.py > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err ) 2>&1 | grep real
tail -c 300 gpurun_out/r02g_bench_n1.json; tail -3 gpurun_out/r02g_bench_n1.err
python bench.py --impl reference > gpurun_out/r02g_bench_ref_n1.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/r02g_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cc_lin_cl -s 2 -c 2 -o gpurun_out/r02k_lin_cl -f python scratch/prof_cl.py > gpurun_out/r02k_ncu_lin_cl.log 2>&1
ncu -i gpurun_out/r02k_lin_cl.ncu-rep --page raw --csv > gpurun_out/r02k_lin_cl_raw.csv 2>/dev/null
PB=65536 PT=1001 ncu --set full --clock-control none -k regex:cc_lin_prep -s 1 -c 1 -o gpurun_out/r02k_prep -f python scratch/prof_cl.py > gpurun_out/r02k_ncu_prep.log 2>&1
ncu -i gpurun_out/r02k_prep.ncu-rep --page raw --csv > gpurun_out/r02k_prep_raw.csv 2>/dev/null
ls -la gpurun_out | grep r02k
