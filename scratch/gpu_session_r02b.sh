#!/bin/bash
# round-2 (second session) measurement pass: bench (both arms), launch list, ncu full captures of the new kernels
set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err
tail -c 600 gpurun_out/r02b_bench_n1.json
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02b_bench_ref_n1.json 2>/dev/null
python scratch/sweep_cfg3_r02.py gpurun_out/r02b_sweep_cfg3.jsonl > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/r02b_ncu_launches.log 2>&1
PS=64 PB=65536 PT=100 ncu --set full --clock-control none --import-source on -k regex:cc_mlp -s 1 -c 1 -o gpurun_out/r02b_mlp64 -f python scratch/prof_mlp.py > gpurun_out/r02b_ncu_mlp64.log 2>&1
PS=128 PB=18944 PT=50 ncu --set full --clock-control none --import-source on -k regex:cc_mlp128_fwd -s 1 -c 1 -o gpurun_out/r02b_mlp128 -f python scratch/prof_mlp.py > gpurun_out/r02b_ncu_mlp128.log 2>&1
ncu -i gpurun_out/r02b_mlp64.ncu-rep --page raw --csv > gpurun_out/r02b_mlp64_raw.csv 2>/dev/null
ncu -i gpurun_out/r02b_mlp128.ncu-rep --page raw --csv > gpurun_out/r02b_mlp128_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
