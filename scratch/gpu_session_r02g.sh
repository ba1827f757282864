#!/bin/bash
# round-2 final measurement pass (fourth session): bench lines, launch list, ncu captures of the family-6 and env kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02g_test_all.log 2>&1; tail -2 gpurun_out/r02g_test_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02g_smoke.log 2>&1; tail -3 gpurun_out/r02g_smoke.log
PTIME=1 python scratch/prof_env.py > gpurun_out/r02g_env_time.log 2>&1; cat gpurun_out/r02g_env_time.log
( time python bench.py > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err ) 2>&1 | grep real
tail -c 400 gpurun_out/r02g_bench_n1.json; tail -3 gpurun_out/r02g_bench_n1.err
( time python bench.py --impl reference > gpurun_out/r02g_bench_ref_n1.json 2>/dev/null ) 2>&1 | grep real
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/r02g_ncu_launches.log 2>&1
PS=64 PB=18944 PT=16 ncu --set full --clock-control none --import-source on -k regex:cc_mlpb_bwd -c 1 -o gpurun_out/r02g_mlpb -f python scratch/prof_mlpb.py > gpurun_out/r02g_ncu_mlpb.log 2>&1
PB=65536 PT=64 ncu --set full --clock-control none --import-source on -k regex:cc_env -c 1 -o gpurun_out/r02g_env -f python scratch/prof_env.py > gpurun_out/r02g_ncu_env.log 2>&1
for k in mlpb env; do
  ncu -i gpurun_out/r02g_$k.ncu-rep --page raw --csv > gpurun_out/r02g_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02g_$k.ncu-rep --page source --csv > gpurun_out/r02g_${k}_src.csv 2>/dev/null
done
ls -la gpurun_out | tail -12
