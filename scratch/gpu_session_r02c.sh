#!/bin/bash
# round-2 final measurement pass
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02c_test_all.log 2>&1; tail -4 gpurun_out/r02c_test_all.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02c_smoke.log 2>&1; tail -5 gpurun_out/r02c_smoke.log
for S in 32 64 128; do PS=$S PB=300 PT=4 timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python scratch/prof_mlp.py > gpurun_out/r02c_memcheck_$S.log 2>&1; echo "memcheck S=$S rc=$?"; tail -2 gpurun_out/r02c_memcheck_$S.log; done
( time python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err ) 2>&1 | grep real
tail -c 300 gpurun_out/r02b_bench_n1.json
( time python bench.py --impl reference > gpurun_out/r02b_bench_ref_n1.json 2>/dev/null ) 2>&1 | grep real
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-others > gpurun_out/r02b_ncu_launches.log 2>&1
PS=64 PB=65536 PT=100 ncu --set full --clock-control none --import-source on -k regex:cc_mlp -s 1 -c 1 -o gpurun_out/r02b_mlp64 -f python scratch/prof_mlp.py > gpurun_out/r02b_ncu_mlp64.log 2>&1
PS=128 PB=18944 PT=50 ncu --set full --clock-control none --import-source on -k regex:cc_mlp128_fwd -s 1 -c 1 -o gpurun_out/r02b_mlp128 -f python scratch/prof_mlp.py > gpurun_out/r02b_ncu_mlp128.log 2>&1
ncu -i gpurun_out/r02b_mlp64.ncu-rep --page raw --csv > gpurun_out/r02b_mlp64_raw.csv 2>/dev/null
ncu -i gpurun_out/r02b_mlp128.ncu-rep --page raw --csv > gpurun_out/r02b_mlp128_raw.csv 2>/dev/null
ncu -i gpurun_out/r02b_mlp64.ncu-rep --page source --csv > gpurun_out/r02b_mlp64_src.csv 2>/dev/null
ncu -i gpurun_out/r02b_mlp128.ncu-rep --page source --csv > gpurun_out/r02b_mlp128_src.csv 2>/dev/null
