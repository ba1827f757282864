#!/bin/bash
# round-end measurement set (run under gpurun from the repo root); outputs under gpurun_out/
set -x
timeout 800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_r01_tc.json 2> gpurun_out/bench_r01_tc.err
python bench.py --workload cfg4 --steps 5 --warmup 3 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
python bench.py --workload cfg2 --steps 20 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
python bench.py --workload cfg1 --steps 20 --warmup 5 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_tc.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_cfg4.csv python bench.py --workload cfg4 --steps 2 --warmup 3 > gpurun_out/ncu_launches_cfg4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cc_tc -s 6 -c 2 -o gpurun_out/prof_r01_tc_bench -f python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cc_tcw -s 3 -c 1 -o gpurun_out/prof_r01_tcw_bench -f python bench.py --workload cfg4 --steps 2 --warmup 3 > gpurun_out/ncu_full_cfg4.log 2>&1
for f in bench_r01_tc bench_cfg4 bench_cfg2 bench_cfg1 bench_reference; do python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
    print("$f", d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("launch_ms"), (d.get("roofline") or {}).get("frac"), (d.get("clocks") or {}).get("reasons"))
except Exception as e:
    print("$f FAILED", e)
PY
done
