#!/bin/bash
timeout 800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
python bench.py --workload cfg4 --steps 5 --warmup 3 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
python bench.py > gpurun_out/bench_r01_tc.json 2> gpurun_out/bench_r01_tc.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_cfg4.csv python bench.py --workload cfg4 --steps 2 --warmup 3 > gpurun_out/ncu_launches_cfg4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cc_tcw -s 3 -c 1 -o gpurun_out/prof_r01_tcw_bench -f python bench.py --workload cfg4 --steps 2 --warmup 3 > gpurun_out/ncu_full_cfg4.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for f in bench_r01_tc bench_cfg4; do python - <<PY
import json
d = json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1])
print("$f", d.get("value"), d.get("ms_per_step"), d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["launch_ms"], d["roofline"]["frac"], d["roofline"].get("executed", {}).get("frac"), d["clocks"]["reasons"])
PY
done
