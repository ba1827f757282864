#!/bin/bash
nvidia-smi topo -m 2>&1 | head -20
for d in /sys/bus/pci/devices/*; do c=$(cat $d/class 2>/dev/null); if [ "$c" = "0x030200" ] || [ "$c" = "0x030000" ]; then echo "$d numa=$(cat $d/numa_node)"; fi; done
ls /sys/devices/system/node/ | head; for n in /sys/devices/system/node/node*; do echo "$n: $(cat $n/cpulist)"; done
python -c "import os; print('affinity', sorted(os.sched_getaffinity(0)))"
for nb in 1 0; do CC_B200_NUMA_BIND=$nb python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 2952$nb bench.py --gpus 4 --steps 20 --warmup 3 --no-others --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bind=$nb', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e']['h2d_gbs_this_box'], d['e2e'].get('host_numa_binding'))"; done
