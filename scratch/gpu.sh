#!/bin/bash
# scratch helper: rebuild the in-tree .so (abort on failure), then run a command on the GPU box
set -e
cd /root/repo
python -c "import __graft_entry__ as g; g.build()" > /tmp/build.log 2>&1 || { grep -E "error|Error" -A3 /tmp/build.log | head -30; exit 1; }
T=${GPU_TIMEOUT:-1500}
/usr/local/graft/bin/gpurun --timeout $T -- "$@"
