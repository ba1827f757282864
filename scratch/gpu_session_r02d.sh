#!/bin/bash
# round-2, fourth session: two-segment data source + mlpb NH variants; full GPU suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_two_segments.py -m gpu -q -x > gpurun_out/r02d_env_test.log 2>&1; tail -15 gpurun_out/r02d_env_test.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02d_test_all.log 2>&1; tail -4 gpurun_out/r02d_test_all.log
