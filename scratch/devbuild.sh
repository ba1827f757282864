#!/bin/bash
# development build: only the TN=13 kernels
cd /root/repo/chain_control_b200/csrc && nvcc -DCC_DEV_BUILD_TN13 -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xptxas -v -shared -Xcompiler -fPIC -o libccb200.so cc_abi.cu 2>&1 | grep -E "error|Compiling entry.*rollout|registers|spill" | grep -A2 "entry.*rollout" | grep -E "entry|Used|spill" | sed -e 's/ptxas info    : //' | paste - - - | sed -e "s/Compiling entry function '_Z21cc_rollout_//; s/Ev9CcKParams' for 'sm_100a'//"
