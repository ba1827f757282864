#!/bin/bash
# development build: only the TN given (default 13) is rebuilt; other TN objects are reused if present
TN=${1:-13}
cd /root/repo/chain_control_b200/csrc
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
nvcc $F -c cc_abi.cu -o cc_abi.o &
nvcc $F -c cc_tc.cu -o cc_tc.o &
nvcc $F -Xptxas -v -DCC_TN_VALUE=$TN -c cc_tn.cu -o cc_tn$TN.o 2>&1 | grep -E "error|Compiling entry.*rollout|registers|spill" | grep -A2 "entry.*rollout" | grep -E "entry|Used|spill" | sed -e 's/ptxas info    : //' | paste - - - | sed -e "s/Compiling entry function '_Z21cc_rollout_//; s/Ev9CcKParams' for 'sm_100a'//" | cut -c1-140
wait
for t in 8 13 16; do [ -f cc_tn$t.o ] || nvcc $F -DCC_TN_VALUE=$t -c cc_tn.cu -o cc_tn$t.o & done; wait
nvcc -shared -o libccb200.so cc_abi.o cc_tc.o cc_tn8.o cc_tn13.o cc_tn16.o && ls -la libccb200.so
