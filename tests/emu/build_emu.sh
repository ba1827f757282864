#!/bin/bash
# TEST INFRASTRUCTURE: compiles the device code of chain_control_b200/csrc with g++ against the CPU
# fiber emulation in cuda_emu.h.  Output: tests/emu/libccb200_emu.so (same C ABI, host pointers).
set -e
cd "$(dirname "$0")"
SAN=${CC_EMU_SAN:-}
/usr/bin/g++ -std=c++17 -O1 -g -fPIC -shared -DCC_EMULATE $SAN -I. -I../../chain_control_b200/csrc \
  -x c++ ../../chain_control_b200/csrc/cc_abi.cu -o libccb200_emu.so
