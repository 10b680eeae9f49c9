#!/bin/bash
# usage: scripts/build_variant.sh out.so [-Dflags...]   -- builds an A/B variant of libbathgpu.so and prints regs/spills
out=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -ftz=true -Xptxas -v -shared -Xcompiler -fPIC "$@" -o $out bath_b200/csrc/bathgpu.cu 2>&1 | grep -E "error|Compiling|Used|spill" | paste - - - | sed -E 's/.*kernelILi([0-9]+)ELb([01]).*bytes stack frame, ([0-9]+) bytes spill stores, ([0-9]+) bytes spill loads.*Used ([0-9]+) registers.*/J=\1 xmx=\2 spillst=\3 spillld=\4 regs=\5/' | grep -E "xmx=0" | grep -E "J=(4|5|6|7|8|10|12) " | tr '\n' ';'; echo
