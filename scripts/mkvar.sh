#!/bin/bash
# kernel-tuning experiments: scripts/mkvar.sh NAME [nvcc flags...] -> scripts/variant_NAME.so (same sources, extra -D flags for brb_kernels.cu)
set -e
name=$1; shift
cd "$(dirname "$0")/../balance_robot_b200/csrc"
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -ftz=true -Xcompiler -fPIC"
mkdir -p /tmp/brbobj
for f in brb_cabi brb_policy brb_policy_tc; do
  [ /tmp/brbobj/$f.o -nt $f.cu ] || nvcc $F -c -o /tmp/brbobj/$f.o $f.cu 2>/dev/null &
done
nvcc $F "$@" -Xptxas -v -c -o /tmp/brbobj/k_$name.o brb_kernels.cu 2>&1 | grep -A2 "brb_step_kernelILi1" | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | head -3
wait
nvcc -shared -o ../../scripts/variant_$name.so /tmp/brbobj/k_$name.o /tmp/brbobj/brb_cabi.o /tmp/brbobj/brb_policy.o /tmp/brbobj/brb_policy_tc.o
