#!/bin/bash
# Debug build of the library with pipeline time stamps in the tcgen05 kernel (-DMBO_TC_TRACE); used by tools/tc_trace.py
set -e
cd "$(dirname "$0")/../metabo_b200/csrc"
mkdir -p ../_build_trace
for f in api gp mlp_simt reduce neural_af_tc ppo_update ppo_update_simt comm env; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC \
    --expt-relaxed-constexpr -DMBO_TC_TRACE -DMBO_UPD_TRACE -c $f.cu -o ../_build_trace/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libmetabo_b200_trace.so ../_build_trace/*.o
echo built ../libmetabo_b200_trace.so
