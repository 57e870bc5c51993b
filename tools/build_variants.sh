#!/bin/bash
# Developer helper: builds liboptistate_kf.so variants of the decoupled-group FP64 kernels (minimum resident blocks per SM,
# summary sums in shared memory or registers) into variants/<name>/ so that one gpurun call can time them all:
#     LD_LIBRARY_PATH=variants/<name> python tools/rate.py f64:summary
# (the extension finds the library through RUNPATH=$ORIGIN, which LD_LIBRARY_PATH overrides)
set -e
cd "$(dirname "$0")/.."
CS=optistate_b200/csrc
OBJ=optistate_b200/_lib/obj
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
for cfg in "$@"; do
  minb=${cfg%%:*}; acc=${cfg##*:}
  name=b${minb}_a${acc}
  mkdir -p variants/$name /tmp/var_$name
  for tu in kf_seq_tma_f64_sum_blk kf_seq_tma_f64_nosum_blk; do
    nvcc $FLAGS -DOKF_BLK_MINB=$minb -DOKF_BLK_ACC_SMEM=$acc -c -o /tmp/var_$name/$tu.o $CS/$tu.cu &
  done
  wait
  others=$(ls $OBJ/*.o | grep -v "kf_seq_tma_f64_sum_blk.o\|kf_seq_tma_f64_nosum_blk.o")
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/$name/liboptistate_kf.so $others /tmp/var_$name/*.o
  echo built variants/$name
done
