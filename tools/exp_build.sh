#!/bin/bash
# Dev tool: build experiment variants of the library (joint_tc.cu compiled with -DCTCVR_EXP=<n>) as
# ctc-vr_b200/build/exp/<n>/libctcvr.so; timing tools load them through CTCVR_LIB.
set -e
cd "$(dirname "$0")/.."
for n in "$@"; do
  d=ctc-vr_b200/build/exp/$n
  mkdir -p $d
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
     -DCTCVR_EXP=$n -c ctc-vr_b200/csrc/joint_tc.cu -o $d/joint_tc.o
  objs=$(ls ctc-vr_b200/build/*.o | grep -v joint_tc.o | grep -v beams_stub)
  /usr/local/cuda/bin/nvcc -shared -o $d/libctcvr.so $d/joint_tc.o $objs -lcuda
done
