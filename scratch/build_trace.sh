#!/bin/bash
# builds scratch/libtrace.so = the product sources with -DD3B_TRACE (development aid, not shipped)
set -e
cd "$(dirname "$0")/.."
mkdir -p scratch/obj
for f in det3d_b200/csrc/*.cu; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Iinclude -DD3B_TRACE -c $f -o scratch/obj/$(basename $f .cu).o &
done
wait
nvcc -shared -Wno-deprecated-gpu-targets -o scratch/libtrace.so scratch/obj/*.o
echo built scratch/libtrace.so
