#!/bin/bash
# A/B builds of the streaming bidiagonalization kernels for same-box comparisons (boxes differ by several percent):
#   bench/ab_build.sh            -> build/ab/libsvdgpu_<tag>.so for every variant below
#   SVD_GPU_LIB=build/ab/libsvdgpu_<tag>.so python bench/exp_knobs.py --n 16384 --values-only
# Only bidiag.cu depends on the switches; the other objects come from the regular build (run `make` first).
set -e
cd "$(dirname "$0")/.."
mkdir -p build/ab
variants=(${AB_VARIANTS:-"hw2_b1_e0:-DSVDGPU_FZ_HW=2 -DSVDGPU_FZ_BATCH=1 -DSVDGPU_FZ_EARLY=0"
          "hw4_b4_e0:-DSVDGPU_FZ_HW=4 -DSVDGPU_FZ_BATCH=4 -DSVDGPU_FZ_EARLY=0"
          "hw3_b4_e0:-DSVDGPU_FZ_HW=3 -DSVDGPU_FZ_BATCH=4 -DSVDGPU_FZ_EARLY=0"
          "hw5_b4_e0:-DSVDGPU_FZ_HW=5 -DSVDGPU_FZ_BATCH=4 -DSVDGPU_FZ_EARLY=0"})
others=$(ls build/obj/*.o | grep -v "/bidiag.o")
for v in "${variants[@]}"; do
  tag=${v%%:*}; flags=${v#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $flags -c ddc_svd_b200/csrc/bidiag.cu -o build/ab/bidiag_$tag.o &
done
wait
for v in "${variants[@]}"; do
  tag=${v%%:*}
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/ab/libsvdgpu_$tag.so build/ab/bidiag_$tag.o $others -lcudart -ldl -lpthread
  echo built build/ab/libsvdgpu_$tag.so
done
