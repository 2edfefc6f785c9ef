#!/bin/bash
# A/B builds of the streaming bidiagonalization kernels for SAME-BOX comparisons (the boxes of the pool differ by
# several percent, consecutive runs on one box by < 0.01 %):
#   make && bench/ab_build.sh [tag:"-DFLAG=.. -DFLAG=.." ...]     -> build/ab/libsvdgpu_<tag>.so per variant
#   SVD_GPU_LIB=build/ab/libsvdgpu_<tag>.so python bench/exp_knobs.py --n 16384 --values-only --reps 2
# Only bidiag.cu depends on the switches (bidiag_fused.cuh: SVDGPU_FZ_HW helper warps, SVDGPU_FZ_BATCH shared-memory
# loads per batch in the sweeps, SVDGPU_FZ_EARLY early panel-row requests, SVDGPU_FZ_WP split lane reduction); the other
# objects come from the regular build.  Results of the round: profiles/r02_ab_helper_warps.log.
set -e
cd "$(dirname "$0")/.."
mkdir -p build/ab
if [ $# -gt 0 ]; then variants=("$@"); else
variants=("hw2_b1_e0_wp0:-DSVDGPU_FZ_HW=2 -DSVDGPU_FZ_BATCH=1 -DSVDGPU_FZ_EARLY=0 -DSVDGPU_FZ_WP=0"
          "hw3_b4_e0_wp0:-DSVDGPU_FZ_HW=3 -DSVDGPU_FZ_BATCH=4 -DSVDGPU_FZ_EARLY=0 -DSVDGPU_FZ_WP=0"
          "hw3_b4_e0_wp1:-DSVDGPU_FZ_HW=3 -DSVDGPU_FZ_BATCH=4 -DSVDGPU_FZ_EARLY=0 -DSVDGPU_FZ_WP=1"
          "hw4_b4_e0_wp1:-DSVDGPU_FZ_HW=4 -DSVDGPU_FZ_BATCH=4 -DSVDGPU_FZ_EARLY=0 -DSVDGPU_FZ_WP=1")
fi
others=$(ls build/obj/*.o | grep -v "/bidiag.o")
for v in "${variants[@]}"; do
  tag=${v%%:*}; flags=${v#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC $flags \
       -c ddc_svd_b200/csrc/bidiag.cu -o build/ab/bidiag_$tag.o &
done
wait
for v in "${variants[@]}"; do
  tag=${v%%:*}
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/ab/libsvdgpu_$tag.so build/ab/bidiag_$tag.o $others -lcudart -ldl -lpthread
  rm -f build/ab/bidiag_$tag.o
  echo built build/ab/libsvdgpu_$tag.so
done
