#!/bin/bash
# development build with per-phase cycle counters in the fused L-BFGS kernels (TP_WF_TIMING: warp form; TP_LBFGS_TIMING=n:
# round-1 block form), loaded through TP_B200_LIB=trajectory_planner_b200/csrc/build/libtp_timing.so
set -e
cd "$(dirname "$0")"
mkdir -p build
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -Xcompiler -fPIC,-ffp-contract=off ${TIMING_DEFS:--DTP_WF_TIMING} $EXTRA -c tp_vigo.cu -o build/tp_vigo_timing.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/libtp_timing.so build/tp_map.o build/tp_frontend.o build/tp_vigo_timing.o -cudart static
