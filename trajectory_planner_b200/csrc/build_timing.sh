#!/bin/bash
# development build with per-phase cycle counters in the fused L-BFGS kernel (TP_LBFGS_TIMING), loaded through TP_B200_LIB
set -e
cd "$(dirname "$0")"
mkdir -p build
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --fmad=false -Xcompiler -fPIC,-ffp-contract=off -DTP_LBFGS_TIMING=${TIMING_LEVEL:-2} $EXTRA -c tp_vigo.cu -o build/tp_vigo_timing.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o build/libtp_timing.so build/tp_map.o build/tp_frontend.o build/tp_vigo_timing.o -cudart static
