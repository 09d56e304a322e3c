#!/bin/sh
# builds tools/streamprobe/streamprobe for sm_100a (cross-compiles without a GPU)
cd "$(dirname "$0")" && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o streamprobe streamprobe.cu
