#!/bin/sh
cd "$(dirname "$0")" && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gatherprobe gatherprobe.cu
