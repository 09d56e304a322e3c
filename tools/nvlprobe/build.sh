#!/bin/sh
cd "$(dirname "$0")" && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -cudart static nvlprobe.cu -o libnvlprobe.so
