#!/bin/bash
# Instrumented build of the library (conv_slab.cu with -DDFF_SLAB_TRACE): tools/libdff_trace.so.  Use with
#   DFF_B200_LIB=tools/libdff_trace.so DFF_SLAB_TRACE=1 python tools/trace_conv.py ...
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()"
B=dffinthewild_b200/csrc/build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DDFF_SLAB_TRACE -c dffinthewild_b200/csrc/conv_slab.cu -o /tmp/conv_slab_trace.o
OBJS=$(ls $B/*.o | grep -v conv_slab.o)
nvcc -shared -o tools/libdff_trace.so $OBJS /tmp/conv_slab_trace.o -lcudart_static -ldl -lrt -lpthread
echo built tools/libdff_trace.so
