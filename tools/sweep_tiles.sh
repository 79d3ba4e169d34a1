#!/bin/bash
# Build tile-shape variants of the library (CPU box): bash tools/sweep_tiles.sh name:W:H:NT:PCAP:TCAP:MINB ...
cd "$(dirname "$0")/.."
mkdir -p wrach_b200/lib/sweep
for v in "$@"; do
  IFS=: read name W H NT PCAP TCAP MINB extra <<< "$v"
  (cd wrach_b200/csrc && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -ccbin /usr/bin/g++ -Xcompiler -fPIC \
     -DWRACH_TILE_W=$W -DWRACH_TILE_H=$H -DWRACH_TILE_NT=$NT -DWRACH_TILE_PCAP=$PCAP -DWRACH_TILE_TCAP=$TCAP -DWRACH_TILE_MINB=$MINB $extra \
     -Xptxas -v -shared -o ../lib/sweep/lib_$name.so wrach_worker.cu wrach_host.cpp -ldl 2>&1 | grep -A2 "Function properties for _ZN5wrach12k_tile_frameILi1" | grep -E "Used" | sed 's/ptxas info    : //' | tr '\n' '|'; echo " <- $name") &
done
wait
