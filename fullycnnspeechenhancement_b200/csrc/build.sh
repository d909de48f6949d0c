#!/bin/bash
# Builds librced_b200.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
# RCED_EXTRA_FLAGS: development switches, e.g. RCED_EXTRA_FLAGS=-DRCED_TC_TRACING=1 for the RCED_TC_TRACE clock stamps
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr $RCED_EXTRA_FLAGS"
mkdir -p build
pids=()
for f in rced_net rced_net_tc rced_stft rced_istft rced_api rced_host; do
  ( $NVCC $FLAGS -c $f.cu -o build/$f.o > build/$f.log 2>&1 || { cat build/$f.log; exit 1; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../librced_b200.so build/rced_net.o build/rced_net_tc.o build/rced_stft.o build/rced_istft.o build/rced_api.o build/rced_host.o -lcudart
echo "built $(cd .. && pwd)/librced_b200.so"
