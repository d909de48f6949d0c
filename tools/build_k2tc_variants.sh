#!/bin/bash
# Builds experiment variants of tools/k2tc_bench: each argument is name:"nvcc -D flags".
set -e
cd "$(dirname "$0")"
B="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -I../fullycnnspeechenhancement_b200/csrc"
mkdir -p k2v
pids=()
for v in "$@"; do
  name="${v%%:*}"; flags="${v#*:}"
  [ "$flags" = "$v" ] && flags=""
  ( $B $flags -o k2v/$name k2tc_bench.cu > k2v/$name.log 2>&1 || { echo "FAILED $name"; tail -5 k2v/$name.log; } ) &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
ls k2v | grep -v log
