#!/bin/bash
# One GPU call that re-measures the K2-TC experiment switches with the kernel harness (tools/k2tc_bench.cu):
#   bash tools/k2tc_experiments.sh build      # here (no GPU): cross-compiles the variants into tools/k2v/
#   gpurun -- 'bash tools/k2tc_experiments.sh run'   # on the B200: times them, writes gpurun_out/k2tc_experiments.txt
# Round-1 results of the same table: DESIGN.md section 4 (K2-TC, "History" and the role table).
set -e
cd "$(dirname "$0")"
VARIANTS=(
  "default"
  "issuers2:-DRCED_TC_ISSUERS=2"
  "issuers4:-DRCED_TC_ISSUERS=4"
  "spinwait:-DRCED_TC_EPIWAIT=0"
  "inflight2:-DRCED_TC_MAXINFLIGHT=2"
  "l2hint:-DRCED_TC_SKIPHINT=1"
  "l1bypass:-DRCED_TC_SKIPHINT=3"
  "validrows:-DRCED_TC_DIAG_VALIDROWS"
  "boundary0:-DRCED_TC_BOUNDARY=0"
  "trace:-DRCED_TC_TRACING=1"
  "unroll1:-DRCED_TC_UNROLL_UNITS=1"
  "unroll4:-DRCED_TC_UNROLL_UNITS=4"
  "packest1_DIAG:-DRCED_TC_DIAG_PACKEST=1"
  "packest2_DIAG:-DRCED_TC_DIAG_PACKEST=2"
  "skipbulk:-DRCED_TC_SKIP_BULK=1"
  "skipdeferred:-DRCED_TC_SAVE_DEFERRED"
  "noskip_DIAG:-DRCED_TC_DIAG_NOSKIP"
  "nosave_DIAG:-DRCED_TC_DIAG_NOSAVE"
  "noadd_DIAG:-DRCED_TC_DIAG_NOADD"
)
# RCED_K2TC_ONLY="default saveinline" restricts the run to some variants
case "$1" in
  build)
    ./build_k2tc_variants.sh "${VARIANTS[@]}"
    ;;
  run)
    mkdir -p ../gpurun_out
    out=../gpurun_out/k2tc_experiments.txt
    : > "$out"
    for v in "${VARIANTS[@]}"; do
      name="${v%%:*}"
      if [ -n "$RCED_K2TC_ONLY" ] && ! echo " $RCED_K2TC_ONLY " | grep -q " $name "; then continue; fi
      for arch in ${RCED_K2TC_ARCHS:-2 1 3}; do
        timeout 60 ./k2v/"$name" $arch "$name" | tee -a "$out"
      done
    done
    K2TC_PERSIST=1 timeout 60 ./k2v/default 2 persist | tee -a "$out"
    timeout 60 ./k2v/trace 2 trace ../gpurun_out/tc_trace.txt > /dev/null && python tc_trace_report.py ../gpurun_out/tc_trace.txt | grep -E "first_poll|total" >> "$out"
    ;;
  *)
    echo "usage: $0 build|run"; exit 2
    ;;
esac
