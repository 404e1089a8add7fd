#!/bin/bash
# ncu captures of the round-2 kernels: one `--set full` report per kernel (first launch of one big operator instance, see
# tools/op_probe.py), summarised by tools/ncu_summary.py into profiles/r02_<kernel>.txt, plus the engine's own CUDA-event
# timing of the same launches WITHOUT the profiler (gpurun_out/r02_probe_<op>.json).
set -u
mkdir -p gpurun_out
cap() {  # op kernel-regex tag
  REPS=1 ncu --set full --clock-control none --import-source on -k "regex:$2" -c 1 -f -o gpurun_out/r02_$3 python tools/op_probe.py $1 > gpurun_out/r02_ncu_$3.log 2>&1
  python tools/ncu_summary.py gpurun_out/r02_$3.ncu-rep 12 > gpurun_out/r02_$3.txt 2>&1
  head -20 gpurun_out/r02_$3.txt
}
for op in join partition groupby groupby_small filter q1; do python tools/op_probe.py $op > gpurun_out/r02_probe_$op.json 2> gpurun_out/r02_probe_$op.err; done
cap join join_build2 join_build
cap join join_probe2 join_probe
cap partition part_tile_scatter partition_scatter
cap groupby groupby_kernel groupby
cap groupby_small groupby_kernel groupby_small
cap filter fast_filter filter
cap q1 fused_kernel q1_fused
