#!/bin/bash
# sweep the fused kernel's launch geometry (diagnostic; run on the GPU box)
export MSF=${MSF:-10000} REPS=${REPS:-3}
for cfg in "2 512" "2 384" "2 256" "4 256" "4 384" "4 128"; do
  set -- $cfg
  echo "== R=$1 B=$2"
  B200_FUSED_R=$1 B200_FUSED_B=$2 python tools/kernel_probe.py 2>&1 | grep -v "^\[b200\]" | awk '{print $1,$2,$4,$5}' | tr '\n' ';'; echo
done
