#!/bin/bash
# sweep the fused kernel's launch geometry for q1 (diagnostic; run on the GPU box)
export MSF=${MSF:-10000} REPS=${REPS:-4} Q=${Q:-q1}
for cfg in "2 384" "2 352" "4 256" "4 224" "4 192" "2 384" "4 256"; do
  set -- $cfg
  printf "R=%s B=%s: " $1 $2
  B200_FUSED_R=$1 B200_FUSED_B=$2 python tools/kernel_probe.py 2>&1 | tail -2 | awk '{printf "%s ", $4}'; echo
done
