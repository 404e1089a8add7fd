#!/bin/bash
# A/B two builds of libb200exec.so on the same box: interleaved runs of the q1 probe
export MSF=${MSF:-10000} REPS=${REPS:-4} Q=${Q:-q1}
for round in 1 2; do
  for lib in "" "$1"; do
    for cfg in ${CFGS:-"4:256" "2:384"}; do
      r=${cfg%%:*}; b=${cfg##*:}
      printf "lib=%s R=%s B=%s " "${lib:-default}" "$r" "$b"
      B200EXEC_LIB=$lib B200_FUSED_R=$r B200_FUSED_B=$b python tools/kernel_probe.py | tail -2 | awk '{printf "%s ", $4}'; echo
    done
  done
done
