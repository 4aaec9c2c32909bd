#!/bin/bash
# DDA tuning sweep: kernel time of raycast_dda under different settings
# columns: NEAR OCC SEG ELECT
while read -r a b c d; do
  export FDEM_RAY_NEAR=$a FDEM_RAY_OCC=$b FDEM_RAY_SEG=$c FDEM_RAY_ELECT=$d
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"raycast_dda" --launch-skip 2 --launch-count 2 --csv python tools/c4_steps.py c4_dense_raycast 4 2>/dev/null | grep -E "raycast_dda" | awk -F'","' -v c="$a $b $c $d" '{print c, $5, $NF}' | sed 's/(.*)//'
done <<CFG
64 6 256 0
64 8 256 0
64 6 128 0
64 8 128 0
96 0 256 0
128 0 256 0
64 8 512 0
CFG
