#!/bin/bash
# Developer helper: scaling of the fused kernels with resident warps per SM.
for nw in 4 8 12 16; do
  PIVB200_NWARPS=$nw python tools/_sweep.py 16 2>&1 | tail -1
done
