#!/bin/bash
# usage: tools/sweep_phase_delay.sh 0 1000 2000 ...   (cycles; one-time start offset of the second CTA per SM)
for d in "$@"; do
  RT_PRUNE_DMMA_PHASE_DELAY=$d python tools/time_c3_up.py 2>&1 | tail -1
done
