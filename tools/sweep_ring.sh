#!/bin/bash
for r in "$@"; do
  sed -i "s/#define RT_DENSE_RING [0-9]*/#define RT_DENSE_RING $r/" raoteh_b200/csrc/rt_prune_small.cu
  python -m raoteh_b200._build > /dev/null 2>&1 || { echo "build failed $r"; continue; }
  python - <<PY
import sys, os, json
sys.path.insert(0, os.getcwd())
import torch, bench
from raoteh_b200 import engine
from raoteh_b200.lowering import TreeSchedule
import numpy as np
class A: steps=10
dev=torch.device('cuda:0')
cfg=bench.c2_workload(0)
sched=TreeSchedule(cfg['parent'],cfg['length'])
obs=engine.Observations.from_leaf_codes(sched,cfg['codes'],cfg['leaves'],device=dev)
peaks,_=bench.measured_peaks()
r=bench.bench_c2_loglik(dev,cfg,sched,obs,A,peaks)
print('ring=$r dense ms %.4f frac %.3f  codes ms %.4f' % (r['dense_emissions']['ms'], r['dense_emissions']['hbm_frac'], r['codes']['ms']))
PY
done
