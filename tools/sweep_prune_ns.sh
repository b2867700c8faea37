#!/bin/bash
# usage: tools/sweep_prune_ns.sh 1 2 4     (sites per thread of the S <= 4 pruning kernel)
for ns in "$@"; do
  sed -i "s/constexpr int NS = (S <= 4) ? [0-9]* : 1;/constexpr int NS = (S <= 4) ? $ns : 1;/" raoteh_b200/csrc/rt_prune_small.cu
  python -m raoteh_b200._build > /dev/null 2>&1 || { echo "build failed $ns"; continue; }
  python bench.py --no-cpu --steps 10 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); x=d['extra']['c2_loglik_only']
print('NS=$ns step %.3f up(store) %.3f codes %.4f dense %.4f (%.3f of HBM)' % (d['ms_per_step'], d['roofline']['also']['ms'], x['codes']['ms'], x['dense_emissions']['ms'], x['dense_emissions']['hbm_frac']))"
done
sed -i "s/constexpr int NS = (S <= 4) ? [0-9]* : 1;/constexpr int NS = (S <= 4) ? 2 : 1;/" raoteh_b200/csrc/rt_prune_small.cu
