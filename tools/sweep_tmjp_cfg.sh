#!/bin/bash
# usage: tools/sweep_tmjp_cfg.sh "8 2" "16 1" ...   (warps per CTA, min CTAs per SM)
export C5_SITES=${C5_SITES:-60000} C5_LL_SITES=1000
for cfg in "$@"; do
  set -- $cfg
  sed -i "s/#define RT_TMJP_WARPS [0-9]*/#define RT_TMJP_WARPS $1/; s/#define RT_TMJP_MINB [0-9]*/#define RT_TMJP_MINB $2/" raoteh_b200/csrc/rt_tmjp.cu
  python -m raoteh_b200._build > /dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  echo "== warps/CTA=$1 minB=$2"
  timeout -s KILL 300 python tools/bench_tmjp.py c5 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    g=d['gibbs']; print('sweep %.3g /s  sweep+summary %.3g /s  summary %.3g traj/s' % (g['sweep']['sweeps_per_sec'], g['sweep_plus_summary']['sweeps_per_sec'], g['summary_only']['trajectories_per_sec']))
"
done
