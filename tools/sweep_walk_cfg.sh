#!/bin/bash
# usage: tools/sweep_walk_cfg.sh "NS BLOCK" ...   (sites per thread, threads per CTA of the walk kernel)
for cfg in "$@"; do
  set -- $cfg
  sed -i "s/#define RT_WALK_NS [0-9]*/#define RT_WALK_NS $1/; s/#define RT_WALK_BLOCK [0-9]*/#define RT_WALK_BLOCK $2/" raoteh_b200/csrc/rt_posterior_small.cu
  python -m raoteh_b200._build > /dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  python bench.py --no-extra --no-cpu --steps 10 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('NS=$1 block=$2 step %.3f e2e %.3f walk %.3f up %.3f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['ms'], d['roofline']['also']['ms']))"
done
