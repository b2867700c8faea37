"""C4 Rao-Teh sweeps at full size (128 chains x 10000 sites), GPU part of the bench leg only."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench_legs
class A: no_cpu = True
r = bench_legs.bench_c4(torch.device('cuda:0'), A(), launches=3)
print(json.dumps({k: r[k] for k in r if k in ('ms_per_launch', 'sweeps_per_sec', 'mean_real_jumps_per_trajectory')}))
