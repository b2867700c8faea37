"""C4 sweeps/s at a reduced number of chains (same per-trajectory work as the full configuration)."""
import sys
import json
import torch
sys.path.insert(0, '.')
import bench_legs
n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 256
r = bench_legs.bench_c4_sharded(torch.device('cuda:0'), 0, 1, n_chains=n_chains, timed_sweeps=50, cpu_leg=False)
print(json.dumps(dict(value=r['value'], ms=r['ms'], jumps=r['mean_real_jumps_per_trajectory'],
                      dwell_err=r['dwell_per_sweep_per_site_minus_tree_length'])))
