import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench_legs
print(bench_legs.bench_c4(torch.device('cuda:0'), None, n_chains=32, sweeps_per_launch=10, launches=2))
