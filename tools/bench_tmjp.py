import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench_legs as raoteh_bench
dev = torch.device('cuda:0')
which = sys.argv[1] if len(sys.argv) > 1 else 'all'
if which in ('all', 'c5'):
    print(json.dumps(raoteh_bench.bench_c5(dev, None, n_sites=int(os.environ.get('C5_SITES', 100000)),
                                           loglik_sites=int(os.environ.get('C5_LL_SITES', 1000000)))))
if which in ('all', 'codon'):
    print(json.dumps(raoteh_bench.bench_codon_raoteh(dev, None)))
