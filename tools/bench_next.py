import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench_legs
print(json.dumps(bench_legs.bench_next_rows(torch.device('cuda:0'), None)))
