import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
import test_gpu_tmjp as T
from raoteh_b200 import tmjp

def dump(ch, t):
    ns, e = ch.primary_trajectory(t)
    print('primary nodes', ns)
    for c, (tt, ss) in sorted(e.items()):
        print('  edge', c, 'len', ch.sched.length[c], 'times', np.round(tt, 7).tolist(), 'states', ss.tolist())
    for cls in range(3):
        bits, te = ch.tolerance_trajectory(t, cls)
        print(' tol', cls, 'nodes', bits.tolist())
        for c, (tt, ss) in sorted(te.items()):
            print('   edge', c, 'times', np.array(tt, dtype=np.float64).tolist(), 'states', ss.tolist())

for seed in range(4000, 4016):
    ch = T._toy_chains(512, seed, T.DISEASE)[0]
    ch.initialize()
    bad = None
    for sw in range(180):
        ch._run(tmjp.MODE_SWEEP, n_sweeps=1, flags=0)
        ch.sweeps_done += 1
        st = ch.status.cpu().numpy()
        if (st != 0).any():
            bad = (sw, int(np.nonzero(st)[0][0]), int(st[np.nonzero(st)[0][0]]))
            break
    print('seed', seed, 'bad', bad)
    if bad:
        sw, t, code = bad
        print('state after failing sweep (primary new if tol failed):')
        dump(ch, t)
        ch2 = T._toy_chains(512, seed, T.DISEASE)[0]
        ch2.initialize()
        for k in range(sw):
            ch2._run(tmjp.MODE_SWEEP, n_sweeps=1, flags=0)
            ch2.sweeps_done += 1
        print('state before failing sweep:')
        dump(ch2, t)
        break
