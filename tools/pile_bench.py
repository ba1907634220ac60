import sys, time
sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "..", "tests")); sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), ".."))
import numpy as np
import daccord_b200 as d
from daccord_b200.host import Dataset
mb = float(sys.argv[1]) if len(sys.argv) > 1 else 10
ds = Dataset.simulate(int(mb * 1e6 / 40), read_len=10000, coverage=40, seed=0)
pi, pd, cor = ds.profile()
e = d.Engine(d.Params.default(p_i=pi, p_d=pd, est_cor=cor), 0)
e.set_reads(np.array(ds.packed(), copy=True))
t0 = time.time(); b = ds.pile(); t1 = time.time()
ovl, trace, boff, rlen = ds.overlaps(); t2 = time.time()
for it in range(3):
    t3 = time.time(); nw, ns = e.pile(ovl, trace, ds.tspace, boff, rlen); t4 = time.time()
    print("host pile %.2fs (select+export %.2fs)  gpu pile %.3fs  windows %d slices %d" % (t1 - t0, t2 - t1, t4 - t3, nw, ns), flush=True)
win, sl = e.get_windows(with_slices=True)
print("identical:", bool((win == b.win).all() and (sl == b.sl).all()))
