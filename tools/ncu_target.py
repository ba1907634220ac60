"""Target for ncu captures: windows of the bench workload (simulated pile), resident in HBM, `launches` device-timed launches.
   python tools/ncu_target.py [Mb] [coverage] [launches] [k]      (ncu: -k regex:dcu_window -s 1 -c 1 skips the warm-up launch)"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import daccord_b200 as d
from daccord_b200.host import Dataset
mb = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
cov = float(sys.argv[2]) if len(sys.argv) > 2 else 40.0
nl = int(sys.argv[3]) if len(sys.argv) > 3 else 2
k = int(sys.argv[4]) if len(sys.argv) > 4 else 8
ds = Dataset.simulate(int(mb * 1e6 / cov), read_len=10000, coverage=cov, seed=0)
batch = ds.pile(0, ds.nreads, w=40, a=10, nthreads=os.cpu_count() or 1)
pi, pd, cor = ds.profile()
e = d.Engine(d.Params.default(w=40, k_lo=k, k_hi=k, p_i=pi, p_d=pd, est_cor=cor), 0)
e.set_reads(np.array(ds.packed(), copy=True))
e.upload(batch.win, batch.sl)
for it in range(nl):
    ms = e.launch()
    print("coverage %g windows %d kernel %.2f ms  %.3f Mwin/s stats %s" % (cov, len(batch.win), ms, len(batch.win) / ms / 1e3, e.stats()), flush=True)
