import sys, time, os
sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "..", "tests")); sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), ".."))
import numpy as np
from common import default_params, run_oracle
from daccord_b200.host import Dataset
ds = Dataset.simulate(250000, read_len=10000, coverage=40, seed=0)
b = ds.pile(nthreads=64)
pi, pd, cor = ds.profile(); p = default_params(p_i=pi, p_d=pd, est_cor=cor)
packed = np.array(ds.packed(), copy=True)
print("windows", len(b.win), "lscpu:", os.popen("lscpu | grep -E 'Model name|Thread|Core|Socket' | tr -s ' ' | tr '\n' ';'").read())
for t in (1, 8, 16, 32, 64, 128):
    n = min(len(b.win), 3000 * t)
    r, _, _, dt = run_oracle(p, packed, b.win[:n].copy(), b.sl, t)
    print("threads %3d: %8.0f win/s  (%.0f per thread)" % (t, n / dt, n / dt / t), flush=True)
