"""small mixed batch through every kernel (window kernel both tiers, piling kernels); run under compute-sanitizer"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import daccord_b200 as d
from daccord_b200.host import Dataset
from common import synth_batch
for depth, n, rf, kw in ((40, 96, 0.2, {}), (8, 96, 0.5, {}), (6, 64, 0.3, dict(min_ff=0, max_ff=0))):
    packed, win, sl, _ = synth_batch(n, depth, seed=depth, repeat_frac=rf, depth_jitter=2)
    e = d.Engine(d.Params.default(**kw), 0); e.set_reads(packed); out = e.run(win, sl); print(depth, np.bincount(out[0]["status"], minlength=3), e.stats()); e.close()
ds = Dataset.simulate(6000, read_len=1500, coverage=12, seed=3)
pi, pd, cor = ds.profile()
e = d.Engine(d.Params.default(p_i=pi, p_d=pd, est_cor=cor), 0); e.set_reads(np.array(ds.packed(), copy=True))
ovl, trace, boff, rlen = ds.overlaps()
print("pile", e.pile(ovl, trace, ds.tspace, boff, rlen)); e.launch(); out = e.download(); print(np.bincount(out[0]["status"], minlength=3)); e.close()
