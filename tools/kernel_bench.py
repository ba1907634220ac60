import sys, time, os
sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "..", "tests")); sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), ".."))
import numpy as np
from common import default_params, synth_batch
import daccord_b200 as d
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 200
packed, win, sl, _ = synth_batch(2000, depth, seed=5)
W = np.tile(win, rep)
e = d.Engine(d.Params.default(), 0)
e.set_reads(packed)
e.upload(W, sl)
for it in range(3):
    ms = e.launch()
    print("depth %d windows %d kernel %.2f ms  %.3f Mwin/s stats %s" % (depth, len(W), ms, len(W) / ms / 1e3, e.stats()), flush=True)
t0 = time.time(); out = e.run(W, sl); t1 = time.time()
print("e2e run %.1f ms -> %.3f Mwin/s" % ((t1 - t0) * 1e3, len(W) / (t1 - t0) / 1e6))
