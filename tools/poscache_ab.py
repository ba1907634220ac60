"""CPU only: effect of the per-traverse cache of unitig position slots (DCU_POSCACHE, window_core.cuh stretch_positions) on windows of the
bench workload (simulated 40x / 20x pile through the host piler): warp collectives per window (32-lane emulation) and single-lane
emulation time, cache off vs on; the results of both are compared with each other as well.
   python tools/poscache_ab.py [genome_len] [coverage] [windows]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys, time, os
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
from common import default_params, run_emu, run_emu_lanes
from daccord_b200.host import Dataset
glen, cov, nw = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3])
ds = Dataset.simulate(glen, read_len=10000, coverage=cov, seed=1)
pi, pd, cor = ds.profile()
p = default_params(p_i=pi, p_d=pd, est_cor=cor)
b = ds.pile(nthreads=8)
win = b.win[:: max(1, len(b.win) // nw)].copy(); sl = b.sl.copy(); packed = np.ascontiguousarray(ds.packed())
run_emu(p, packed, win[:50], sl, 1)
t = time.time(); r = run_emu(p, packed, win, sl, 1); te = time.time() - t
rl = run_emu_lanes(p, packed, win, sl, 1, 2, 7)
assert (r[0] == rl[0]).all() and (r[1] == rl[1]).all()
np.save(sys.argv[4], np.concatenate([r[0].view(np.uint8).ravel(), r[1], r[2]]))
print("windows %%d consensus %%d: single-lane emulation %%.1f us / window, %%.0f warp collectives / window" %% (len(win), int((r[0]["status"] == 1).sum()), 1e6 * te / len(win), rl[4] / len(win)))
""" % (ROOT, os.path.join(ROOT, "tests"))


def main():
    a = sys.argv[1:] + ["60000", "40", "6000"][len(sys.argv) - 1:]
    outs = []
    for pc in ("0", "1"):
        env = dict(os.environ, DCU_POSCACHE=pc)
        out = "/tmp/poscache_ab_%s_%d.npy" % (pc, os.getpid())
        r = subprocess.run([sys.executable, "-c", CHILD] + a + [out], env=env, capture_output=True, text=True)
        print("DCU_POSCACHE=%s  %s" % (pc, r.stdout.strip() or r.stderr[-500:]))
        outs.append(out)
    import numpy as np
    x, y = np.load(outs[0]), np.load(outs[1])
    print("results identical:", bool((x == y).all()))
    for o in outs:
        os.remove(o)


if __name__ == "__main__":
    main()
