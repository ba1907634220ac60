"""CPU only: differential fuzzing of the kernel source (host emulation, single lane and 32 lanes under random lane schedules) against the oracle over random parameter sets (w, k range,
depth, error rates, repeats, -m, filter frequencies, -e).  Prints one line per batch; any mismatch is a parity bug.
   python tools/fuzz_parity.py [seconds] [first_seed]"""
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import default_params, synth_batch, run_oracle, run_emu, run_emu_lanes, compare_results  # noqa: E402


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 600
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    t0 = time.time(); nb = nw = nbad = 0
    while time.time() - t0 < budget:
        rng = np.random.default_rng(seed)
        w = int(rng.choice([16, 24, 32, 40, 40, 40, 48, 56, 59]))
        klo = int(rng.integers(4, min(13, w // 2)))
        khi = int(min(14, klo + rng.integers(0, 3)))
        depth = int(rng.choice([3, 5, 8, 12, 20, 30, 40, 60, 90]))
        e = float(rng.uniform(0.02, 0.28))
        p_ins, p_del, p_sub = e * 0.55, e * 0.3, e * 0.15
        maxff = int(rng.choice([2, 2, 2, 1, 3])); minff = int(rng.integers(0, maxff + 1))
        kw = dict(w=w, k_lo=klo, k_hi=khi, min_cov=int(rng.integers(2, 6)), max_ff=maxff, min_ff=minff, p_i=p_ins, p_d=p_del, est_cor=float(rng.choice([1 - e, 1 - e, 0.0])))
        if rng.random() < 0.25:
            kw["max_err"] = int(rng.integers(depth, depth * 12))
        p = default_params(**kw)
        n = int(max(20, 4000 // depth))
        packed, win, sl, _ = synth_batch(n, depth, seed=seed, w=w, p_ins=p_ins, p_del=p_del, p_sub=p_sub, repeat_frac=float(rng.choice([0.0, 0.2, 0.6])), depth_jitter=int(rng.integers(0, 3)))
        ref = run_oracle(p, packed, win, sl, 8)
        got = run_emu(p, packed, win, sl, 1)
        bad = compare_results(ref, got)
        # the same windows through the 32-lane emulation under a shuffled lane schedule (races, divergent collectives)
        sched = int(rng.integers(0, 4))
        gl = run_emu_lanes(p, packed, win, sl, 1, sched, seed)
        bad = sorted(set(bad) | set(compare_results(ref, gl)))
        nb += 1; nw += n; nbad += len(bad)
        print("seed %d w %d k %d..%d depth %d err %.2f ff %d..%d windows %d ok %d overflow %d %s" % (seed, w, klo, khi, depth, e, maxff, minff, n, int((ref[0]["status"] == 1).sum()), got[3],
                                                                                                     "OK" if not bad else "MISMATCH %s" % bad[:5]), flush=True)
        seed += 1
    print("batches %d windows %d mismatching windows %d in %.0f s" % (nb, nw, nbad, time.time() - t0))
    return 1 if nbad else 0


if __name__ == "__main__":
    sys.exit(main())
