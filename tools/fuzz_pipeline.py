"""CPU only: differential fuzzing of the read-level path.  Oracle driver (its own LAS / DB readers, trace reconstruction, window loop,
vote) against the product's host + GPU-stage sources run as host emulation: GPU piling (pile_core.cuh), window kernel
(window_core.cuh), GPU vote (vote_core.cuh) -- over random data sets and command-line parameters.  Byte-identical FastA or a bug.
   python tools/fuzz_pipeline.py [seconds] [first_seed]"""
import ctypes as C
import os
import sys
import tempfile
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import default_params, run_emu, emu_lib, emu_vote, oracle_lib  # noqa: E402
from daccord_b200.host import Dataset, format_segments                      # noqa: E402
from daccord_b200 import WINDOW_DT, SLICE_DT                                 # noqa: E402


def emu_pile(ds, w, a, d, D):
    ovl, trace, boff, rlen = ds.overlaps(0, ds.nreads, maxinput=D)
    cap_w = int(rlen.sum() // a + 4 * ds.nreads + 16); cap_s = cap_w * (min(d, 200) + 1) if d < 10**9 else cap_w * 120
    win = np.zeros(cap_w, WINDOW_DT); sl = np.zeros(min(cap_s, 40_000_000), SLICE_DT)
    nw, ns = C.c_uint64(0), C.c_uint64(0)
    packed = np.array(ds.packed(), copy=True)
    P = lambda x: x.ctypes.data_as(C.c_void_p)
    rc = emu_lib().emu_pile(P(ovl), C.c_uint64(len(ovl)), P(trace), C.c_uint64(len(trace)), C.c_int32(ds.tspace), P(packed), P(boff), P(rlen), C.c_uint64(ds.nreads), C.c_uint32(w), C.c_uint32(a),
                            C.c_uint64(d), P(win), C.c_uint64(len(win)), P(sl), C.c_uint64(len(sl)), C.byref(nw), C.byref(ns))
    assert rc == 0, rc
    return win[:nw.value].copy(), sl[:ns.value].copy(), packed, boff, rlen


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 600
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    lib = oracle_lib(); lib.oracle_daccord_files.restype = C.c_void_p
    t0 = time.time(); nb = nbad = 0
    with tempfile.TemporaryDirectory() as tmp:
        while time.time() - t0 < budget:
            rng = np.random.default_rng(seed)
            w, a = [(40, 10), (40, 10), (40, 5), (40, 20), (32, 8), (48, 12), (56, 14), (24, 24), (16, 4), (40, 7), (40, 15), (33, 10), (59, 13)][int(rng.integers(0, 13))]
            glen = int(rng.integers(1500, 9000)); rl = int(rng.integers(400, 2600)); cov = float(rng.integers(4, 28))
            e = float(rng.uniform(0.04, 0.2))
            ds = Dataset.simulate(glen, read_len=rl, coverage=cov, seed=seed, repeat_frac=float(rng.choice([0.0, 0.0, 0.3])), min_ovl=int(rng.integers(150, 600)),
                                  p_ins=e * 0.6, p_del=e * 0.3, p_sub=e * 0.1, tspace=int(rng.choice([100, 100, 64, 126])))
            if ds.novl == 0:
                seed += 1; continue
            d = int(rng.choice([2**64 - 1, 2**64 - 1, 3, 8, 20])); D = int(rng.choice([5000, 5000, 4, 12])); full = int(rng.random() < 0.3); minlen = int(rng.choice([0, 0, rl // 2]))
            k = int(rng.choice([8, 8, 6, 10])); k = min(k, w // 2)
            las, db = os.path.join(tmp, "f.las"), os.path.join(tmp, "f.db")
            ds.write(las, db)
            pi, pd, cor = ds.profile()
            p = default_params(w=w, k_lo=k, k_hi=k, p_i=pi, p_d=pd, est_cor=cor)
            n = C.c_uint64(0)
            ptr = lib.oracle_daccord_files(C.byref(p), C.c_uint32(a), C.c_uint64(d), C.c_uint64(D), C.c_int(full), C.c_uint64(minlen), las.encode(), db.encode(), C.c_int64(0), C.c_int64(-1), C.c_int(8), C.byref(n), None)
            want = C.string_at(ptr, n.value); lib.oracle_free(C.c_void_p(ptr))
            win, sl, packed, boff, rlen = emu_pile(ds, w, a, d, D)
            r = run_emu(p, packed, win, sl, 1)
            seg, chars = emu_vote(win, r[0], r[1], r[2], w, full, minlen, packed, boff, rlen)
            got = format_segments(seg, chars)[0]
            ok = got == want
            # error-profile estimation: oracle restatement vs the product's host implementation on the same files
            out = (C.c_uint64 * 7)(); dd = (C.c_double * 2)()
            assert lib.oracle_estimate_profile(las.encode(), db.encode(), C.c_int64(0), C.c_int64(-1), C.c_uint64(d), C.c_uint64(D), out, dd) == 0
            g = ds.estimate_profile(0, None, d, D, 4)
            ok = ok and [g[x] for x in ("matches", "mismatches", "insertions", "deletions", "usable", "unusable", "reads")] == [int(x) for x in out]
            nb += 1; nbad += (not ok)
            print("seed %d w %d a %d k %d reads %d x %d cov %.0f err %.2f tspace %d -d %s -D %d -f %d -l %d windows %d fasta %d %s" % (
                seed, w, a, k, ds.nreads, rl, cov, e, ds.tspace, "inf" if d > 10**9 else d, D, full, minlen, len(win), len(want), "OK" if ok else "MISMATCH"), flush=True)
            seed += 1
    print("data sets %d mismatches %d in %.0f s" % (nb, nbad, time.time() - t0))
    return 1 if nbad else 0


if __name__ == "__main__":
    sys.exit(main())
