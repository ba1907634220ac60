"""Workspace footprint study for the next kernel generation: runs the kernel source as host emulation (-DDCU_EMU_STATS) over windows
of the bench workload and prints percentiles of the per-window peaks of every workspace counter, plus the bytes a compact
(common-case) layout would need.  CPU only.   python tools/footprint.py [genome_len] [coverage]"""
import ctypes as C
import os
import subprocess
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import default_params, alloc_out, _ptr   # noqa: E402
from daccord_b200.host import Dataset               # noqa: E402

NAMES = ["slices", "bases", "nodes", "instances", "gapfill_extras", "unitigs", "unitig_positions", "reverse_links", "reverse_paths", "emulation_ns", "forward_paths",
         "score_intervals", "raw_unitigs", "unitig_link_symbols", "distinct_kmers"]


def main():
    glen = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
    cov = float(sys.argv[2]) if len(sys.argv) > 2 else 40.0
    out = os.path.join(ROOT, "tests", "emu", "_build", "libemu_stats.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-DDCU_EMU_STATS", "-o", out, os.path.join(ROOT, "tests", "emu", "emu.cpp")])
    ds = Dataset.simulate(glen, read_len=10000, coverage=cov, seed=1)
    pi, pd, cor = ds.profile()
    p = default_params(p_i=pi, p_d=pd, est_cor=cor)
    b = ds.pile(nthreads=8)
    win = b.win[:: max(1, len(b.win) // 20000)].copy()
    fn = "/tmp/footprint_%d.txt" % os.getpid()
    os.environ["DCU_FOOTPRINT_OUT"] = fn
    lib = C.CDLL(out)
    res, cons, ops = alloc_out(len(win))
    nov = C.c_uint64(0)
    packed = np.ascontiguousarray(ds.packed())
    sl = b.sl.copy()
    rc = lib.emu_run_batch(C.byref(p), _ptr(packed), _ptr(win), C.c_uint64(len(win)), _ptr(sl), _ptr(res), _ptr(cons), _ptr(ops), C.c_int(1), C.byref(nov))
    assert rc == 0
    a = np.loadtxt(fn, dtype=np.int64)
    os.remove(fn)
    print("windows %d (coverage %.0f), consensus %d, overflow %d" % (len(win), cov, int((res["status"] == 1).sum()), nov.value))
    print("%-22s %8s %8s %8s %8s %8s" % ("counter", "median", "p90", "p99", "p99.9", "max"))
    for i, n in enumerate(NAMES):
        c = a[:, i]
        print("%-22s %8d %8d %8d %8d %8d" % (n, np.median(c), np.percentile(c, 90), np.percentile(c, 99), np.percentile(c, 99.9), c.max()))
    # bytes of a compact layout: hash 8 B / slot at load <= 0.5, instances 2 x 1 B + 4 B slot, nodes ~40 B, unitig positions 3 doubles, paths ~32 B
    def layout(ix):
        v = {n: np.percentile(a[:, i], ix) for i, n in enumerate(NAMES)}
        hs = 2 ** int(np.ceil(np.log2(max(2 * v["distinct_kmers"], 64))))
        return (v["bases"] + 8 * hs + 6 * v["instances"] + 40 * v["nodes"] + 3 * v["unitig_link_symbols"] + 24 * v["unitigs"] + 24 * v["unitig_positions"] +
                4 * v["reverse_links"] + 36 * v["reverse_paths"] + 28 * v["forward_paths"] + 26 * v["score_intervals"] + 4096)
    # imbalance of the phase-synchronous groups: consecutive windows share a group of 16 warps; a group advances at the pace of its slowest member
    t = a[:, 9].astype(np.float64)
    g = t[: len(t) // 16 * 16].reshape(-1, 16)
    print("single-lane emulation time per window: median %.0f us, p99 %.0f us, max %.0f us; sum over groups of 16 of (16 x max) / sum = %.2f" %
          (np.median(t) / 1e3, np.percentile(t, 99) / 1e3, t.max() / 1e3, 16 * g.max(axis=1).sum() / g.sum()))
    print("compact workspace bytes: median %.0f KB, p90 %.0f KB, p99 %.0f KB, p99.9 %.0f KB" % tuple(layout(q) / 1024 for q in (50, 90, 99, 99.9)))


if __name__ == "__main__":
    main()
