"""CPU only: where a window's warp collectives come from.  Runs windows through the 32-lane emulation (tests/emu/emu_lanes.cpp) with
DCU_EMU_PROFILE, which counts every collective / wsync by call site, and resolves the sites to source lines of window_core.cuh
(addr2line on the emulation library).  Collectives are a proxy for the serial depth of the lane-parallel code: ~30 issue cycles each on the GPU.
   python tools/lane_collectives.py [k] [depth] [windows] [repeat_frac]      (synthetic window batch)
   python tools/lane_collectives.py pile [coverage] [windows]               (windows of a simulated pile, as in bench.py)"""
import collections
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    pile = len(sys.argv) > 1 and sys.argv[1] == "pile"
    if pile:
        sys.argv[1] = "8"
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    depth = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    rf = float(sys.argv[4]) if len(sys.argv) > 4 else 0.1
    out = tempfile.mktemp()
    os.environ["DCU_EMU_PROFILE"] = out
    from common import default_params, synth_batch, run_emu_lanes, build_emu_lanes
    p = default_params(k_lo=k, k_hi=k)
    if pile:
        import numpy as np
        from daccord_b200.host import Dataset
        ds = Dataset.simulate(60000, read_len=10000, coverage=depth, seed=1)
        pi, pd, cor = ds.profile()
        p = default_params(p_i=pi, p_d=pd, est_cor=cor)
        b = ds.pile(nthreads=8)
        win = b.win[:: max(1, len(b.win) // n)].copy(); sl = b.sl.copy(); packed = np.ascontiguousarray(ds.packed()); n = len(win)
    else:
        packed, win, sl, _ = synth_batch(n, depth, seed=18, repeat_frac=rf, depth_jitter=3, w=p.w)
    r = run_emu_lanes(p, packed, win, sl, 1, 0, 1)
    rows = [l.split() for l in open(out)]
    os.unlink(out)
    lib = build_emu_lanes()
    # addr2line -i prints a (function, file:line) pair per inlining level; one address at a time keeps the parsing simple.  The
    # counted address is the return address of emu_xchg: minus one byte is inside the call instruction.
    by_line = collections.Counter(); by_fn = collections.Counter()
    for a, cnt in rows:
        fr = subprocess.run(["addr2line", "-f", "-i", "-C", "-e", lib, hex(int(a, 16) - 1)], capture_output=True, text=True).stdout.strip().split("\n")
        frames = [(fr[i], fr[i + 1]) for i in range(0, len(fr) - 1, 2)]
        core = [(f, l) for f, l in frames if "window_core.cuh" in l]
        # innermost frame inside window_core.cuh that is not one of the collective wrappers
        def lineno(l):
            try:
                return int(l.split(":")[-1].split(" ")[0])
            except ValueError:
                return 0
        site = next(((f, l) for f, l in core if lineno(l) > 140), core[0] if core else ("?", "?"))      # the wrappers are defined above line 140
        by_line[(site[0].split("(")[0], site[1].split("/")[-1].split(" ")[0])] += int(cnt)
        by_fn[site[0].split("(")[0]] += int(cnt)
    tot = sum(by_fn.values())
    print("k %d depth %d windows %d: %.0f collectives per window, statuses ok %d" % (k, depth, n, tot / n, int((r[0]["status"] == 1).sum())))
    for (fn, line), c in by_line.most_common(25):
        print("  %8.0f / window  %5.1f %%  %-28s %s" % (c / n, 100.0 * c / tot, fn, line))


if __name__ == "__main__":
    main()
