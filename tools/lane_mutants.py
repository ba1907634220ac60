"""CPU only: mutation test of the 32-lane emulation (tests/emu/emu_lanes.cpp).  Drops one wsync() site of window_core.cuh at a time
(DCU_EMU_SKIP_SYNC_LINE) and reports whether the harness notices -- a result differing from the oracle, lanes disagreeing, or a deadlock under
one of the lane schedules.  A site that survives every schedule is either redundant (the next collective already orders the accesses) or not
exercised by the data; the list is printed so that it can be reviewed by hand.
With --tsan the same mutants run in the ThreadSanitizer build (tests/emu/emu_tsan.cpp: lanes = OS threads, collectives = barriers), which
reports the racing source lines whatever the schedule: a dropped sync without a report there orders no conflicting accesses on this data.
   python tools/lane_mutants.py [windows_per_case] [--tsan]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
from common import default_params, synth_batch, run_oracle, run_emu_lanes, run_emu_tsan, compare_results
import re
n = int(sys.argv[1]); tsan = len(sys.argv) > 2 and sys.argv[2] == "tsan"
cases = [(dict(), dict(depth=30, seed=3, rf=0.2)), (dict(min_ff=0, max_ff=2, k_lo=6, k_hi=8), dict(depth=8, seed=7, rf=0.4)), (dict(), dict(depth=12, seed=6, rf=0.6))]
for kw, gen in cases:
    p = default_params(**kw)
    packed, win, sl, _ = synth_batch(n, gen["depth"], seed=gen["seed"], repeat_frac=gen["rf"], depth_jitter=3, w=p.w)
    ref = run_oracle(p, packed, win, sl, 4)
    if tsan:
        got = run_emu_tsan(p, packed, win, sl, 1)
        bad = compare_results(ref, got[:3])
        lines = sorted(set(int(x) for x in re.findall(r"window_core\.cuh:(\d+)", got[3]) if int(x) < 1550))
        if bad or got[3].count("WARNING: ThreadSanitizer"):
            print("DETECTED races %%d lines %%s mismatches %%s" %% (got[3].count("WARNING: ThreadSanitizer"), lines[:6], bad[:3])); sys.exit(3)
        continue
    for sched in (0, 1, 2, 3):
        got = run_emu_lanes(p, packed, win, sl, 1, sched, 11 + sched)
        bad = compare_results(ref, got)
        if bad:
            print("DETECTED mismatch schedule %%d windows %%s" %% (sched, bad[:4])); sys.exit(3)
print("SURVIVED")
""" % (ROOT, os.path.join(ROOT, "tests"))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    tsan = ["tsan"] if "--tsan" in sys.argv else []
    n = int(args[0]) if args else (6 if tsan else 24)
    src = open(os.path.join(ROOT, "daccord_b200", "csrc", "window_core.cuh")).read().split("\n")
    sites = [i + 1 for i, l in enumerate(src) if re.search(r"\bwsync\(\)", l.split("//")[0]) and "define" not in l and "inline" not in l]
    print("%d wsync() sites" % len(sites))
    env = dict(os.environ)
    env.pop("DCU_EMU_SKIP_SYNC_LINE", None)
    r = subprocess.run([sys.executable, "-c", CHILD, str(n)] + tsan, env=env, capture_output=True, text=True)
    assert "SURVIVED" in r.stdout, ("baseline must be clean", r.stdout, r.stderr)
    surv = []
    for ln in sites:
        env["DCU_EMU_SKIP_SYNC_LINE"] = str(ln)
        r = subprocess.run([sys.executable, "-c", CHILD, str(n)] + tsan, env=env, capture_output=True, text=True)
        out = (r.stdout.strip().split("\n") or [""])[-1]
        if r.returncode != 0 and "DETECTED" not in out:
            out = "DETECTED " + (r.stderr.strip().split("\n") or ["?"])[-1][:120]
        print("line %4d  %-60s | %s" % (ln, out[:60], src[ln - 1].strip()[:90]), flush=True)
        if out.startswith("SURVIVED"):
            surv.append(ln)
    print("%d of %d dropped syncs detected; surviving lines: %s" % (len(sites) - len(surv), len(sites), surv))


if __name__ == "__main__":
    main()
