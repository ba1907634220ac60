"""CPU only, planning data: which workspace fields a window really touches.  The kernel source (single-lane emulation, tests/emu/emu.cpp) is
compiled with -fsanitize=thread and linked against tests/emu/trace_rt.cpp instead of libtsan, so that every load / store into the per-warp
workspace slab is attributed to its field (DCU_WS_FIELDS of window_core.cuh).  Prints, per field and per window of the bench workload: the
distinct 32-byte sectors (= bytes the memory system has to deliver at least once), loads and stores (element accesses: a lane-parallel loop
counts once per element), and the bytes per access -- the fields with many accesses on few bytes are the ones a shared-memory resident
part of the workspace should hold.
   python tools/field_traffic.py [coverage] [windows]"""
import ctypes as C
import os
import re
import subprocess
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import default_params, alloc_out, _ptr   # noqa: E402
from daccord_b200.host import Dataset               # noqa: E402


def field_names():
    s = open(os.path.join(ROOT, "daccord_b200", "csrc", "window_core.cuh")).read()
    a = s.index("#define DCU_WS_FIELDS(X)"); b = s.index("static inline void make_layout")
    return [(m.group(1), m.group(2)) for m in re.finditer(r"\bX\((\w+),\s*([\w ]+?),", s[a:b]) if m.group(1) != "name"]


def build():
    bd = os.path.join(ROOT, "tests", "emu", "_build")
    os.makedirs(bd, exist_ok=True)
    out = os.path.join(bd, "libemu_trace.so")
    o1, o2 = os.path.join(bd, "emu_trace.o"), os.path.join(bd, "trace_rt.o")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-fsanitize=thread", "-DDCU_EMU_TRACE", "-c", os.path.join(ROOT, "tests", "emu", "emu.cpp"), "-o", o1])
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-c", os.path.join(ROOT, "tests", "emu", "trace_rt.cpp"), "-o", o2])
    subprocess.check_call(["/usr/bin/g++", "-shared", "-o", out, o1, o2])       # no -fsanitize here: trace_rt.cpp is the runtime
    return out


def main():
    cov = float(sys.argv[1]) if len(sys.argv) > 1 else 40.0
    nw = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
    lib = C.CDLL(build())
    lib.trace_report.restype = C.c_uint64
    ds = Dataset.simulate(60000, read_len=10000, coverage=cov, seed=1)
    pi, pd, cor = ds.profile()
    p = default_params(p_i=pi, p_d=pd, est_cor=cor)
    b = ds.pile(nthreads=8)
    win = b.win[:: max(1, len(b.win) // nw)].copy(); sl = b.sl.copy(); packed = np.ascontiguousarray(ds.packed())
    res, cons, ops = alloc_out(len(win)); nov = C.c_uint64(0)
    rc = lib.emu_run_batch(C.byref(p), _ptr(packed), _ptr(win), C.c_uint64(len(win)), _ptr(sl), _ptr(res), _ptr(cons), _ptr(ops), C.c_int(0), C.byref(nov))
    assert rc == 0
    names = field_names()
    rep = np.zeros(4 * len(names), np.uint64)
    n = lib.trace_report(_ptr(rep), C.c_int(len(names)))
    rep = rep.reshape(-1, 4).astype(np.float64) / n
    tot_b = rep[:, 2].sum() * 32
    print("coverage %.0f, %d windows (tier-0 layout), consensus %d: %.1f KB of distinct 32-byte sectors per window, %.0f loads and %.0f stores per window"
          % (cov, n, int((res["status"] == 1).sum()), tot_b / 1024, rep[:, 0].sum(), rep[:, 1].sum()))
    print("%-12s %-10s %10s %7s %9s %9s %12s %14s" % ("field", "type", "bytes/win", "%", "loads", "stores", "acc / byte", "load-first B"))
    order = np.argsort(-rep[:, 2])
    cum = 0.0
    for i in order:
        if rep[i, 2] == 0:
            continue
        bts = rep[i, 2] * 32; cum += bts
        print("%-12s %-10s %10.0f %6.1f%% %9.0f %9.0f %12.2f %14.0f" % (names[i][0], names[i][1].strip(), bts, 100 * bts / tot_b, rep[i, 0], rep[i, 1], (rep[i, 0] + rep[i, 1]) / bts, rep[i, 3] * 32))
    # hot and small: accesses per byte
    print("\nhottest fields by accesses per touched byte (candidates for a shared-memory resident part):")
    dens = [((rep[i, 0] + rep[i, 1]) / (rep[i, 2] * 32), i) for i in range(len(names)) if rep[i, 2] > 0]
    acc_tot = rep[:, 0].sum() + rep[:, 1].sum(); cb = ca = 0.0
    for d, i in sorted(dens, reverse=True)[:25]:
        cb += rep[i, 2] * 32; ca += rep[i, 0] + rep[i, 1]
        print("  %-12s %8.0f B  %8.0f accesses  %6.2f acc/B   cumulative %6.1f KB hold %5.1f %% of all accesses" % (names[i][0], rep[i, 2] * 32, rep[i, 0] + rep[i, 1], d, cb / 1024, 100 * ca / acc_tot))


if __name__ == "__main__":
    main()
