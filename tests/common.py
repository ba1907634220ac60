"""Shared test plumbing: ctypes views of the C ABI structs, loaders for the oracle / emulation /
product libraries, and a small synthetic window-batch generator (numpy)."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class DcuParams(C.Structure):
    _fields_ = [("w", C.c_uint32), ("k_lo", C.c_uint32), ("k_hi", C.c_uint32), ("min_cov", C.c_uint32),
                ("min_ff", C.c_int32), ("max_ff", C.c_int32), ("max_err", C.c_uint64),
                ("p_i", C.c_double), ("p_d", C.c_double), ("est_cor", C.c_double)]


SLICE_DT = np.dtype([("gpos", "<u4"), ("len", "<u2"), ("flags", "<u2")])
WINDOW_DT = np.dtype([("slice_begin", "<u4"), ("slice_cnt", "<u2"), ("reserved", "<u2"), ("aread", "<u4"), ("astart", "<u4")])
RESULT_DT = np.dtype([("status", "u1"), ("k", "u1"), ("ff", "i1"), ("clen", "u1"), ("err", "<u4"), ("nops", "<u2"),
                      ("ncand", "<u2"), ("elength", "<i4")])
assert SLICE_DT.itemsize == 8 and WINDOW_DT.itemsize == 16 and RESULT_DT.itemsize == 16
CONS_STRIDE, OPS_STRIDE = 64, 128


def default_params(**kw):
    p = DcuParams(w=40, k_lo=8, k_hi=8, min_cov=3, min_ff=0, max_ff=2, max_err=2**64 - 1, p_i=0.09, p_d=0.045, est_cor=0.85)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    return os.path.join(ROOT, "oracle", "_build", "liboracle.so")


def build_emu():
    out = os.path.join(ROOT, "tests", "emu", "_build", "libemu.so")
    src = os.path.join(ROOT, "tests", "emu", "emu.cpp")
    deps = [src] + [os.path.join(ROOT, "daccord_b200", "csrc", f) for f in ("window_core.cuh", "window_types.cuh", "host_tables.hpp", "host_caps.hpp", "pile_core.cuh", "pile_host.hpp", "vote_core.cuh", "vote_host.hpp")] + [os.path.join(ROOT, "tests", "emu", "emu_builds.hpp")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["/usr/bin/g++", "-O2", "-g", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", out, src])
    return out


def build_emu_lanes():
    out = os.path.join(ROOT, "tests", "emu", "_build", "libemu_lanes.so")
    src = os.path.join(ROOT, "tests", "emu", "emu_lanes.cpp")
    deps = [src] + [os.path.join(ROOT, "daccord_b200", "csrc", f) for f in ("window_core.cuh", "window_types.cuh", "host_tables.hpp", "host_caps.hpp")] + [os.path.join(ROOT, "tests", "emu", "emu_builds.hpp")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["/usr/bin/g++", "-O2", "-g", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", out, src])
    return out


_libs = {}


def oracle_lib():
    if "o" not in _libs:
        lib = C.CDLL(build_oracle())
        lib.oracle_run_batch.restype = C.c_double
        lib.oracle_get_tables.restype = C.c_int64
        _libs["o"] = lib
    return _libs["o"]


def emu_lib():
    if "e" not in _libs:
        lib = C.CDLL(build_emu())
        lib.emu_get_tables.restype = C.c_int64
        _libs["e"] = lib
    return _libs["e"]


def alloc_out(nwin):
    return (np.zeros(nwin, RESULT_DT), np.zeros(nwin * CONS_STRIDE, np.uint8), np.zeros(nwin * OPS_STRIDE, np.uint8))


def run_oracle(params, packed, win, sl, nthreads=1):
    res, cons, ops = alloc_out(len(win))
    t = oracle_lib().oracle_run_batch(C.byref(params), _ptr(packed), _ptr(win), C.c_uint64(len(win)), _ptr(sl), _ptr(res), _ptr(cons), _ptr(ops), C.c_int(nthreads))
    assert t >= 0
    return res, cons, ops, t


def run_emu(params, packed, win, sl, tier=0):
    res, cons, ops = alloc_out(len(win))
    nov = C.c_uint64(0)
    rc = emu_lib().emu_run_batch(C.byref(params), _ptr(packed), _ptr(win), C.c_uint64(len(win)), _ptr(sl), _ptr(res), _ptr(cons), _ptr(ops), C.c_int(tier), C.byref(nov))
    assert rc == 0
    return res, cons, ops, nov.value


def run_emu_lanes(params, packed, win, sl, tier=0, schedule=0, seed=1):
    """window_core.cuh as 32 cooperative fibers per warp (tests/emu/emu_lanes.cpp); schedule 0 ascending, 1 descending, >= 2 shuffled.
    Returns (res, cons, ops, overflows, collectives); raises on a lane deadlock or on lanes that disagree about the result record."""
    if "l" not in _libs:
        _libs["l"] = C.CDLL(build_emu_lanes())
    res, cons, ops = alloc_out(len(win))
    nov, ncoll = C.c_uint64(0), C.c_uint64(0)
    rc = _libs["l"].emu_lanes_run_batch(C.byref(params), _ptr(packed), _ptr(win), C.c_uint64(len(win)), _ptr(sl), _ptr(res), _ptr(cons), _ptr(ops), C.c_int(tier), C.byref(nov),
                                        C.c_int(schedule), C.c_uint64(seed), C.byref(ncoll))
    assert rc == 0, "lane emulation: %s" % {101: "deadlock (lanes diverged around a warp collective)", 102: "lanes disagree on the result record"}.get(rc, rc)
    return res, cons, ops, nov.value, ncoll.value


class TsanUnavailable(Exception):
    pass


def build_emu_tsan():
    out = os.path.join(ROOT, "tests", "emu", "_build", "emu_tsan")
    src = os.path.join(ROOT, "tests", "emu", "emu_tsan.cpp")
    deps = [src] + [os.path.join(ROOT, "daccord_b200", "csrc", f) for f in ("window_core.cuh", "window_types.cuh", "host_tables.hpp", "host_caps.hpp")] + [os.path.join(ROOT, "tests", "emu", "emu_builds.hpp")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        r = subprocess.run(["/usr/bin/g++", "-fsanitize=thread", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-pthread", "-o", out, src], capture_output=True, text=True)
        if r.returncode != 0:
            if "tsan" in r.stderr.lower():                    # no libtsan on this host
                raise TsanUnavailable(r.stderr[-500:])
            raise RuntimeError(r.stderr[-3000:])
    return out


def run_emu_tsan(params, packed, win, sl, tier=1, timeout=600):
    """window_core.cuh with the 32 lanes of a warp as OS threads under ThreadSanitizer (tests/emu/emu_tsan.cpp).
    Returns (res, cons, ops, tsan_report_text); the report is empty when no two lanes race."""
    import tempfile
    exe = build_emu_tsan()
    with tempfile.TemporaryDirectory() as tmp:
        fin, fout = os.path.join(tmp, "in.bin"), os.path.join(tmp, "out.bin")
        packed = np.ascontiguousarray(packed); win = np.ascontiguousarray(win); sl = np.ascontiguousarray(sl)
        with open(fin, "wb") as f:
            f.write(bytes(params)); f.write(np.array([packed.nbytes, len(win), len(sl)], np.uint64).tobytes())
            f.write(packed.tobytes()); f.write(win.tobytes()); f.write(sl.tobytes())
        env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=4")
        r = subprocess.run([exe, fin, fout, str(tier)], capture_output=True, text=True, timeout=timeout, env=env)
        if r.returncode not in (0, 66) and "FATAL: ThreadSanitizer" in r.stderr:    # the sanitizer runtime itself cannot start on this host (e.g. ASLR entropy): not a finding
            raise TsanUnavailable(r.stderr[-500:])
        assert r.returncode in (0, 66), (r.returncode, r.stderr[-2000:])       # 66: TSan found races (reported through the text)
        res, cons, ops = alloc_out(len(win))
        raw = open(fout, "rb").read()
        a, b = res.nbytes, res.nbytes + cons.nbytes
        res[:] = np.frombuffer(raw[:a], RESULT_DT); cons[:] = np.frombuffer(raw[a:b], np.uint8); ops[:] = np.frombuffer(raw[b:], np.uint8)
    return res, cons, ops, r.stderr


def emu_vote(win, res, cons, ops, w, producefull, minlen, packed, boff, rlen):
    """vote_core.cuh (the GPU pile vote) compiled for the host; returns (segments, chars)"""
    from daccord_b200 import SEGMENT_DT
    seg = np.zeros(len(win) + 16, SEGMENT_DT)
    chars = np.zeros(int(rlen.sum()) * 2 + 1024, np.uint8)
    ns, nc = C.c_uint64(0), C.c_uint64(0)
    packed = np.ascontiguousarray(packed)
    rc = emu_lib().emu_vote(_ptr(win), _ptr(res), _ptr(cons), _ptr(ops), C.c_uint64(len(win)), C.c_uint32(w), C.c_int(1 if producefull else 0), C.c_uint64(minlen),
                            _ptr(packed), _ptr(boff), _ptr(rlen), C.c_uint64(len(rlen)), _ptr(seg), C.c_uint64(len(seg)), _ptr(chars), C.c_uint64(len(chars)),
                            C.byref(ns), C.byref(nc))
    assert rc == 0, rc
    return seg[:ns.value], chars[:nc.value]


def random_placements(win, w, rng, pfail, pins):
    """random but well-formed window results (status, consensus, placement trace covering exactly w A bases) to stress the votes:
    clumps of failed windows (gaps, short runs), insertion runs, trailing insertions"""
    from daccord_b200 import RESULT_DT
    n = len(win)
    res = np.zeros(n, RESULT_DT); cons = np.zeros(n * CONS_STRIDE, np.uint8); ops = np.zeros(n * OPS_STRIDE, np.uint8)
    st = np.ones(n, np.uint8); i = 0
    while i < n:
        if rng.random() < pfail * 0.2:
            ln = int(rng.integers(1, 30)); st[i:i + ln] = 2; i += ln
        else:
            i += 1
    acgt = np.frombuffer(b"ACGT", np.uint8)
    for i in range(n):
        o, c, j = [], [], 0
        while j < w:
            while rng.random() < pins and len(c) < 60 and len(o) < 120:
                o.append(2); c.append(int(rng.integers(0, 4)))
            if len(c) >= 62 or len(o) >= 124:
                o.append(3); j += 1; continue
            x = rng.random()
            if x < 0.1:
                o.append(3)
            else:
                o.append(0 if x < 0.9 else 1); c.append(int(rng.integers(0, 4)))
            j += 1
        while rng.random() < pins * 2 and len(c) < 63 and len(o) < 127:
            o.append(2); c.append(int(rng.integers(0, 4)))
        res[i]["status"] = st[i]; res[i]["nops"] = len(o); res[i]["clen"] = len(c)
        ops[i * OPS_STRIDE:i * OPS_STRIDE + len(o)] = o
        if c:
            cons[i * CONS_STRIDE:i * CONS_STRIDE + len(c)] = acgt[np.array(c, dtype=np.int64)]
    return res, cons, ops


def get_tables(lib, fn, params, which, klimn=64):
    n = getattr(lib, fn)(C.byref(params), C.c_int(which), None, C.c_int64(0), C.c_int(klimn))
    out = np.zeros(n, np.float64)
    getattr(lib, fn)(C.byref(params), C.c_int(which), _ptr(out), C.c_int64(n), C.c_int(klimn))
    return out


def compare_results(a, b, what="results"):
    """bit-exact comparison of two (res, cons, ops) triples; returns list of differing window indices"""
    ra, ca, oa = a[:3]
    rb, cb, ob = b[:3]
    bad = []
    for i in range(len(ra)):
        same = ra[i] == rb[i]
        if same and ra[i]["status"] == 1:
            n, m = int(ra[i]["clen"]), int(ra[i]["nops"])
            same = (ca[i * CONS_STRIDE:i * CONS_STRIDE + n] == cb[i * CONS_STRIDE:i * CONS_STRIDE + n]).all() and \
                   (oa[i * OPS_STRIDE:i * OPS_STRIDE + m] == ob[i * OPS_STRIDE:i * OPS_STRIDE + m]).all()
        if not same:
            bad.append(i)
    return bad


# ---------------------------------------------------------------- synthetic window batches
_COMP = np.array([3, 2, 1, 0], np.uint8)


def _noisy(rng, t, p_ins, p_del, p_sub):
    out = []
    for b in t:
        while rng.random() < p_ins:
            out.append(rng.integers(4))
        x = rng.random()
        if x < p_del:
            continue
        if x < p_del + p_sub:
            out.append((b + 1 + rng.integers(3)) & 3)
        else:
            out.append(b)
    return np.array(out, np.uint8)


def pack_bases(codes):
    n = len(codes)
    pad = (-n) % 4
    c = np.concatenate([codes, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    return ((c[:, 0] << 6) | (c[:, 1] << 4) | (c[:, 2] << 2) | c[:, 3]).astype(np.uint8)


def synth_batch(nwin, depth, seed=1, w=40, p_ins=0.09, p_del=0.045, p_sub=0.015, repeat_frac=0.0, depth_jitter=0):
    """nwin independent windows: truth segment, an A window of exactly w bases and `depth` noisy B slices
    (about half stored reverse-complemented in the database).  Returns packed, win, sl, truths."""
    rng = np.random.default_rng(seed)
    allb = []
    pos = 0
    win = np.zeros(nwin, WINDOW_DT)
    sl = []
    truths = []
    for i in range(nwin):
        if rng.random() < repeat_frac:
            unit = rng.integers(0, 4, rng.integers(1, 7)).astype(np.uint8)
            truth = np.tile(unit, 80 // len(unit) + 1)[:w + 24]
            nmut = rng.integers(0, 4)
            truth = truth.copy()
            for _ in range(nmut):
                truth[rng.integers(len(truth))] = rng.integers(4)
        else:
            truth = rng.integers(0, 4, w + 24).astype(np.uint8)
        # A window: noisy copy, exactly w bases; t = truth bases it covers
        a = []
        t = 0
        while len(a) < w and t < len(truth):
            piece = _noisy(rng, truth[t:t + 1], p_ins, p_del, p_sub)
            a.extend(piece.tolist())
            t += 1
        a = np.array(a[:w], np.uint8)
        if len(a) < w:
            a = np.concatenate([a, rng.integers(0, 4, w - len(a)).astype(np.uint8)])
        seg = truth[:t]
        truths.append(seg)
        d = depth + (rng.integers(-depth_jitter, depth_jitter + 1) if depth_jitter else 0)
        d = max(0, d)
        win[i] = (len(sl), 1 + d, 0, i, 0)
        pos += int(rng.integers(0, 4))
        allb.append(np.zeros(pos - sum(len(x) for x in allb), np.uint8) if False else np.zeros(0, np.uint8))
        # A slice
        start = sum(len(x) for x in allb)
        allb.append(a)
        sl.append((start, len(a), 0))
        for _ in range(d):
            b = _noisy(rng, seg, p_ins, p_del, p_sub)
            if len(b) > 250:
                b = b[:250]
            start = sum(len(x) for x in allb) if False else None
            rc = int(rng.random() < 0.5)
            stored = (_COMP[b][::-1] if rc else b)
            allb.append(stored)
            sl.append((None, len(b), rc))
    # assign gpos cumulatively (cheap second pass)
    off = 0
    k = 0
    gp = []
    for x in allb:
        gp.append(off)
        off += len(x)
    # allb has one zero-length filler per window before its A slice; map slices to non-filler arrays
    arrs = [x for x in allb]
    sl_arr = np.zeros(len(sl), SLICE_DT)
    ai = 0
    si = 0
    offs = np.cumsum([0] + [len(x) for x in arrs])
    # walk: per window: filler, A, then d B's
    idx = 0
    for i in range(nwin):
        idx += 1  # filler
        cnt = int(win[i]["slice_cnt"])
        for j in range(cnt):
            sl_arr[si] = (offs[idx], sl[si][1], sl[si][2])
            idx += 1
            si += 1
    codes = np.concatenate(arrs) if arrs else np.zeros(0, np.uint8)
    packed = pack_bases(codes)
    packed = np.concatenate([packed, np.zeros(16, np.uint8)])
    return packed, win, sl_arr, truths
