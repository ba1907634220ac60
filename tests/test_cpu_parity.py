"""CPU-side tests (no GPU): oracle sanity, host table builder vs oracle tables, the kernel source run as a
single-lane host emulation vs the oracle, and the C-ABI export list."""
import ctypes as C
import os
import re
import subprocess
import sys
import numpy as np
import pytest
from common import (ROOT, default_params, synth_batch, run_oracle, run_emu, run_emu_lanes, run_emu_tsan, TsanUnavailable, compare_results, get_tables, oracle_lib, emu_lib,
                    CONS_STRIDE)


def test_tables_bit_identical():
    for kw in ({}, {"w": 32, "p_i": 0.12, "p_d": 0.03}, {"w": 56, "k_lo": 6, "k_hi": 10, "est_cor": 0.9}):
        p = default_params(**kw)
        for which in range(6):
            a = get_tables(oracle_lib(), "oracle_get_tables", p, which)
            b = get_tables(emu_lib(), "emu_get_tables", p, which)
            assert a.shape == b.shape and (a.view(np.uint64) == b.view(np.uint64)).all(), (kw, which)


def test_oracle_recovers_truth_at_40x():
    p = default_params()
    packed, win, sl, truths = synth_batch(200, 40, seed=21)
    res, cons, ops, _ = run_oracle(p, packed, win, sl, 4)
    assert (res["status"] == 1).all()
    exact = 0
    for i, t in enumerate(truths):
        c = bytes(cons[i * CONS_STRIDE:i * CONS_STRIDE + int(res[i]["clen"])]).decode()
        exact += c == "".join("ACGT"[x] for x in t)
    assert exact >= 195


def test_oracle_scoring_distance_equals_dp():
    """the bit-parallel edit distance of the oracle's candidate scoring (the CPU arm's hot loop) is the DP's number on random pairs, every length
    0..70 on both sides (64 is the word limit, longer pairs take the DP) and on near-identical pairs like candidate vs slice"""
    lib = oracle_lib()
    rng = np.random.default_rng(5)
    out = (C.c_uint64 * 2)()
    for it in range(4000):
        la, lb = int(rng.integers(0, 71)), int(rng.integers(0, 71))
        a = rng.integers(0, 4, la).astype(np.uint8) + 65
        if it % 2 and la:
            b = a.copy()[:lb] if lb <= la else np.concatenate([a, rng.integers(0, 4, lb - la).astype(np.uint8) + 65])
            for _ in range(int(rng.integers(0, 8))):
                if len(b):
                    b[int(rng.integers(0, len(b)))] = 65 + int(rng.integers(0, 4))
        else:
            b = rng.integers(0, 4, lb).astype(np.uint8) + 65
        lib.oracle_edit_distances(a.ctypes.data_as(C.c_void_p), C.c_uint64(len(a)), b.ctypes.data_as(C.c_void_p), C.c_uint64(len(b)), out)
        assert out[0] == out[1], (it, la, lb, out[0], out[1])


CASES = [
    ("d40", dict(depth=40, n=250, seed=3, rf=0.0), {}),
    ("d10", dict(depth=10, n=250, seed=4, rf=0.0), {}),
    ("d3", dict(depth=3, n=200, seed=5, rf=0.0), {}),
    ("repeats", dict(depth=12, n=300, seed=6, rf=0.6), {}),
    ("gapfill", dict(depth=8, n=200, seed=7, rf=0.3), dict(min_ff=0, max_ff=0)),
    ("multik", dict(depth=12, n=150, seed=8, rf=0.3), dict(k_lo=6, k_hi=10)),
    ("k12", dict(depth=30, n=120, seed=9, rf=0.2), dict(k_lo=12, k_hi=12)),
    ("ebound", dict(depth=20, n=200, seed=11, rf=0.2), dict(max_err=120)),
    ("w32", dict(depth=25, n=200, seed=12, rf=0.2), dict(w=32)),
    ("w56", dict(depth=25, n=150, seed=13, rf=0.2), dict(w=56)),
    ("nocor", dict(depth=25, n=150, seed=14, rf=0.2), dict(est_cor=0.0)),
    ("deep", dict(depth=200, n=40, seed=15, rf=0.3), {}),
    ("w59", dict(depth=20, n=80, seed=16, rf=0.2), dict(w=59)),          # largest window the 64-bit aligners take
    ("k3", dict(depth=20, n=80, seed=17, rf=0.2), dict(k_lo=3, k_hi=3)),  # smallest k
    ("k14", dict(depth=25, n=80, seed=18, rf=0.1), dict(k_lo=14, k_hi=14)),
    ("w8", dict(depth=15, n=80, seed=19, rf=0.0), dict(w=8, k_lo=4, k_hi=4)),
]


@pytest.mark.parametrize("name,gen,kw", CASES, ids=[c[0] for c in CASES])
def test_kernel_emulation_matches_oracle(name, gen, kw):
    p = default_params(**kw)
    packed, win, sl, _ = synth_batch(gen["n"], gen["depth"], seed=gen["seed"], repeat_frac=gen["rf"], depth_jitter=min(gen["depth"], 3), w=p.w)
    ro = run_oracle(p, packed, win, sl, 4)
    for tier in (1, 0, 2, 3):                # 2 = the shared-memory build, 3 = the hybrid build (small capacities: what overflows goes to the HBM passes)
        re_ = run_emu(p, packed, win, sl, tier)
        bad = [i for i in compare_results(ro, re_) if re_[0][i]["status"] != 250]   # 250 = tier overflow, re-run in the next pass by the product
        assert not bad, (name, tier, bad[:5])
        if tier == 1:
            assert re_[3] == 0


@pytest.mark.parametrize("name,gen,kw", CASES, ids=[c[0] for c in CASES])
def test_lane_emulation_matches_oracle_under_every_schedule(name, gen, kw):
    """The kernel source as 32 cooperative fibers per warp (tests/emu/emu_lanes.cpp): the lanes run between two warp collectives in
    ascending, descending and shuffled order.  A missing __syncwarp, a collective in divergent code or lanes that disagree on the
    per-window state would show as a mismatch against the oracle, a deadlock or a lane disagreement (tools/lane_mutants.py measures
    how many dropped syncs this notices)."""
    p = default_params(**kw)
    n = {"k14": 8, "multik": 24, "gapfill": 40}.get(name, max(10, min(120, 3000 // gen["depth"])))     # the three are slow (filter-frequency descent, gap filling)
    packed, win, sl, _ = synth_batch(n, gen["depth"], seed=gen["seed"] + 1000, repeat_frac=gen["rf"], depth_jitter=min(gen["depth"], 3), w=p.w)
    ro = run_oracle(p, packed, win, sl, 4)
    for tier, schedule in ((1, 0), (1, 1), (1, 2), (0, 3), (0, 1), (2, 0), (2, 1), (2, 4)):      # tier 2 = the shared-memory build
        rl = run_emu_lanes(p, packed, win, sl, tier, schedule, seed=gen["seed"])
        bad = [i for i in compare_results(ro, rl) if rl[0][i]["status"] != 250]
        assert not bad, (name, tier, schedule, bad[:5])
        assert tier != 1 or rl[3] == 0
        assert rl[4] > 0 or (tier == 2 and rl[3] == len(win))    # collectives were executed: this was the 32-lane build (unless every window left the small shared-memory tier at once)


TSAN_CASES = [("d30", dict(depth=30, n=8, seed=31, rf=0.2), {}),
              ("gapfill", dict(depth=8, n=10, seed=37, rf=0.4), dict(min_ff=0, max_ff=2, k_lo=6, k_hi=8)),
              ("repeats", dict(depth=12, n=10, seed=36, rf=0.6), {}),
              ("tier0", dict(depth=40, n=6, seed=38, rf=0.0), {}),
              ("smem", dict(depth=40, n=6, seed=39, rf=0.0), {}),
              ("smem_ff", dict(depth=10, n=12, seed=40, rf=0.5), dict(min_ff=0, max_ff=2))]


@pytest.mark.parametrize("name,gen,kw", TSAN_CASES, ids=[c[0] for c in TSAN_CASES])
def test_no_race_between_lanes_under_thread_sanitizer(name, gen, kw):
    """The kernel source with the 32 lanes of a warp as OS threads (collectives and __syncwarp = barriers, atomics = atomics) under
    ThreadSanitizer: no two lanes may touch the same workspace bytes between two barriers unless both only read -- and the result is still the
    oracle's.  (tools/lane_mutants.py --tsan: dropping a needed wsync() is reported with the two racing source lines.)"""
    p = default_params(**kw)
    packed, win, sl, _ = synth_batch(gen["n"], gen["depth"], seed=gen["seed"], repeat_frac=gen["rf"], depth_jitter=3, w=p.w)
    ro = run_oracle(p, packed, win, sl, 4)
    tier = 0 if name == "tier0" else (2 if name.startswith("smem") else 1)
    try:
        res, cons, ops, report = run_emu_tsan(p, packed, win, sl, tier)
    except TsanUnavailable as e:
        pytest.skip("ThreadSanitizer cannot run on this host: %s" % str(e)[-200:])
    assert "ThreadSanitizer" not in report, report[:3000]
    bad = [i for i in compare_results(ro, (res, cons, ops)) if res[i]["status"] != 250]
    assert not bad, (name, bad[:5])


def test_position_slot_cache_switch_changes_nothing():
    """DCU_POSCACHE=0 (every (first,last) pair recomputes all unitig position slots, the round-1 behaviour) and the default (slots of unsplit
    unitigs kept across the pairs of a traverse) give identical results on a repeat-rich shallow pile, where windows walk through many pairs."""
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "import numpy as np\n"
            "from common import default_params, synth_batch, run_emu, run_emu_lanes\n"
            "p = default_params(k_lo=7, k_hi=9)\n"
            "packed, win, sl, _ = synth_batch(80, 10, seed=41, repeat_frac=0.6, depth_jitter=3, w=p.w)\n"
            "r = run_emu(p, packed, win, sl, 1); l = run_emu_lanes(p, packed, win, sl, 1, 2, 5)\n"
            "assert (r[0] == l[0]).all() and (r[1] == l[1]).all() and (r[2] == l[2]).all()\n"
            "np.save(sys.argv[1], np.concatenate([r[0].view(np.uint8).ravel(), r[1], r[2]]))\n") % (ROOT, os.path.join(ROOT, "tests"))
    import tempfile
    outs = []
    with tempfile.TemporaryDirectory() as tmp:
        for pc in ("0", "1"):
            out = os.path.join(tmp, "r%s.npy" % pc)
            subprocess.check_call([sys.executable, "-c", code, out], env=dict(os.environ, DCU_POSCACHE=pc))
            outs.append(np.load(out))
    assert (outs[0] == outs[1]).all()


def test_trace_runtime_attributes_workspace_accesses():
    """tests/emu/trace_rt.cpp (a stand-in for the TSan runtime that attributes every workspace access to its field; tools/field_traffic.py):
    results unchanged under instrumentation, and the hash table, the instance bytes and the unitig slots all show up with plausible byte counts."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("field_traffic", os.path.join(ROOT, "tools", "field_traffic.py"))
    ft = importlib.util.module_from_spec(spec); spec.loader.exec_module(ft)
    lib = C.CDLL(ft.build()); lib.trace_report.restype = C.c_uint64
    p = default_params()
    packed, win, sl, _ = synth_batch(40, 30, seed=71, repeat_frac=0.1, depth_jitter=3, w=p.w)
    from common import alloc_out, _ptr
    res, cons, ops = alloc_out(len(win)); nov = C.c_uint64(0)
    assert lib.emu_run_batch(C.byref(p), _ptr(packed), _ptr(win), C.c_uint64(len(win)), _ptr(sl), _ptr(res), _ptr(cons), _ptr(ops), C.c_int(0), C.byref(nov)) == 0
    ro = run_oracle(p, packed, win, sl, 4)
    assert not compare_results(ro, (res, cons, ops))
    names = [n for n, _ in ft.field_names()]
    rep = np.zeros(4 * len(names), np.uint64)
    n = lib.trace_report(_ptr(rep), C.c_int(len(names)))
    assert n == len(win)
    per = dict(zip(names, rep.reshape(-1, 4)[:, 2].astype(np.float64) * 32 / n))
    assert 2000 < per["hkey"] < 60000 and per["ipos"] > 100 and per["scs"] > 1000 and 20000 < sum(per.values()) < 300000, per


def test_edge_cases_empty_and_ragged():
    p = default_params()
    packed, win, sl, _ = synth_batch(30, 6, seed=31)
    # ragged: empty window, window with only the A slice, zero-length B slices
    win = win.copy(); sl = sl.copy()
    win[0]["slice_cnt"] = 0
    win[1]["slice_cnt"] = 1
    sl[win[2]["slice_begin"] + 1]["len"] = 0
    sl[win[2]["slice_begin"] + 2]["len"] = 3          # shorter than k
    ro = run_oracle(p, packed, win, sl, 1)
    re_ = run_emu(p, packed, win, sl, 1)
    assert not compare_results(ro, re_)
    assert ro[0][0]["status"] == 0 and ro[0][1]["status"] == 0


def test_long_and_degenerate_slices():
    """slices of the maximum length (255), slices shorter than k, identical slices, and a 400-deep pile"""
    p = default_params()
    rng = np.random.default_rng(5)
    from common import pack_bases, WINDOW_DT, SLICE_DT
    truth = rng.integers(0, 4, 300).astype(np.uint8)
    wins, sls, pos = [], [], 0
    chunks = []
    def add(seq):
        nonlocal pos
        chunks.append(seq); start = pos; pos += len(seq); return start
    # window 0: A window + 30 long slices (the first 255 bases of noisy copies) -- elength fallbacks, no consensus expected
    s0 = len(sls); sls.append((add(truth[:40]), 40, 0))
    for _ in range(30):
        b = truth[:255].copy(); b[rng.integers(0, 255, 20)] = rng.integers(0, 4, 20); sls.append((add(b), 255, 0))
    wins.append((s0, 31, 0, 0, 0))
    # window 1: identical slices (a clean pile) ; window 2: slices shorter than k ; window 3: 400 identical-ish slices
    s1 = len(sls); sls.append((add(truth[:40]), 40, 0))
    for _ in range(12): sls.append((add(truth[:40]), 40, 0))
    wins.append((s1, 13, 0, 1, 0))
    s2 = len(sls); sls.append((add(truth[:40]), 40, 0))
    for _ in range(8): sls.append((add(truth[:5]), 5, 0))
    wins.append((s2, 9, 0, 2, 0))
    s3 = len(sls); sls.append((add(truth[:40]), 40, 0))
    for _ in range(399):
        b = truth[:40].copy(); b[rng.integers(0, 40)] = rng.integers(0, 4); sls.append((add(b), 40, int(rng.integers(0, 2)) * 0))
    wins.append((s3, 400, 0, 3, 0))
    packed = np.concatenate([pack_bases(np.concatenate(chunks)), np.zeros(16, np.uint8)])
    win = np.array(wins, dtype=WINDOW_DT); sl = np.array(sls, dtype=SLICE_DT)
    ro = run_oracle(p, packed, win, sl, 2)
    re_ = run_emu(p, packed, win, sl, 1)
    assert not compare_results(ro, re_), (ro[0], re_[0])
    assert ro[0][1]["status"] == 1 and bytes(ro[1][64:64 + 40]) == bytes(b"ACGT"[x] for x in truth[:40])
    assert ro[0][3]["status"] == 1


def test_c_abi_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "daccord_b200.h")).read()
    declared = set(re.findall(r"\b(dcu_[a-z_]+)\s*\(", hdr))
    assert {"dcu_create", "dcu_run", "dcu_set_reads", "dcu_upload", "dcu_launch", "dcu_download"} <= declared
    import daccord_b200
    if not os.path.exists(daccord_b200.LIB_PATH):
        from daccord_b200 import build
        build.build()
    lib = C.CDLL(daccord_b200.LIB_PATH)      # loads without a GPU; no compute call is made
    for name in declared:
        assert hasattr(lib, name), name


def test_product_has_no_oracle_or_cpu_path():
    csrc = os.path.join(ROOT, "daccord_b200")
    for dp, _, files in os.walk(csrc):
        if "_build" in dp:
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".hpp", ".cpp", ".py", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r'#include\s*[<"][^>"]*oracle|liboracle|oracle_run|import\s+oracle|from\s+oracle', txt), f


def test_gap_fill_table_check_is_uniform_across_lanes():
    """found by tools/fuzz_parity.py (seed 5031): gap_fill compared hstate[0] + extras with the table size lane by lane while lanes that had
    passed the check were already inserting extras (which counts hstate[0] up) -- slower lanes could take the overflow branch alone.  w = 59,
    depth 60, k 7..8, filter frequencies down to 0: window 7 deadlocked the 32-lane emulation under every schedule."""
    p = default_params(w=59, k_lo=7, k_hi=8, min_cov=3, max_ff=2, min_ff=0, p_i=0.12523667712661934, p_d=0.06831091479633782, est_cor=0.0, max_err=556)
    e = 0.22770304932112606
    packed, win, sl, _ = synth_batch(66, 60, seed=5031, w=59, p_ins=e * 0.55, p_del=e * 0.3, p_sub=e * 0.15, repeat_frac=0.6, depth_jitter=1)
    ro = run_oracle(p, packed, win, sl, 4)
    for sched in (0, 2):
        rl = run_emu_lanes(p, packed, win, sl, 1, sched, 5031)
        assert not compare_results(ro, rl)
