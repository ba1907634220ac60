"""GPU parity tests (-m gpu): the CUDA library, called through its C ABI, against the oracle on the same
seeded batches; bit-exact on status, k, filterfreq, error, consensus bytes and placement trace."""
import numpy as np
import pytest
from common import default_params, synth_batch, run_oracle, compare_results, get_tables, oracle_lib

pytestmark = pytest.mark.gpu


def _engine(p):
    import daccord_b200 as d
    pp = d.Params.default(w=p.w, k_lo=p.k_lo, k_hi=p.k_hi, min_cov=p.min_cov, min_ff=p.min_ff, max_ff=p.max_ff, max_err=p.max_err,
                          p_i=p.p_i, p_d=p.p_d, est_cor=p.est_cor)
    return d.Engine(pp, 0)


def test_tables_match_oracle():
    p = default_params()
    e = _engine(p)
    for which in range(4):
        a = get_tables(oracle_lib(), "oracle_get_tables", p, which)
        b = e.tables(which)
        assert a.shape == b.shape and (a.view(np.uint64) == b.view(np.uint64)).all(), which
    e.close()


CASES = [
    ("d40", dict(depth=40, n=1500, seed=103, rf=0.0), {}),
    ("d10", dict(depth=10, n=800, seed=104, rf=0.0), {}),
    ("d3", dict(depth=3, n=400, seed=105, rf=0.0), {}),
    ("repeats", dict(depth=12, n=800, seed=106, rf=0.6), {}),
    ("repeats40", dict(depth=40, n=800, seed=116, rf=0.6), {}),
    ("gapfill", dict(depth=8, n=400, seed=107, rf=0.3), dict(min_ff=0, max_ff=0)),
    ("multik", dict(depth=12, n=300, seed=108, rf=0.3), dict(k_lo=6, k_hi=10)),
    ("k12", dict(depth=30, n=300, seed=109, rf=0.2), dict(k_lo=12, k_hi=12)),
    ("k14", dict(depth=30, n=200, seed=110, rf=0.2), dict(k_lo=14, k_hi=14)),
    ("ebound", dict(depth=20, n=400, seed=111, rf=0.2), dict(max_err=120)),
    ("w32", dict(depth=25, n=300, seed=112, rf=0.2), dict(w=32)),
    ("w56", dict(depth=25, n=300, seed=113, rf=0.2), dict(w=56)),
    ("deep200", dict(depth=200, n=100, seed=115, rf=0.3), {}),
    ("w59", dict(depth=20, n=200, seed=116, rf=0.2), dict(w=59)),
    ("k3", dict(depth=20, n=200, seed=117, rf=0.2), dict(k_lo=3, k_hi=3)),
    ("w8", dict(depth=15, n=200, seed=119, rf=0.0), dict(w=8, k_lo=4, k_hi=4)),
    ("deep400", dict(depth=400, n=30, seed=121, rf=0.2), {}),
]


@pytest.mark.parametrize("name,gen,kw", CASES, ids=[c[0] for c in CASES])
def test_cuda_matches_oracle(name, gen, kw):
    p = default_params(**kw)
    packed, win, sl, _ = synth_batch(gen["n"], gen["depth"], seed=gen["seed"], repeat_frac=gen["rf"], depth_jitter=min(gen["depth"], 3), w=p.w)
    ro = run_oracle(p, packed, win, sl, 8)
    e = _engine(p)
    e.set_reads(packed)
    rg = e.run(win, sl)
    bad = compare_results(ro, rg)
    st = e.stats()
    e.close()
    assert not bad, (name, len(bad), bad[:5], [(ro[0][i], rg[0][i]) for i in bad[:3]])
    assert st["launches"] >= 1


def test_repeatable_and_order_independent():
    p = default_params()
    packed, win, sl, _ = synth_batch(600, 20, seed=120, repeat_frac=0.3)
    e = _engine(p)
    e.set_reads(packed)
    a = e.run(win, sl)
    b = e.run(win, sl)
    assert not compare_results(a, b)
    perm = np.random.default_rng(1).permutation(len(win))
    c = e.run(win[perm].copy(), sl)
    inv = np.argsort(perm)
    c2 = (c[0][inv], c[1].reshape(-1, 64)[inv].reshape(-1), c[2].reshape(-1, 128)[inv].reshape(-1))
    assert not compare_results(a, c2)
    e.close()


def test_errors_are_loud():
    import daccord_b200 as d
    p = default_params()
    e = _engine(p)
    packed, win, sl, _ = synth_batch(4, 5, seed=1)
    with pytest.raises(d.DcuError):
        e.run(win, sl)                 # no database set
    e.set_reads(packed)
    bad = sl.copy(); bad[3]["gpos"] = 2**31
    with pytest.raises(d.DcuError):
        e.run(win, bad)                # slice outside the database
    with pytest.raises(d.DcuError):
        d.Engine(d.Params.default(w=100), 0)   # unsupported window size
    e.close()


def test_position_slot_cache_switch_changes_nothing_on_the_gpu(monkeypatch):
    """DCU_POSCACHE=0 (all unitig position slots recomputed for every (first,last) pair) against the default (slots of unsplit unitigs kept across
    the pairs of a traverse), on a repeat-rich shallow pile whose windows walk through many pairs: identical to each other and to the oracle."""
    p = default_params(k_lo=7, k_hi=9)
    packed, win, sl, _ = synth_batch(600, 10, seed=141, repeat_frac=0.6, depth_jitter=3)
    ro = run_oracle(p, packed, win, sl, 8)
    out = []
    for pc in ("0", "1"):
        monkeypatch.setenv("DCU_POSCACHE", pc)       # read by dcu_create
        e = _engine(p)
        e.set_reads(packed)
        out.append(e.run(win, sl))
        e.close()
    assert not compare_results(ro, out[0])
    assert not compare_results(ro, out[1])
    assert not compare_results(out[0], out[1])


def test_small_first_pass_capacities(monkeypatch):
    """DCU_T0_SMALL=1 shrinks the first-pass workspace to about the p99.9 of a clean 40x pile (slab 399 KB instead of 950 KB per warp); on a deep
    small-k repeat-rich batch a few windows overflow it and take the large-workspace launch: the results are still the oracle's."""
    p = default_params(k_lo=6, k_hi=6)
    packed, win, sl, _ = synth_batch(400, 60, seed=151, repeat_frac=0.6, depth_jitter=3)
    ro = run_oracle(p, packed, win, sl, 8)
    monkeypatch.setenv("DCU_T0_SMALL", "1")          # read by the host when the batch is uploaded
    e = _engine(p)
    e.set_reads(packed)
    rg = e.run(win, sl)
    st = e.stats()
    e.close()
    assert not compare_results(ro, rg)
    assert st["launches"] >= 1


def test_pile_shaped_batch_of_1e5_windows_matches_oracle_and_truth():
    """a real pile (1 Mb of A-reads at 40x, ~1.04e5 windows: overlapping windows of neighbouring positions, both builds of the kernel and the
    overflow passes in one launch) against the oracle on every window -- result record, consensus bytes and placement trace -- then the
    read-level GPU path (dcu_pile + launch + dcu_vote) against the simulated genome: an anchor outside the oracle (the reference quotes the
    error rate of its output against the true sequence, README.md:442)."""
    import os
    import sys
    import daccord_b200 as d
    from daccord_b200.host import Dataset, format_segments
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from bench import full_compare
    ds = Dataset.simulate(25000, read_len=10000, coverage=40, seed=7, keep_truth=True)
    batch = ds.pile(0, ds.nreads, w=40, a=10, nthreads=os.cpu_count() or 1)
    assert len(batch.win) >= 100000
    pi, pd, cor = ds.profile()
    p = default_params(p_i=pi, p_d=pd, est_cor=cor)
    packed = np.array(ds.packed(), copy=True)
    ro = run_oracle(p, packed, batch.win, batch.sl, os.cpu_count() or 1)
    e = _engine(p)
    e.set_reads(packed)
    rg = e.run(batch.win, batch.sl)
    st = e.stats()
    diff = full_compare(ro, rg)
    assert not diff.any(), ("windows that differ from the oracle", np.nonzero(diff)[0][:10], st)
    assert (rg[0]["status"] == 1).mean() > 0.97 and st["lost_windows"] == 0
    ovl, trace, boff, rlen = ds.overlaps()
    e.pile(ovl, trace, ds.tspace, boff, rlen)
    e.launch()
    seg, chars = e.vote()
    e.close()
    fasta, nseq = format_segments(seg, chars)
    assert fasta == batch.vote(*rg)[0]
    acc = ds.truth_eval(fasta)
    # raw reads: 15 % error events per base; corrected: a few 1e-4 (measured 3.9e-4 on this generator at 40x)
    assert acc["reads"] == ds.nreads and acc["corrected_bases"] > 0.95 * 1e6 and acc["raw_erate"] > 0.14
    assert acc["erate"] < 1e-3, acc


def test_shared_memory_build_matches_oracle(monkeypatch):
    """DCU_SMEM=1: the first pass runs the shared-memory build of the kernel (graph fields in the warp's shared-memory arena, slices of the next
    window staged by bulk copies, two-bitmap pre-filter); windows beyond its small capacities go on to the HBM passes.  Same results as the oracle,
    on clean deep piles (nearly everything stays in the first pass) and on repeat-rich ones (many hand-overs); without staging too."""
    for seed, depth, rf, stage in ((201, 40, 0.0, "1"), (202, 30, 0.5, "1"), (203, 12, 0.3, "0")):
        p = default_params()
        packed, win, sl, _ = synth_batch(1200, depth, seed=seed, repeat_frac=rf, depth_jitter=3)
        ro = run_oracle(p, packed, win, sl, 8)
        monkeypatch.setenv("DCU_SMEM", "1"); monkeypatch.setenv("DCU_STAGE", stage)
        e = _engine(p)
        e.set_reads(packed)
        rg = e.run(win, sl)
        st = e.stats()
        e.close()
        assert st["smem_warps"] >= 2 and st["lost_windows"] == 0
        assert not compare_results(ro, rg), (seed, st)


def test_hybrid_build_matches_oracle(monkeypatch):
    """DCU_HYBRID=1: first pass with the window's k-mer table (pre-filtered, 512 slots) in shared memory at 32 warps per SM, everything else in the HBM
    slab; windows whose table fills up or that need the filter frequency 1 graph go on to the plain HBM passes.  Same results as the oracle."""
    for seed, depth, rf in ((211, 40, 0.0), (212, 30, 0.4)):
        p = default_params()
        packed, win, sl, _ = synth_batch(1000, depth, seed=seed, repeat_frac=rf, depth_jitter=3)
        ro = run_oracle(p, packed, win, sl, 8)
        monkeypatch.setenv("DCU_HYBRID", "1")
        e = _engine(p)
        e.set_reads(packed)
        rg = e.run(win, sl)
        st = e.stats()
        e.close()
        assert st["smem_warps"] == 32 and st["lost_windows"] == 0
        assert not compare_results(ro, rg), (seed, st)
