"""Generates the committed golden fixtures from the CPU oracle (the reference itself cannot be built here: libmaus2 is
absent, see oracle/README.md -- so these vectors pin the oracle, "parity unpinned" upstream).
  kat1.{las,db,..}   BASELINE config 1: genome 1 kb, 21 reads of 1 kb covering it fully (1 A-read pile with 20 B-reads), tspace 100
  kat1.fasta         oracle FastA for `-w40 -a10 -k8 -I0,0` ; kat1_all.fasta for all 21 reads
  windows_small.npz  a seeded window batch (packed DB, windows, slices) with the oracle's per-window results
Run from the repo root:  python tests/golden/make_golden.py
"""
import ctypes as C
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import default_params, synth_batch, run_oracle, oracle_lib
from daccord_b200.host import Dataset


def oracle_fasta(p, las, db, first, last, a=10, threads=4):
    lib = oracle_lib()
    lib.oracle_daccord_files.restype = C.c_void_p
    n = C.c_uint64(0)
    st = (C.c_uint64 * 3)()
    ptr = lib.oracle_daccord_files(C.byref(p), C.c_uint32(a), C.c_uint64(2**64 - 1), C.c_uint64(5000), C.c_int(0), C.c_uint64(0), las.encode(), db.encode(),
                                   C.c_int64(first), C.c_int64(last), C.c_int(threads), C.byref(n), st)
    out = C.string_at(ptr, n.value)
    lib.oracle_free(C.c_void_p(ptr))
    return out, [int(x) for x in st]


def main():
    ds = Dataset.simulate(1000, read_len=1000, coverage=21, seed=1)
    las, db = os.path.join(HERE, "kat1.las"), os.path.join(HERE, "kat1.db")
    ds.write(las, db)
    pi, pd, cor = ds.profile()
    p = default_params(p_i=pi, p_d=pd, est_cor=cor)
    fa, st = oracle_fasta(p, las, db, 0, 0)
    open(os.path.join(HERE, "kat1.fasta"), "wb").write(fa)
    fa_all, st_all = oracle_fasta(p, las, db, 0, -1)
    open(os.path.join(HERE, "kat1_all.fasta"), "wb").write(fa_all)
    print("kat1: reads", ds.nreads, "overlaps", ds.novl, "windows/attempted/ok", st, "all", st_all)
    p2 = default_params()
    packed, win, sl, _ = synth_batch(400, 18, seed=2024, repeat_frac=0.4, depth_jitter=3)
    res, cons, ops, _ = run_oracle(p2, packed, win, sl, 4)
    np.savez_compressed(os.path.join(HERE, "windows_small.npz"), packed=packed, win=win, sl=sl, res=res, cons=cons, ops=ops)
    print("windows_small: status", np.bincount(res["status"], minlength=3))


if __name__ == "__main__":
    main()
