"""CPU only: .las ingest -- sequential whole-file reader vs record-offset index + ranged, parallel decode.
   python tools/las_ingest_bench.py [Mb] [shards]"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from daccord_b200.host import Dataset  # noqa: E402


def main():
    mb = float(sys.argv[1]) if len(sys.argv) > 1 else 50
    shards = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    ds = Dataset.simulate(int(mb * 1e6 / 40), read_len=10000, coverage=40, seed=1)
    with tempfile.TemporaryDirectory() as tmp:
        las, db = os.path.join(tmp, "x.las"), os.path.join(tmp, "x.db")
        ds.write(las, db)
        size = os.path.getsize(las)
        t = time.time(); full = Dataset.load(las, db); t_full = time.time() - t
        t = time.time(); a = Dataset.load_range(las, db, 0, 10**9); t_first = time.time() - t        # builds + caches the index
        t = time.time(); a = Dataset.load_range(las, db, 0, 10**9); t_all = time.time() - t          # cached index, all host threads
        n = full.nreads; per = (n + shards - 1) // shards
        t = time.time(); s = Dataset.load_range(las, db, per, 2 * per); t_shard = time.time() - t
        assert a.novl == full.novl
        print("las %.1f MB, %d overlaps, %d reads" % (size / 1e6, full.novl, n))
        print("sequential whole file (incl. DB)      %.3f s  %.0f MB/s" % (t_full, size / 1e6 / t_full))
        print("index build + ranged parallel decode  %.3f s" % t_first)
        print("cached index, whole file, %3d threads %.3f s  %.0f MB/s" % (os.cpu_count(), t_all, size / 1e6 / t_all))
        print("cached index, shard 1 of %d (%d overlaps) %.3f s" % (shards, s.novl, t_shard))


if __name__ == "__main__":
    main()
