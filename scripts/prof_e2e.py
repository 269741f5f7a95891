"""Host-side breakdown of the end-to-end call (amod_osp_build with host pointers + amod_pool_fetch into host
buffers) on the bench snapshots.  Run on a GPU box: python scripts/prof_e2e.py"""
import glob
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amod_b200 import marshal, pool as apool, synth  # noqa: E402


def main():
    files = sorted(glob.glob(os.path.join(ROOT, "bench_data", "cfg2_epoch*.npz")))
    eps = [marshal.Epoch.from_npz_dict(np.load(f), "ep_") for f in files]
    tt = synth.make_road_graph(seed=0).tt
    b = apool.PoolBuilder(tt)
    for i in range(16):
        b.build_pool(eps[i % len(eps)], reuse=True)
    tb = tf = 0.0
    n = 64
    for i in range(n):
        ep = eps[i % len(eps)]
        t0 = time.perf_counter()
        b.build(ep)
        t1 = time.perf_counter()
        b.fetch(True)
        t2 = time.perf_counter()
        tb += t1 - t0; tf += t2 - t1
    print(f"build(host ptrs) {tb / n * 1e3:.3f} ms  fetch(host bufs) {tf / n * 1e3:.3f} ms  stats ms_build={b.last_stats.get('ms_build')}")


if __name__ == "__main__":
    main()
