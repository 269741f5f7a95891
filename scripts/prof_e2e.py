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
    tag = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    files = sorted(glob.glob(os.path.join(ROOT, "bench_data", f"{tag}_epoch*.npz")))
    eps = [marshal.Epoch.from_npz_dict(np.load(f), "ep_") for f in files]
    tt = synth.make_road_graph(seed=0).tt
    b = apool.PoolBuilder(tt)
    for i in range(16):
        b.build_pool(eps[i % len(eps)], reuse=True)
    tb = tf = td = 0.0
    n = 64
    for i in range(n):
        ep = eps[i % len(eps)]
        t0 = time.perf_counter()
        b.build(ep)
        t1 = time.perf_counter()
        b.fetch(True)
        t2 = time.perf_counter()
        tb += t1 - t0; tf += t2 - t1; td += b.last_stats['ms_build']
    print(f"{tag}: build(host ptrs) {tb / n * 1e3:.3f} ms  fetch(host bufs) {tf / n * 1e3:.3f} ms  device time of the build (upload + graph) {td / n:.3f} ms"
          f"  evals/epoch {b.last_stats['n_evals']}  launches {b.last_stats['n_kernel_launches']}")


if __name__ == "__main__":
    main()
