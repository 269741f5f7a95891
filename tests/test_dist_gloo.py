"""N > 1 host logic on CPU: two gloo ranks each build their interleaved vehicle slice (with the
oracle standing in for the per-rank GPU build) and all-gather the pool; every rank must end up with
the reference's vehicle-major pool."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_golden
from amod_b200 import marshal


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, family, case, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from amod_b200 import dist as adist
    from conftest import load_golden as lg
    tt, cases = lg(family)
    ep, gold = cases[case]
    sub, sel = ep.shard(rank, world)
    sba = ep.mode == marshal.MODE_SBA
    local = (oracle.sba_build if sba else oracle.osp_build)(sub, tt)
    if sba:   # a rank's SBA slice carries its own empty pairs; they merge by vehicle
        pass
    full = adist.allgather_pool(local, sel, ep.n_veh, mode_sba=sba, search=ep.search)
    ok, why = marshal.pools_equal(full, gold, check_kind=True)
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([int(ok)]))
    if not ok:
        open(os.path.join(out_dir, f"why{rank}.txt"), "w").write(why)
    dist.destroy_process_group()


@pytest.mark.parametrize("family,case", [("rand_real_osp_cap4", 0), ("sim_osp_cap4", 2), ("rand_ties_sba_cap4", 1)])
def test_two_rank_gloo_allgather_reproduces_reference_pool(tmp_path, family, case):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), family, case, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        ok = int(np.load(tmp_path / f"ok{r}.npy")[0])
        why = (tmp_path / f"why{r}.txt").read_text() if not ok else ""
        assert ok, f"rank {r}: {why}"


def test_sba_merge_follows_the_search_order_not_the_request_index():
    """the reference's SBA pool is ordered by position in new_received_rids (dispatch_sba.py:56); a
    non-monotone search list must merge to the single-rank pool."""
    import oracle
    from amod_b200 import dist as adist
    tt, cases = load_golden("rand_ties_sba_cap4")
    ep, _gold = cases[1]
    ep.search = np.ascontiguousarray(ep.search[::-1])
    whole = oracle.sba_build(ep, tt)
    parts = []
    for rank in range(3):
        sub, sel = ep.shard(rank, 3)
        parts.append((oracle.sba_build(sub, tt), sel))
    ok, why = marshal.pools_equal(adist.merge_sba(parts, ep.n_veh, ep.search), whole, check_kind=True)
    assert ok, why
