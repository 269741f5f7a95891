"""Multi-rank pool assembly over peer-mapped memory (amod_gpool_*): two ranks (two GPUs when the box
has them, otherwise two processes sharing GPU 0 — cudaIpc works either way), each builds its
interleaved vehicle slice and scatters it into every rank's global pool; both must hold the
reference's full vehicle-major pool.  The tiny collectives use gloo here so the test also runs on
one GPU; bench.py uses NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    two_gpus = torch.cuda.device_count() >= world
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from amod_b200 import dist as adist, marshal, synth
    from amod_b200.pool import PoolBuilder
    dev_index = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev_index)
    dev = torch.device("cuda", dev_index)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    tt = synth.make_road_graph(900, seed=4).tt
    b = PoolBuilder(tt, device=dev_index)
    b.set_stream(stream.cuda_stream)
    ok = True
    why = ""
    pp = None
    for seed in (1, 2, 3):
        ep = synth.random_epoch(tt, n_veh=101, n_pending=50, cap=4, mode=marshal.MODE_OSP, seed=seed, wait_sec=300, near=True, T=4000)
        ref = oracle.osp_build(ep, tt)
        sub, _sel = ep.shard(rank, world)
        b.build(sub)
        if pp is None:
            pp = adist.PeerPool(b, world, rank, ep.n_veh, 4 * ref.n_entries, 4 * ref.trip_req.size + 64, 4 * ref.sche_code.size + 64, dev)
            # with one GPU per rank also exercise the flag-based assembly (kernels of different ranks wait on each
            # other, which must never be done by two processes sharing one GPU)
            pp.fused = two_gpus and seed >= 2
        pp.fused = two_gpus and seed >= 2
        pp.assemble()
        got = pp.fetch()
        o, w = marshal.pools_equal(got, ref, rtol=0.0, check_kind=True)
        ok &= o
        why = why or w
        dist.barrier()    # nobody starts the next epoch's scatter while a peer still reads
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([int(ok)]))
    if not ok:
        open(os.path.join(out_dir, f"why{rank}.txt"), "w").write(why)
    dist.destroy_process_group()


def test_two_ranks_scatter_over_peer_memory(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        ok = int(np.load(tmp_path / f"ok{r}.npy")[0])
        why = (tmp_path / f"why{r}.txt").read_text() if not ok else ""
        assert ok, f"rank {r}: {why}"


def _single_rank_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from amod_b200 import dist as adist, marshal, synth
    from amod_b200.pool import PoolBuilder
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    tt = synth.make_road_graph(900, seed=4).tt
    b = PoolBuilder(tt, device=0)
    ok, why, pp = True, "", None
    for seed in (5, 6, 7):
        ep = synth.random_epoch(tt, n_veh=157, n_pending=60, cap=4, mode=marshal.MODE_OSP, seed=seed, wait_sec=300, near=True, T=4000)
        ref = oracle.osp_build(ep, tt)
        b.build(ep)
        if pp is None:
            pp = adist.PeerPool(b, 1, 0, ep.n_veh, 4 * ref.n_entries, 4 * ref.trip_req.size + 64, 4 * ref.sche_code.size + 64, dev)
        pp.fused = True          # push -> flags -> offsets -> merge; with one rank no kernel waits on another process
        pp.assemble()
        o, w = marshal.pools_equal(pp.fetch(), ref, rtol=0.0, check_kind=True)
        ok &= o
        why = why or w
    np.save(os.path.join(out_dir, "ok.npy"), np.array([int(ok)]))
    if not ok:
        open(os.path.join(out_dir, "why.txt"), "w").write(why)
    dist.destroy_process_group()


def test_flag_based_assembly_kernels_with_one_rank(tmp_path):
    """The NVLink assembly path (bulk push, epoch flags, offsets, entry-parallel merge) run by a single rank on one
    GPU: the same kernels bench.py uses at N > 1, checked against the oracle's full pool."""
    mp.spawn(_single_rank_worker, args=(1, _free_port(), str(tmp_path)), nprocs=1, join=True)
    ok = int(np.load(tmp_path / "ok.npy")[0])
    assert ok, (tmp_path / "why.txt").read_text() if not ok else ""


def _attached_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from amod_b200 import dist as adist, marshal, synth
    from amod_b200.pool import PoolBuilder
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    tt = synth.make_road_graph(900, seed=4).tt
    b = PoolBuilder(tt, device=rank)
    ok, why, pp = True, "", None
    rng = np.random.default_rng(17)
    for it, seed in enumerate((11, 12, 13, 14)):
        ep = synth.random_epoch(tt, n_veh=203, n_pending=70, cap=4, mode=marshal.MODE_OSP, seed=seed, wait_sec=300, near=True, T=4000)
        ref = oracle.osp_build(ep, tt)
        root = it % world                                    # the gathering rank changes too
        if it < 2:     # interleaved placement, then an arbitrary load-balanced one
            veh_rank, veh_local, sels = None, None, [np.arange(r, ep.n_veh, world) for r in range(world)]
        else:
            veh_rank, veh_local, sels = adist.balanced_placement(rng.pareto(1.1, size=ep.n_veh) + 1.0, world)
        sub = ep.take(sels[rank])
        if pp is None:
            b.build(sub)
            pp = adist.PeerPool(b, world, rank, ep.n_veh, 4 * ref.n_entries, 4 * ref.trip_req.size + 64, 4 * ref.sche_code.size + 64, dev)
        pp.attach(root=root, veh_rank=veh_rank, veh_local=veh_local)
        for _rep in range(3):                                # graph replay: the assembly kernels are nodes of the epoch graph
            b.build(sub)
            if rank == root:
                o, w = marshal.pools_equal(pp.fetch(), ref, rtol=0.0, check_kind=True)
                ok &= o
                why = why or w
        dist.barrier()
    # SBA on the same peers: request-major merge on the root (sbam.cuh), a non-monotone search list, OSP and SBA epochs alternating
    for it, seed in enumerate((21, 22, 23)):
        ep = synth.random_epoch(tt, n_veh=203, n_pending=60, cap=4, mode=marshal.MODE_SBA, seed=seed, wait_sec=400, near=True, frac_new=0.8, T=4000)
        if it == 1:
            ep.search = np.ascontiguousarray(ep.search[::-1])
        ref = oracle.sba_build(ep, tt)
        veh_rank, veh_local, sels = adist.balanced_placement(rng.pareto(1.1, size=ep.n_veh) + 1.0, world)
        pp.attach(root=0, veh_rank=veh_rank, veh_local=veh_local)
        b.build(ep.take(sels[rank]))
        if rank == 0:
            o, w = marshal.pools_equal(pp.fetch(), ref, rtol=0.0, check_kind=True)
            ok &= o
            why = why or ("sba: " + w if w else "")
        dist.barrier()
        ep2 = synth.random_epoch(tt, n_veh=203, n_pending=40, cap=4, mode=marshal.MODE_OSP, seed=seed + 100, wait_sec=300, near=True, T=4000)
        b.build(ep2.take(sels[rank]))
        if rank == 0:
            o, w = marshal.pools_equal(pp.fetch(), oracle.osp_build(ep2, tt), rtol=0.0, check_kind=True)
            ok &= o
            why = why or ("osp after sba: " + w if w else "")
        dist.barrier()
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([int(ok)]))
    if not ok:
        open(os.path.join(out_dir, f"why{rank}.txt"), "w").write(why)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs: kernels of different ranks wait on each other")
def test_two_gpus_gather_to_root_inside_the_build_graph(tmp_path):
    """bench.py's N > 1 path on real peers: PeerPool.attach -> every build pushes its slice to the root over NVLink and the
    root merges, all inside the captured epoch graph; interleaved and load-balanced placements, both ranks as root."""
    world = 2
    mp.spawn(_attached_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        ok = int(np.load(tmp_path / f"ok{r}.npy")[0])
        why = (tmp_path / f"why{r}.txt").read_text() if not ok else ""
        assert ok, f"rank {r}: {why}"
