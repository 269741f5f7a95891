import os, sys, time, glob
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from amod_b200 import dist as adist, marshal, synth, capi
from amod_b200.marshal import Epoch
from amod_b200.pool import PoolBuilder
sys.path.insert(0, ROOT)
import bench
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
tt = synth.make_road_graph(seed=0).tt
b = PoolBuilder(tt, device=local); b.set_stream(stream.cuda_stream)
ep = Epoch.from_npz_dict(np.load(sorted(glob.glob(os.path.join(ROOT, 'bench_data/cfg2_*.npz')))[1]), "ep_")
full = bench.tile_fleet(ep, world); sub, sel = full.shard(rank, world)
b.build(sub); P, NT, NS = b.pool_sizes()
pp = adist.PeerPool(b, world, rank, full.n_veh, 3*world*P, 3*world*NT+1024, 3*world*NS+1024, dev)
T = {"build":0,"counts":0,"allgather":0,"scatter":0,"barrier":0}
def sync(): torch.cuda.synchronize()
for it in range(25):
    t0=time.perf_counter(); b.build(sub); sync(); t1=time.perf_counter()
    pp.counts.zero_(); b._check(b.lib.amod_pool_veh_counts(b.h, pp.counts.data_ptr())); sync(); t2=time.perf_counter()
    dist.all_gather_into_tensor(pp.counts_all, pp.counts); sync(); t3=time.perf_counter()
    b._check(b.lib.amod_gpool_scatter(b.h, pp.counts_all.data_ptr(), pp.vmax, pp.n_veh_total)); sync(); t4=time.perf_counter()
    dist.barrier(); sync(); t5=time.perf_counter()
    if it >= 5:
        for k,v in zip(T, (t1-t0,t2-t1,t3-t2,t4-t3,t5-t4)): T[k]+=v
if rank == 0: print({k: round(v/20*1e3,3) for k,v in T.items()}, flush=True)
dist.destroy_process_group()
