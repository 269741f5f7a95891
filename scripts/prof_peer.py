import os, sys, time, glob
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from amod_b200 import dist as adist, marshal, synth, capi
from amod_b200.marshal import Epoch
from amod_b200.pool import PoolBuilder
import bench
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
tt = synth.make_road_graph(seed=0).tt
b = PoolBuilder(tt, device=local); b.set_stream(stream.cuda_stream)
eps = [Epoch.from_npz_dict(np.load(f), "ep_") for f in sorted(glob.glob(os.path.join(ROOT, 'bench_data/cfg2_*.npz')))]
subs = []
for ep in eps:
    full = bench.tile_fleet(ep, world); sub, sel = full.shard(rank, world); subs.append(sub)
b.build(subs[0]); P, NT, NS = b.pool_sizes()
pp = adist.PeerPool(b, world, rank, full.n_veh, 3*world*P, 3*world*NT+1024, 3*world*NS+1024, dev)
tb = ta = 0.0; n = 0
for it in range(40):
    sub = subs[it % len(subs)]
    dist.barrier(); torch.cuda.synchronize()
    t0=time.perf_counter(); b.build(sub); t1=time.perf_counter(); pp.assemble(); t2=time.perf_counter()
    if it >= 8: tb += t1-t0; ta += t2-t1; n += 1
res = torch.tensor([tb/n*1e3, ta/n*1e3], device=dev, dtype=torch.float64)
allr = [torch.zeros_like(res) for _ in range(world)]
dist.all_gather(allr, res)
if rank == 0:
    print("per-rank build ms:", [round(float(x[0]),3) for x in allr]); print("per-rank assemble ms (incl. waiting for peers):", [round(float(x[1]),3) for x in allr], flush=True)
dist.destroy_process_group()
