"""2+ GPU diagnostic (torchrun): how long do the pieces of one multi-GPU bench step take on each rank?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from scanner_colmap_b200 import SiftMatcher, synth, sharding
import bench
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
sp = sharding.plan([8192] * (100 * world), 10, world, rank)
own = list(range(*sp.own)); halo = [r for r, _ in sp.recv]
imgs = [synth.make_image(i, 8192) for i in own]
m = SiftMatcher(device=local); m.put_images(own, imgs)
buf = {r: torch.empty(8192 * 128, dtype=torch.uint8, device=dev) for r in halo}
def view(row):
    p, n = m.image_device_ptr(row); return torch.as_tensor(bench._DevView(p, n * 128), device=dev)
def ex():
    sharding.exchange_halo(sp, view, lambda r: buf[r]); torch.cuda.current_stream().synchronize()
def put():
    if halo: m.put_images_device(halo, [buf[r].data_ptr() for r in halo], [8192] * len(halo))
def ex_async():
    sharding.exchange_halo(sp, view, lambda r: buf[r])
    if halo: m.put_images_device_async(halo, [buf[r].data_ptr() for r in halo], [8192] * len(halo), torch.cuda.current_stream().cuda_stream)
for _ in range(3): ex(); put(); m.match_pairs_count(sp.pairs)
ts = {"exchange": [], "put": [], "match": [], "barrier": []}
for _ in range(10):
    t0 = time.perf_counter(); dist.barrier(); torch.cuda.synchronize(); t1 = time.perf_counter()
    ex(); t2 = time.perf_counter(); put(); t3 = time.perf_counter(); m.match_pairs_count(sp.pairs); t4 = time.perf_counter()
    for k, v in zip(ts, (t2 - t1, t3 - t2, t4 - t3, t1 - t0)): ts[k].append(v * 1e3)
print(f"rank {rank}: own {len(own)} halo {len(halo)} pairs {len(sp.pairs)} " + " ".join(f"{k}={np.median(v):.3f}ms" for k, v in ts.items()), flush=True)
ta = {"enqueue": [], "match": []}
for _ in range(3): ex_async(); m.match_pairs_count(sp.pairs)
for _ in range(10):
    dist.barrier(); torch.cuda.synchronize(); t1 = time.perf_counter()
    ex_async(); t2 = time.perf_counter(); m.match_pairs_count(sp.pairs); t3 = time.perf_counter()
    ta["enqueue"].append((t2 - t1) * 1e3); ta["match"].append((t3 - t2) * 1e3)
print(f"rank {rank} async: " + " ".join(f"{k}={np.median(v):.3f}ms" for k, v in ta.items()), flush=True)
m.clear_images(); m.put_images(own, imgs); ex(); put()
tm = []
for _ in range(10):
    dist.barrier(); torch.cuda.synchronize(); t1 = time.perf_counter(); m.match_pairs_count(sp.pairs); tm.append((time.perf_counter() - t1) * 1e3)
print(f"rank {rank} resident halo, match only: {np.median(tm):.3f}ms", flush=True)
dist.barrier(); dist.destroy_process_group()
