import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scanner_colmap_b200 import SiftMatcher, synth, sequential_pairs
ids = list(range(100)); imgs = [torch.from_numpy(synth.make_image(i, 8192)).pin_memory().numpy() for i in ids]
pairs = sequential_pairs(ids, 10)
m = SiftMatcher()
def T(f, n=10):
    f(); m.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    m.synchronize(); return (time.perf_counter() - t0) / n * 1e3
def up_sync(): m.clear_images(); m.put_images(ids, imgs)
def up_async():
    m.clear_images()
    for c in range(4): m.put_images_async(ids[25*c:25*c+25], imgs[25*c:25*c+25])
    m.synchronize()
print("upload sync  %.3f ms" % T(up_sync)); print("upload async %.3f ms" % T(up_async))
bounds=[0,25,50,75,100]
groups=[[] for _ in range(4)]
for a,b in pairs.tolist(): groups[min(3,max(a,b)//25)].append((a,b))
groups=[np.asarray(g,dtype=np.uint32) for g in groups]
print("group sizes", [len(g) for g in groups])
up_sync()
print("match all    %.3f ms" % T(lambda: m.match_pairs_count(pairs)))
print("match groups %.3f ms" % T(lambda: [m.match_pairs_count(g) for g in groups]))
def e2e_seq(): up_sync(); m.match_pairs_count(pairs)
def e2e_pipe():
    m.clear_images()
    for c in range(4): m.put_images_async(ids[25*c:25*c+25], imgs[25*c:25*c+25])
    for g in groups: m.match_pairs_count(g)
print("e2e sequential %.3f ms" % T(e2e_seq)); print("e2e pipelined  %.3f ms" % T(e2e_pipe))
