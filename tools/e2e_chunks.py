"""Sweep the chunk boundaries of the pipelined end-to-end step (upload chunk k+1 overlaps matching group k)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scanner_colmap_b200 import SiftMatcher, synth, sequential_pairs
ids = list(range(100)); imgs = [torch.from_numpy(synth.make_image(i, 8192)).pin_memory().numpy() for i in ids]
pairs = sequential_pairs(ids, 10)
m = SiftMatcher()
def T(f, n=10):
    f(); f(); m.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    m.synchronize(); return (time.perf_counter() - t0) / n * 1e3
m.put_images(ids, imgs)
print("match all (resident) %.3f ms" % T(lambda: m.match_pairs_count(pairs)))
for bounds in ([0, 25, 50, 75, 100], [0, 12, 40, 70, 100], [0, 11, 33, 66, 100], [0, 11, 55, 100], [0, 16, 58, 100], [0, 11, 100],
               [0, 11, 30, 53, 76, 100], [0, 8, 20, 40, 60, 80, 100]):
    n = len(bounds) - 1
    groups = [[] for _ in range(n)]
    for a, b in pairs.tolist():
        c = max(k for k in range(n) if bounds[k] <= max(a, b))
        groups[c].append((a, b))
    groups = [np.asarray(g, dtype=np.uint32).reshape(-1, 2) for g in groups]
    def step():
        m.clear_images()
        for c in range(n): m.put_images_async(ids[bounds[c]:bounds[c + 1]], imgs[bounds[c]:bounds[c + 1]])
        for g in groups:
            if len(g): m.match_pairs_count(g)
    print(bounds, [len(g) for g in groups], "e2e %.3f ms" % T(step), flush=True)
