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
for bounds in ([0, 16, 40, 100], [0, 14, 36, 100], [0, 18, 45, 100], [0, 20, 50, 100], [0, 16, 36, 64, 100], [0, 16, 48, 100],
               [0, 14, 30, 56, 100], [0, 24, 56, 100], [0, 8, 20, 44, 100], [0, 4, 12, 28, 56, 100], [0, 6, 14, 30, 60, 100],
               [0, 12, 30, 60, 100], [0, 10, 20, 40, 70, 100]):
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
    def step1():
        m.clear_images()
        for c in range(n): m.put_images_async(ids[bounds[c]:bounds[c + 1]], imgs[bounds[c]:bounds[c + 1]])
        m.match_pairs_count(pairs)
    print(bounds, [len(g) for g in groups], "e2e per-chunk calls %.3f ms, one call %.3f ms" % (T(step), T(step1)), flush=True)
