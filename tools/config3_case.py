"""BASELINE configs[2] on one GPU: 1000 images x 8192, overlap 20 -> 18,810 pairs in ONE call (sub-batched internally)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from concurrent.futures import ThreadPoolExecutor
from scanner_colmap_b200 import SiftMatcher, synth, sequential_pairs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
t0 = time.time()
with ThreadPoolExecutor(16) as ex: imgs = list(ex.map(lambda i: synth.make_image(i, 8192), range(n)))
print(f"generated {n} images in {time.time()-t0:.1f}s", flush=True)
ids = list(range(n)); pairs = sequential_pairs(ids, 20)
m = SiftMatcher(profile=True); m.put_images(ids, imgs)
for _ in range(2):
    t0 = time.perf_counter(); tot = m.match_pairs_count(pairs); wall = (time.perf_counter() - t0) * 1e3
    t = m.timing()
    print(f"{len(pairs)} pairs: wall={wall:.1f}ms ({len(pairs)/wall*1e3:.0f} pairs/s) score={t['score_ms']:.1f}ms launches={t['score_launches']} "
          f"TOPS={t['ops']/t['score_ms']/1e9:.0f} matches={tot}", flush=True)
sub = pairs[::997]
got = m.match_pairs(sub); back = m.match_pairs(sub[:, ::-1].copy())
ok = all(np.array_equal(a, b[:, ::-1][np.argsort(b[:, 1], kind='stable')]) for a, b in zip(got, back))
print("swap symmetry on", len(sub), "pairs:", ok)
