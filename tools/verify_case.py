"""Throughput of the GPU two-view verification on the bench workload's matches (100 x 8192, overlap 10, 855 pairs)
with random keypoint positions (every pair is pure outliers geometrically: both RANSACs run to their trial caps --
the worst case for time)  and on geometrically consistent synthetic scenes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scanner_colmap_b200 import SiftMatcher, synth, sequential_pairs
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ids = list(range(n_img)); imgs = synth.make_images(n_img, 8192); pairs = sequential_pairs(ids, 10)
rng = np.random.default_rng(1)
m = SiftMatcher()
m.put_images(ids, imgs)
for i in ids:
    m.put_keypoints(i, rng.uniform(0, 4000, size=(8192, 2)).astype(np.float32))
r = m.match_pairs_result(pairs)
for _ in range(2):
    t0 = time.perf_counter(); r.verify(seed=3); dt = time.perf_counter() - t0
    cfg = np.bincount([r.tvg(k)["config"] for k in range(len(pairs))], minlength=7)
    print(f"{len(pairs)} pairs, {r.total} matches, random keypoints: verify {dt*1e3:.1f} ms ({len(pairs)/dt:.0f} pairs/s) configs {cfg.tolist()} "
          f"mean trials F {np.mean([r.tvg(k)['trials_F'] for k in range(len(pairs))]):.0f} H {np.mean([r.tvg(k)['trials_H'] for k in range(len(pairs))]):.0f}", flush=True)
r.release(); m.close()
