"""Throughput on a ragged set (BASELINE configs[3] shape at 1/10 of the image count): load balance check."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scanner_colmap_b200 import SiftMatcher, synth, sequential_pairs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
sizes = synth.ragged_sizes(n).tolist()
ids = list(range(n)); imgs = [synth.make_image(i, s) for i, s in zip(ids, sizes)]
pairs = sequential_pairs(ids, 10)
m = SiftMatcher(profile=True); m.put_images(ids, imgs)
for _ in range(3):
    t0 = time.perf_counter(); tot = m.match_pairs_count(pairs); wall = (time.perf_counter() - t0) * 1e3
    t = m.timing()
    print(f"{n} ragged images, {len(pairs)} pairs, sizes {min(sizes)}..{max(sizes)}: wall={wall:.2f}ms score={t['score_ms']:.2f}ms "
          f"TOPS={t['ops']/t['score_ms']/1e9:.0f} matches={tot} launches={t['score_launches']} "
          f"cta_busy_max_over_mean={t['cta_busy_max_over_mean']:.4f}", flush=True)
