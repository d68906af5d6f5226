"""Wall-clock of whole match calls on resident descriptors (100 x 8192, overlap 10), for host-pipeline experiments."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scanner_colmap_b200 import SiftMatcher, synth, sequential_pairs
ids = list(range(100)); imgs = synth.make_images(100, 8192); pairs = sequential_pairs(ids, 10)
m = SiftMatcher(profile=True); m.put_images(ids, imgs)
for _ in range(3): m.match_pairs_count(pairs)
ts = []
for _ in range(20):
    t0 = time.perf_counter(); m.match_pairs_count(pairs); ts.append((time.perf_counter() - t0) * 1e3)
t = m.timing()
print(f"SMB_HEAD_FRACTION={os.environ.get('SMB_HEAD_FRACTION','-')}: wall median {np.median(ts):.3f} ms min {min(ts):.3f} | lib total {t['total_ms']:.3f} score {t['score_ms']:.3f} launches {t['total_launches']}")
