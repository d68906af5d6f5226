"""Where does a whole smb_match_pairs call spend its time? (100 x 8192, overlap 10, 855 pairs)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scanner_colmap_b200 import SiftMatcher, synth, sequential_pairs
ids = list(range(100)); imgs = synth.make_images(100, 8192); pairs = sequential_pairs(ids, 10)
m = SiftMatcher(profile=True); m.put_images(ids, imgs)
for _ in range(3): m.match_pairs_count(pairs)
for _ in range(5):
    t0 = time.perf_counter(); tot = m.match_pairs_count(pairs); wall = (time.perf_counter() - t0) * 1e3
    t = m.timing()
    print(f"wall={wall:.3f}ms lib_total(dev events)={t['total_ms']:.3f} score={t['score_ms']:.3f} decide={t['decide_ms']:.3f} matches={tot}")
m.close()
m = SiftMatcher(profile=False); m.put_images(ids, imgs)
for _ in range(3): m.match_pairs_count(pairs)
t0 = time.perf_counter()
for _ in range(20): m.match_pairs_count(pairs)
print(f"no-profile wall per call = {(time.perf_counter()-t0)/20*1e3:.3f} ms")
