"""Short single-GPU case for ncu: 20 images x 8192, overlap 10 (135 pairs), a few repetitions."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scanner_colmap_b200 import SiftMatcher, synth, sequential_pairs
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ids = list(range(n_img)); imgs = synth.make_images(n_img, 8192); pairs = sequential_pairs(ids, 10)
m = SiftMatcher(profile=True)
m.put_images(ids, imgs)
for _ in range(reps):
    tot = m.match_pairs_count(pairs); t = m.timing()
    print(f"{len(pairs)} pairs total={tot} score_ms={t['score_ms']:.3f} TOPS={t['ops']/t['score_ms']/1e9:.1f} pairs/s(score)={len(pairs)/t['score_ms']*1e3:.0f}", flush=True)
m.close()
