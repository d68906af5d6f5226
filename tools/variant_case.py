"""A/B harness for score-kernel build variants: python tools/variant_case.py <lib.so> [n_images] [reps]
(variants are built with  make -C scanner_colmap_b200/csrc OUT=../../tools/bin/libsmb_<name>.so EXTRA="-D...")"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scanner_colmap_b200 import matcher, synth, sequential_pairs
lib = sys.argv[1]
os.environ['SMB_LIB'] = lib if os.path.isabs(lib) else os.path.join(ROOT, lib)
n_img = int(sys.argv[2]) if len(sys.argv) > 2 else 20
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ids = list(range(n_img)); imgs = synth.make_images(n_img, 8192); pairs = sequential_pairs(ids, 10)
m = matcher.SiftMatcher(profile=True)
m.put_images(ids, imgs)
best = 0.0
for _ in range(reps):
    tot = m.match_pairs_count(pairs); t = m.timing()
    best = max(best, t['ops'] / t['score_ms'] / 1e9)
print(f"{os.path.basename(lib):24s} flags={os.environ.get('SMB_DEBUG_FLAGS','0')} {len(pairs)} pairs total={tot} best TOPS={best:.1f}", flush=True)
m.close()
