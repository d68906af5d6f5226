"""Small case for compute-sanitizer: BASELINE configs[0] shape (20 images x 4096, overlap 10 -> 135 pairs) by default,
checked against the oracle so that a detected hazard can be tied to a wrong answer.  python tools/sanitize_case.py [n_images] [n_desc]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scanner_colmap_b200 import SiftMatcher, synth, sequential_pairs
from oracle import oracle
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n_desc = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
ids = list(range(n_img)); imgs = synth.make_images(n_img, n_desc); pairs = sequential_pairs(ids, 10)
with SiftMatcher() as m:
    m.put_images(ids, imgs)
    got = m.match_pairs(pairs)
want, _ = oracle.match_many(imgs, [(int(a), int(b)) for a, b in pairs])
ok = all(np.array_equal(g, w) for g, w in zip(got, want))
print(f"{len(pairs)} pairs of {n_desc}x{n_desc}, {sum(len(g) for g in got)} matches, bit-exact vs oracle: {ok}", flush=True)
sys.exit(0 if ok else 1)
