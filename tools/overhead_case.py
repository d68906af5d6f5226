"""Where does a resident 855-pair match call spend its time outside the score kernel?
python tools/overhead_case.py [n_images]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scanner_colmap_b200 import SiftMatcher, synth, sequential_pairs
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ids = list(range(n_img)); imgs = synth.make_images(n_img, 8192); pairs = sequential_pairs(ids, 10)
for prof in (True, False):
    m = SiftMatcher(profile=prof)
    m.put_images(ids, imgs)
    for _ in range(3):
        m.match_pairs_count(pairs)
    walls, begins = [], []
    for _ in range(10):
        t0 = time.perf_counter(); r = m.match_pairs_begin(pairs); t1 = time.perf_counter(); r.wait(); t2 = time.perf_counter(); r.release()
        walls.append((t2 - t0) * 1e3); begins.append((t1 - t0) * 1e3)
    t = m.timing()
    print(f"profile={prof}: wall {np.median(walls):.3f} ms (begin returns after {np.median(begins):.3f} ms) | device total {t['total_ms']:.3f} "
          f"score {t['score_ms']:.3f} runner_up {t['runner_up_ms']:.3f} decide {t['decide_ms']:.3f} "
          f"-> total - score {t['total_ms'] - t['score_ms']:.3f} | sub-batches {t['sub_batches']} launches {t['total_launches']} "
          f"plan_uploaded {t['plan_uploaded']}", flush=True)
    m.close()
