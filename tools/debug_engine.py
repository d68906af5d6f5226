"""Bring-up helper (run on the GPU box): one engine per process so a trap in one cannot poison the other."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scanner_colmap_b200 import SiftMatcher, synth, sequential_pairs
from oracle import oracle

engine = sys.argv[1] if len(sys.argv) > 1 else "tcgen05"
big = len(sys.argv) > 2 and sys.argv[2] == "big"
print("engine", engine, flush=True)
m = SiftMatcher(engine=engine, profile=True)
print("filter", m.filter, flush=True)
cases = [(128, 256), (300, 260), (1000, 1500), (4096, 4096)]
for n1, n2 in cases:
    a = synth.make_image(0, n1, track_step=32); b = synth.make_image(1, n2, track_step=32)
    t0 = time.time(); got = m.match(a, b); dt = time.time() - t0
    want = oracle.match(a, b)
    ok = np.array_equal(got, want)
    print(f"{n1}x{n2}: got {len(got)} want {len(want)} equal={ok} wall={dt*1e3:.2f}ms timing={m.timing()}", flush=True)
    if not ok:
        sg = set(map(tuple, got.tolist())); sw = set(map(tuple, want.tolist()))
        print("  missing", sorted(sw - sg)[:10], "extra", sorted(sg - sw)[:10], flush=True)
if big:
    ids = list(range(20)); imgs = synth.make_images(20, 8192); pairs = sequential_pairs(ids, 10)
    m.put_images(ids, imgs)
    for rep in range(3):
        t0 = time.time(); tot = m.match_pairs_count(pairs); dt = time.time() - t0
        t = m.timing()
        print(f"20x8192 W=10: {len(pairs)} pairs total={tot} wall={dt*1e3:.2f}ms score_ms={t['score_ms']:.3f} decide_ms={t['decide_ms']:.3f} "
              f"cands={t['candidates']} TOPS={t['ops']/t['score_ms']/1e9:.1f} pairs/s(score)={len(pairs)/t['score_ms']*1e3:.0f}", flush=True)
m.close()
print("done", flush=True)
