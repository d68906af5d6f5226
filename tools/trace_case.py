"""Per-phase clock counts of the score kernel's MMA thread and epilogue warps (CTA 0), from a library built with
-DSMB_TRACE:  make -C scanner_colmap_b200/csrc OUT=../../tools/bin/libsmb_trace.so EXTRA=-DSMB_TRACE"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scanner_colmap_b200 import matcher, synth, sequential_pairs
os.environ['SMB_LIB'] = os.path.join(ROOT, "tools", "bin", os.environ.get("SMB_TRACE_LIB", "libsmb_trace.so"))
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ids = list(range(n_img)); imgs = synth.make_images(n_img, 8192); pairs = sequential_pairs(ids, 10)
m = matcher.SiftMatcher(profile=True)
m.put_images(ids, imgs)
for _ in range(2):
    tot = m.match_pairs_count(pairs); t = m.timing()
    print(f"{len(pairs)} pairs total={tot} score_ms={t['score_ms']:.3f} TOPS={t['ops']/t['score_ms']/1e9:.1f}", flush=True)
m.close()
