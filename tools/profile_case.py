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
if int(os.environ.get('SMB_DEBUG_FLAGS','0'),0) & 32:
    import ctypes, numpy as np
    from scanner_colmap_b200 import matcher as M
    L = M.load_library(); L.smb_debug_clocks.argtypes=[ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    buf = np.zeros(148*16, dtype=np.int64); L.smb_debug_clocks(m._h, buf.ctypes.data, buf.size); b = buf.reshape(148,16)
    tiles = b[:,3].astype(float)
    names = ['mma wait b_full','mma wait t_empty','mma total','tiles','epi wait t_full','epi ld+max','epi push','epi total','prod wait b_empty','prod total']
    for k,nm in enumerate(names): print(f'{nm:20s} mean/tile={np.mean(b[:,k]/tiles):9.1f}  (sum mean {b[:,k].mean():.0f})')
m.close()
