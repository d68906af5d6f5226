"""Op-level throughput: the SequentialMatchingCPU op driven the way integration/feature_matching.py
drives it (stencil range(0, overlap), packets of packet_size rows) through the fake Scanner dispatch, against the
C-ABI end-to-end rate on the same table.  python tools/op_bench.py [n_images] [n_desc] [overlap]
Prints one JSON line per packet size."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scanner_colmap_b200 import SiftMatcher, scanner_sim, synth, sequential_pairs
n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n_desc = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
overlap = int(sys.argv[3]) if len(sys.argv) > 3 else 10
ids = list(range(n_img)); descs = synth.make_images(n_img, n_desc)
kps = [np.zeros((n_desc, 6), np.float32) for _ in ids]
pairs = sequential_pairs(ids, overlap)
enc = scanner_sim.encode_table(ids, kps, descs)

# C-ABI end to end on the same table: pageable host descriptors in, matches in host memory out
m = SiftMatcher()
def abi_step():
    m.clear_images(); m.put_images(ids, descs); return m.match_pairs_count(pairs)
for _ in range(2): abi_step()
t0 = time.perf_counter(); reps = 5
for _ in range(reps): total = abi_step()
abi = len(pairs) * reps / (time.perf_counter() - t0)
m.close()
print(json.dumps({"what": "C ABI, pageable descriptors, one put_images + one match_pairs", "pairs_per_s": abi, "pairs": len(pairs), "matches": total}), flush=True)

rng = np.random.default_rng(0)
kps = [np.concatenate([rng.uniform(0, 4000, size=(n_desc, 2)), np.zeros((n_desc, 4))], axis=1).astype(np.float32) for _ in ids]
enc = scanner_sim.encode_table(ids, kps, descs)
for verify, packet in (("none", 4), ("none", 25), ("none", 100), ("gpu", 25), ("gpu", 100)):
    os.environ["SMB_OP_VERIFY"] = verify      # gpu: two-view verification of every pair (random keypoints: worst case)
    with scanner_sim.OpKernel() as k:
        k.run_table(ids, kps, descs, overlap=overlap, packet_size=packet, decode=False, encoded=enc)   # warm-up
        k.new_stream()
        t0 = time.perf_counter(); reps = 3
        for _ in range(reps):
            out_ids, out_tvg = k.run_table(ids, kps, descs, overlap=overlap, packet_size=packet, decode=False, encoded=enc)
            k.new_stream()
        dt = (time.perf_counter() - t0) / reps
    print(json.dumps({"what": "op through the fake Scanner dispatch", "verify": verify, "packet_size": packet, "rows_per_s": n_img / dt,
                      "pairs_per_s": len(pairs) / dt, "ratio_to_c_abi": len(pairs) / dt / abi, "ms_per_table": dt * 1e3,
                      "output_bytes": int(sum(out_tvg))}), flush=True)
