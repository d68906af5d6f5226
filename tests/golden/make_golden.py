"""Generates tests/golden/matches_small.json from the CPU oracle (run once; commit the output).
The reference itself cannot be imported or built here (COLMAP/Scanner absent), so these vectors freeze the
oracle restatement, not reference output -- see the PARITY UNPINNED note in oracle/sift_match_oracle.c."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import oracle
from scanner_colmap_b200 import synth

cases = []
for (id1, n1, id2, n2, ts, mr, md, cc) in [
    (0, 200, 1, 180, 16, 0.8, 0.7, True), (0, 200, 1, 180, 16, 0.8, 0.7, False), (5, 333, 7, 129, 8, 0.9, 1.0, True),
    (2, 64, 3, 500, 8, 0.8, 0.7, True), (9, 257, 9, 257, 8, 0.8, 0.7, True)]:
    a = synth.make_image(id1, n1, track_step=ts); b = synth.make_image(id2, n2, track_step=ts)
    m = oracle.match(a, b, max_ratio=mr, max_distance=md, cross_check=cc)
    cases.append(dict(id1=id1, n1=n1, id2=id2, n2=n2, track_step=ts, max_ratio=mr, max_distance=md, cross_check=cc,
                      matches=m.tolist()))
json.dump({"generator": "tests/golden/make_golden.py", "cases": cases},
          open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "matches_small.json"), "w"))
print([len(c["matches"]) for c in cases])
