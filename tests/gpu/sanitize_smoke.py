#!/usr/bin/env python
"""Small end-to-end pass over every kernel family for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tests/gpu/sanitize_smoke.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import __graft_entry__ as ge  # noqa: E402

ge.smoke()
from oracle import pasta as O  # noqa: E402
from vdf_b200 import msm as G, minroot as M  # noqa: E402

# plain layout, record levels (tiny S via many entries), batch path via the running prover is in smoke()
cv = O.VESTA
g = G.Generators.progression(cv.cid, 3, 5, 3000, table=False)
sc = [(7 * i + 1) % cv.order if i % 3 else i % 2 for i in range(3000)]
assert g.commit(sc) == cv.msm_known_dlog(sc, 3, 5)
w = M.VestaVDF().step_witness_batch([M.State(5, 6, 40)], 8)
assert len(w[0]) == 33
print("sanitize smoke ok")
