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
# round-2 paths: affine rounds on an arena, graph replay (second identical call), drop-in cache hit, witness bank ->
# commit, un-normalised partial sum, sum-check kernels
import os  # noqa: E402

import numpy as np  # noqa: E402

from vdf_b200 import _lib, nova as N, spartan as SP  # noqa: E402

lib = _lib.load()
os.environ["VDFGPU_MSM_AFFINE"] = "2"
ga = G.Generators.progression(0, 9, 4, 1 << 12, table=True, window_bits=7)
sa = [(11 * i + 3) % O.Q for i in range(1 << 12)]
first = ga.commit(sa)
assert first == O.PALLAS.msm_known_dlog(sa, 9, 4) and ga.commit(sa) == first      # second call: graph replay
os.environ.pop("VDFGPU_MSM_AFFINE")
pts = np.zeros(72 * 2048, dtype=np.uint8)
_lib.check(lib.vdfgpu_gens_export(ga._h, 0, 2048, pts.ctypes.data))
sb = np.frombuffer(O.fes_to_bytes(sa[:2048], O.Q), dtype=np.uint8).copy()
out = np.zeros(96, dtype=np.uint8)
for _ in range(2):                                                                  # miss, then hit
    lib.mult_pippenger_pallas(out.ctypes.data, pts.ctypes.data, 2048, sb.ctypes.data, True)
    assert O.jac_from_bytes(O.PALLAS, out.tobytes()) == O.PALLAS.msm_known_dlog(sa[:2048], 9, 4)
_lib.check(lib.vdfgpu_dropin_cache_clear())
ovdf = O.MinRootVDF(O.FIELD_FQ)
inst = [O.make_step_instance(O.FIELD_FQ, 6, ovdf.eval(O.State(3 + k, 0, 1), 6), aug_cons=20) for k in range(2)]
shape = inst[0][0]
off = shape.num_vars - 25
bank = N.WitnessBank(O.FIELD_FQ, [tuple(W[off - 3:off]) for _, W, _, _ in inst], 6)
gs = N.R1CSShape(O.FIELD_FQ, shape.num_cons, shape.num_vars, shape.num_io, shape.A, shape.B, shape.C)
gg = G.Generators.progression(0, 5, 3, max(shape.num_cons, shape.num_vars), table=True)
pr = N.RunningProver(gs, gg)
pr.set_running(inst[0][1], [0] * shape.num_cons, N.RelaxedR1CSInstance(None, None, list(inst[0][2]), 1))
cW, _, _ = pr.prove_step_bank_bytes(bank, 1, off, O.fes_to_bytes(inst[1][1], shape.m), O.fes_to_bytes(inst[1][2], shape.m), 77)
assert O.jac_from_bytes(O.PALLAS, cW) == O.PALLAS.msm_known_dlog(inst[1][1], 5, 3)
assert shape.is_sat_relaxed(*pr.get_running())
tabs = [[(i * 7 + k) % O.Q for i in range(64)] for k in range(4)]
ch = lambda rnd, e: (sum(e) + rnd + 5) % O.Q  # noqa: E731
assert SP.sumcheck(O.FIELD_FQ, tabs, ch)[2] == O.sumcheck_prove(tabs, O.Q, ch)[2]
assert SP.eq_evals(O.FIELD_FP, [3, 4, 5]) == O.eq_evals([3, 4, 5], O.P)
_lib.check(lib.vdfgpu_trim())
print("sanitize smoke ok")
