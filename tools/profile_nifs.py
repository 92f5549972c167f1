#!/usr/bin/env python
"""One NIFS fold (t = 1024 step circuit + synthetic 9.8k-constraint block, as bench.py's extra.nifs_fold) a few
times: a short target for an ncu launch list of the LATENCY path (Nova-size batched MSM + cross-term + fold)."""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vdf_b200 import _lib, encoding as E, msm as G, nova as N, synthetic as S  # noqa: E402

_lib.check(_lib.load().vdfgpu_init(0))
t, aug = 1024, 9800
cons, nvars, io, A, B, C, W, X = S.step_instance(E.FQ, t, aug, seed=42)
gs = N.R1CSShape(E.FQ, cons, nvars, io, A, B, C)
gens = G.Generators.progression(0, 12345, 678, max(cons, nvars), table=True)
prover = N.RunningProver(gs, gens)
Wb, Xb = E.fes_to_bytes(W, E.Q), E.fes_to_bytes(X, E.Q)
prover.set_running(W, [0] * cons, N.RelaxedR1CSInstance(None, None, list(X), 1))
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for _ in range(2):
    prover.prove_step_bytes(Wb, Xb, 0x1234567890ABCDEF)
t0 = time.perf_counter()
for _ in range(reps):
    prover.prove_step_bytes(Wb, Xb, 0x1234567890ABCDEF)
print(json.dumps({"nifs_fold_ms": (time.perf_counter() - t0) / reps * 1e3, "cons": cons, "vars": nvars,
                  "window_bits": gens.window_bits(max(cons, nvars))}))
