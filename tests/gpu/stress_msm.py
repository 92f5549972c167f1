#!/usr/bin/env python
"""Randomised cross-configuration parity on the GPU at sizes the Python oracle cannot reach: for several sizes,
point sets (distinct points; few distinct points, i.e. buckets full of equal points: doublings and cancellations)
and scalar distributions (uniform; Nova-like: half bits, some small, some full-size; few distinct values), the
result bytes must be identical across generator layouts (table / plain: different window size, bucket sets and
reduction path) and across 0..5 batched-affine halving rounds.  Where the discrete logs are known the result is
also checked against (sum s_i k_i) G computed with Python integers.  Usage: python tests/gpu/stress_msm.py [log2n ...]"""
import json
import os
import random
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import pasta as O  # noqa: E402  (checker only: this is a test tool)
from vdf_b200 import _lib, msm as G  # noqa: E402

lib = _lib.load()
_lib.check(lib.vdfgpu_init(0))
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
_lib.check(lib.vdfgpu_set_stream(st.cuda_stream))
sizes = [int(a) for a in sys.argv[1:]] or [15, 17, 19, 21]
py = random.Random(int(os.environ.get("VDF_STRESS_SEED", "2024")))
failures = 0


def scalars(kind, n, order, seed):
    """n Montgomery representatives < order as an (n, 4) uint64 array (little-endian limbs)."""
    rs = np.random.RandomState(seed)
    limbs = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64)
    limbs[:, 7] &= (1 << 29) - 1                      # < 2^253 < order
    if kind == "nova":
        sel = rs.rand(n)
        limbs[sel < 0.5, 1:] = 0
        limbs[sel < 0.5, 0] &= 1                      # half of the witness: bits
        small = (sel >= 0.5) & (sel < 0.8)
        limbs[small, 2:] = 0                          # 64-bit values
    elif kind == "few":
        vals = rs.randint(0, 1 << 32, size=(7, 8), dtype=np.uint64)
        vals[:, 7] &= (1 << 29) - 1
        limbs = vals[rs.randint(0, 7, size=n)]
    out = (limbs[:, 0::2] | (limbs[:, 1::2] << np.uint64(32))).astype(np.uint64)
    return np.ascontiguousarray(out)


def to_int(rows):
    return [int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192 for r in rows]


for lg in sizes:
    n = 1 << lg
    for curve in (0, 1):
        cv = O.CURVES[curve]
        for pts_kind, (k0, d) in (("distinct", (py.randrange(1, 1 << 60), py.randrange(1, 1 << 40))), ("equal", (py.randrange(1, 1 << 60), 0))):
            for sc_kind in ("uniform", "nova", "few"):
                sc = scalars(sc_kind, n, cv.order, seed=lg * 100 + curve * 10 + len(sc_kind))
                dev = torch.from_numpy(sc.view(np.int64)).cuda()
                # known discrete logs: P_i = (k0 + i d) G  =>  result = (sum s_i (k0 + i d)) G
                if lg <= 19:
                    # the C ABI takes scalars in Montgomery form: these limbs ARE the representatives s * 2^256
                    rinv = pow(1 << 256, -1, cv.order)
                    ints = [v * rinv % cv.order for v in to_int(sc)]
                    want = O.jac_to_bytes(cv, cv.mul(sum(s * (k0 + i * d) for i, s in enumerate(ints)) % cv.order, cv.gen))
                else:
                    want = None
                results = {}
                for layout in ("table", "plain"):
                    g = G.Generators.progression(curve, k0, d, n, table=(layout == "table"))
                    for rounds in ("auto", "0", "1", "3", "5"):
                        if rounds == "auto":
                            os.environ.pop("VDFGPU_MSM_AFFINE", None)
                        else:
                            os.environ["VDFGPU_MSM_AFFINE"] = rounds
                        out = torch.zeros(96, dtype=torch.uint8, device="cuda")
                        _lib.check(lib.vdfgpu_msm_dev(g._h, dev.data_ptr(), n, out.data_ptr()))
                        torch.cuda.synchronize()
                        results[(layout, rounds)] = bytes(out.cpu().numpy().tobytes())
                    g.close()
                ref = want if want is not None else results[("plain", "0")]
                bad = [k for k, v in results.items() if v != ref]
                failures += len(bad)
                print(json.dumps({"log2n": lg, "curve": curve, "points": pts_kind, "scalars": sc_kind,
                                  "checked_against": "python integers" if want is not None else "plain layout, no affine rounds",
                                  "configs": len(results), "mismatches": [list(b) for b in bad]}), flush=True)
os.environ.pop("VDFGPU_MSM_AFFINE", None)
print(json.dumps({"stress_msm": "ok" if failures == 0 else "FAILED", "mismatching_configs": failures}))
sys.exit(1 if failures else 0)
