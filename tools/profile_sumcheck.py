#!/usr/bin/env python
"""One cubic and one quadratic sum-check over 2^24-entry tables plus an IPA generator fold of 2^14 outputs: a short
target for `ncu --set full -k regex:"sc_|EqCombineFn|BindTopFn|PointLinCombFn"` (profiles/r2_sumcheck_ncu.md)."""
import ctypes
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from vdf_b200 import _lib, msm as G  # noqa: E402

lib = _lib.load()
_lib.check(lib.vdfgpu_init(0))
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
_lib.check(lib.vdfgpu_set_stream(st.cuda_stream))
ell = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << ell
tabs = [bench.rand_fe_dev(torch, n) for _ in range(4)]
r_host = (ctypes.c_uint64 * 4)(5, 0, 0, 0)


def cb(_u, rnd, evals, k, r_out):
    ctypes.memmove(r_out, r_host, 32)
    return 0


fn = _lib.ROUND_FN(cb)
final = (ctypes.c_uint8 * 128)()
out = {}
for name, k in (("cubic", 4), ("quad", 2)):
    fresh = [t.clone() for t in tabs[:k]]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    if k == 4:
        _lib.check(lib.vdfgpu_sumcheck_cubic_dev(1, *[t.data_ptr() for t in fresh], ell, ctypes.cast(fn, ctypes.c_void_p), None, final))
    else:
        _lib.check(lib.vdfgpu_sumcheck_quad_dev(1, *[t.data_ptr() for t in fresh], ell, ctypes.cast(fn, ctypes.c_void_p), None, final))
    out[name + "_ms"] = (time.perf_counter() - t0) * 1e3
eq = torch.zeros((n, 4), dtype=torch.int64, device="cuda")
rr = np.frombuffer(bytes(32 * ell), dtype=np.uint8).copy()
rr[::32] = 7
_lib.check(lib.vdfgpu_eq_evals_dev(1, rr.ctypes.data, ell, eq.data_ptr()))
m = 1 << 14
g = G.Generators.progression(0, 5, 7, 2 * m, table=False)
pts = np.zeros(72 * 2 * m, dtype=np.uint8)
_lib.check(lib.vdfgpu_gens_export(g._h, 0, 2 * m, pts.ctypes.data))
w = np.full(64, 0x21, dtype=np.uint8)
outp = np.zeros(72 * m, dtype=np.uint8)
_lib.check(lib.vdfgpu_points_lincomb(0, pts.ctypes.data, pts.ctypes.data + 72 * m, m, w.ctypes.data, w.ctypes.data + 32, outp.ctypes.data))
print(json.dumps(out))
