#!/usr/bin/env python
"""Runs the large-shape R1CS kernels (2^21 constraints: multiply_vec, cross-term, fold) and the batched MinRoot check
(2^16 chains x 1000 rounds) once each, with the same inputs bench.py's `extra` section uses -- a short target for
`ncu --set full -k regex:"CrossTermFn|MultiplyVecFn|FoldFn|MinRootCheckFn"` (profiles/r1_side_kernels_ncu.md)."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

import bench  # noqa: E402
from vdf_b200 import _lib  # noqa: E402

lib = _lib.load()
_lib.check(lib.vdfgpu_init(0))
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
_lib.check(lib.vdfgpu_set_stream(st.cuda_stream))
out = {"r1cs_hbm": bench.r1cs_hbm_measurements(lib, _lib, torch)}
n, t = 1 << 16, 1000
res = torch.randint(0, 1 << 62, (n, 12), dtype=torch.int64, device="cuda")
res[:, 3::4] &= (1 << 61) - 1
orig = torch.zeros_like(res)
ok = torch.zeros(n, dtype=torch.uint8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
_lib.check(lib.vdfgpu_minroot_check_batch_dev(1, res.data_ptr(), orig.data_ptr(), None, t, n, ok.data_ptr()))
e0.record()
_lib.check(lib.vdfgpu_minroot_check_batch_dev(1, res.data_ptr(), orig.data_ptr(), None, t, n, ok.data_ptr()))
e1.record()
torch.cuda.synchronize()
out["minroot_check_ms"] = e0.elapsed_time(e1)
print(json.dumps(out))
