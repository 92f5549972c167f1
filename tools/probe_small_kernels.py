#!/usr/bin/env python
"""A few Nova-size commitments without graph replay, for an ncu launch list (per-kernel times of the latency path).
usage: probe_small_kernels.py [n] [iters]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import os
os.environ.setdefault("VDFGPU_GRAPH", "0")
import numpy as np
import bench as B
from vdf_b200 import _lib, msm as G
lib = _lib.load(); _lib.check(lib.vdfgpu_init(0))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 13904
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
g = G.Generators.progression(0, B.K0, B.D, n, table=True)
rs = np.random.RandomState(1)
raw = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
raw[:, 7] &= 0x3FFFFFFF
out = np.zeros(96, dtype=np.uint8)
for _ in range(iters):
    _lib.check(lib.vdfgpu_msm(g._h, raw.ctypes.data, n, out.ctypes.data))
print(out[:8].tobytes().hex())
g.close()
