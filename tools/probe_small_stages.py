#!/usr/bin/env python
"""Per-stage device time (CUDA events inside the library, graphs off) of Nova-size commitments."""
import ctypes, json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import bench as B
from vdf_b200 import _lib, msm as G
lib = _lib.load(); _lib.check(lib.vdfgpu_init(0))
NAMES = ["digits", "scan", "scatter", "accumulate", "records", "reduce", "final"]
for n in (13904, 75344, 1 << 18, 1 << 20):
    g = G.Generators.progression(0, B.K0, B.D, n, table=True)
    rs = np.random.RandomState(1)
    raw = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    raw[:, 7] &= 0x3FFFFFFF
    out = np.zeros(96, dtype=np.uint8)
    for _ in range(5):
        _lib.check(lib.vdfgpu_msm(g._h, raw.ctypes.data, n, out.ctypes.data))
    _lib.check(lib.vdfgpu_profile_enable(1))
    acc = np.zeros(7)
    buf = (ctypes.c_double * 7)()
    for _ in range(10):
        _lib.check(lib.vdfgpu_msm(g._h, raw.ctypes.data, n, out.ctypes.data))
        _lib.check(lib.vdfgpu_profile_read(buf, 7))
        acc += np.array(list(buf))
    _lib.check(lib.vdfgpu_profile_enable(0))
    print(json.dumps({"n": n, "stage_us": {k: round(v * 100, 1) for k, v in zip(NAMES, acc)}, "total_us": round(acc.sum() * 100, 1)}), flush=True)
    g.close()
