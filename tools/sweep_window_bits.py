#!/usr/bin/env python
"""Window-size sweep of the table layout: device time of one commitment (library stage timers, graphs off) per
(log2 n, c).  usage: sweep_window_bits.py [log2n ...]"""
import ctypes, json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import bench as B
from vdf_b200 import _lib, msm as G
lib = _lib.load(); _lib.check(lib.vdfgpu_init(0))
sizes = [int(a) for a in sys.argv[1:]] or [16, 17, 18, 19, 20]
for lg in sizes:
    n = 1 << lg
    rs = np.random.RandomState(1)
    raw = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    raw[:, 7] &= 0x3FFFFFFF
    out = np.zeros(96, dtype=np.uint8)
    row = {}
    for c in range(lg - 4, lg + 2):
        if c > 20 or c < 4:
            continue
        g = G.Generators.progression(0, B.K0, B.D, n, table=True, window_bits=c)
        for _ in range(3):
            _lib.check(lib.vdfgpu_msm(g._h, raw.ctypes.data, n, out.ctypes.data))
        _lib.check(lib.vdfgpu_profile_enable(1))
        acc = np.zeros(7); buf = (ctypes.c_double * 7)()
        for _ in range(6):
            _lib.check(lib.vdfgpu_msm(g._h, raw.ctypes.data, n, out.ctypes.data))
            _lib.check(lib.vdfgpu_profile_read(buf, 7))
            acc += np.array(list(buf))
        _lib.check(lib.vdfgpu_profile_enable(0))
        row[c] = round(float(acc.sum()) / 6 * 1000, 1)
        g.close()
    auto = G.Generators.progression(0, B.K0, B.D, n, table=True)
    print(json.dumps({"log2n": lg, "auto_c": auto.window_bits(n), "total_us_by_c": row}), flush=True)
    auto.close()
