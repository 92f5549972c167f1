#!/usr/bin/env python
"""Field-multiplication throughput probe: iterated Montgomery products at full occupancy
(vdfgpu_field_mul_batch), to separate the multiplier's own ceiling from the MSM kernel's behaviour."""
import ctypes
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vdf_b200 import _lib  # noqa: E402

lib = _lib.load()
_lib.check(lib.vdfgpu_init(0))
n = 148 * 2048 * 2
a = bytearray(os.urandom(32 * n))
for i in range(n):
    a[32 * i + 31] &= 0x3F
b = bytes(a[32:] + a[:32])
out = bytearray(32 * n)
res = {}
for fid in (0, 1):
    times = {}
    for iters in (1, 4001, 8001):
        _lib.check(lib.vdfgpu_field_mul_batch(fid, _lib.as_ptr(a), _lib.as_ptr(b), n, iters, _lib.as_ptr(out)))
        t0 = time.perf_counter()
        _lib.check(lib.vdfgpu_field_mul_batch(fid, _lib.as_ptr(a), _lib.as_ptr(b), n, iters, _lib.as_ptr(out)))
        times[iters] = time.perf_counter() - t0
    dt = times[8001] - times[4001]
    rate = n * 4000 / dt
    res["fp" if fid == 0 else "fq"] = {"field_mul_per_s": rate, "cycles_per_warp_mul_per_smsp": 148 * 4 * 1.965e9 * 32 / rate,
                                       "ms_4000_iters": dt * 1e3}
w, l, ad = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
_lib.check(lib.vdfgpu_imad_peak(ctypes.byref(w), ctypes.byref(l), ctypes.byref(ad)))
res["imad_wide_per_s"], res["imad_lo_per_s"], res["iadd3_per_s"] = w.value, l.value, ad.value
print(json.dumps(res))
