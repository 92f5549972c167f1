#!/usr/bin/env python
"""Dependent-multiplication latency of ONE warp (the latency path's unit cost): inline vs out-of-line multiplier."""
import json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vdf_b200 import _lib
lib = _lib.load(); _lib.check(lib.vdfgpu_init(0))
n = 32
a = bytearray(os.urandom(32 * n))
for i in range(n): a[32 * i + 31] &= 0x3F
b = bytes(a[32:] + a[:32]); out = bytearray(32 * n)
res = {}
for name, flag in (("inline", 0), ("call", 1 << 31)):
    ts = {}
    for iters in (1000, 41000):
        _lib.check(lib.vdfgpu_field_mul_batch(1, _lib.as_ptr(a), _lib.as_ptr(b), n, iters | flag, _lib.as_ptr(out)))
        t0 = time.perf_counter()
        _lib.check(lib.vdfgpu_field_mul_batch(1, _lib.as_ptr(a), _lib.as_ptr(b), n, iters | flag, _lib.as_ptr(out)))
        ts[iters] = time.perf_counter() - t0
    per = (ts[41000] - ts[1000]) / 40000
    res[name] = {"us_per_mul": per * 1e6, "cycles": per * 1.965e9}
print(json.dumps(res))
