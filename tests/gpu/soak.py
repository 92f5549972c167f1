#!/usr/bin/env python
"""Soak test of the library's caches on the GPU (not a pytest: ~1-2 minutes): hundreds of commitments of random
sizes through every host entry point -- explicit handles (table / plain), the literal pasta-msm symbols over more
generator vectors than the drop-in cache holds, asynchronous slots -- interleaved so that workspace arenas grow,
graphs are captured, replayed and evicted (more plans than the 24-entry cache), staging buffers are reused and
resident sets are dropped and rebuilt.  Every result is checked against the known-discrete-log closed form (O(n)
numpy + a one-point MSM on the GPU); device memory must not creep.  Prints one JSON line."""
import json
import random
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from vdf_b200 import _lib, msm as G  # noqa: E402
from vdf_b200.encoding import CURVE_ORDER, known_dlog_scalar  # noqa: E402

lib = _lib.load()
_lib.check(lib.vdfgpu_init(0))
py = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
rs = np.random.RandomState(7)
K0, D = 0x1234567, 0x89ABCDEF01
one_point = {c: G.Generators.progression(c, 1, 0, 1, table=False) for c in (0, 1)}


def expected(curve, raw, k0, first=0):
    order = CURVE_ORDER[curve]
    s = known_dlog_scalar(raw, k0, D, first=first) * pow(1 << 256, -1, order) % order
    return one_point[curve].commit_bytes((s * (1 << 256) % order).to_bytes(32, "little"))


def scalars(n):
    raw = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    raw[:, 7] &= 0x3FFFFFFF
    if py.random() < 0.3:                      # witness-like: half of them bits
        sel = rs.rand(n) < 0.5
        raw[sel, 1:] = 0
        raw[sel, 0] &= 1
    return raw


sets, hosts = [], []
for k in range(6):                              # 6 generator vectors: 2 more than the drop-in cache holds
    curve = k & 1
    n = 1 << py.choice([11, 12, 13, 14, 16])
    g = G.Generators.progression(curve, K0 + k, D, n, table=(k % 3 != 2))
    pts = np.zeros(72 * n, dtype=np.uint8)
    _lib.check(lib.vdfgpu_gens_export(g._h, 0, n, pts.ctypes.data))
    sets.append((curve, n, g, K0 + k))
    hosts.append(pts)
free0 = torch.cuda.mem_get_info()[0]
checked = mism = 0
out = np.zeros(96, dtype=np.uint8)
pinned = [torch.zeros(96, dtype=torch.uint8).pin_memory() for _ in range(4)]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
for it in range(iters):
    k = py.randrange(len(sets))
    curve, n, g, k0 = sets[k]
    m = py.choice([n, n, n - py.randrange(1, 50), py.randrange(1, n), 1024 + py.randrange(n - 1024) if n > 1100 else n])
    raw = scalars(m)
    want = expected(curve, raw, k0)
    mode = py.randrange(4)
    if mode == 0:                                # explicit handle, synchronous
        _lib.check(lib.vdfgpu_msm(g._h, raw.ctypes.data, m, out.ctypes.data))
        got = out.tobytes()
    elif mode == 1:                              # the literal pasta-msm symbol (resident-set cache, LRU of 4)
        fn = lib.mult_pippenger_pallas if curve == 0 else lib.mult_pippenger_vesta
        fn(out.ctypes.data, hosts[k].ctypes.data, m, raw.ctypes.data, True)
        got = out.tobytes()
    elif mode == 2:                              # asynchronous slot
        slot = it & 3
        hp = torch.from_numpy(raw.view(np.int64).reshape(m, 4)).pin_memory()
        _lib.check(lib.vdfgpu_msm_submit(g._h, hp.data_ptr(), m, pinned[slot].data_ptr(), slot))
        _lib.check(lib.vdfgpu_msm_wait(slot))
        got = pinned[slot].numpy().tobytes()
    else:                                        # device pointers on a stream of our own
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            d = torch.from_numpy(raw.view(np.int64).reshape(m, 4)).cuda()
            o = torch.zeros(96, dtype=torch.uint8, device="cuda")
            _lib.check(lib.vdfgpu_set_stream(st.cuda_stream))
            _lib.check(lib.vdfgpu_msm_dev(g._h, d.data_ptr(), m, o.data_ptr()))
            _lib.check(lib.vdfgpu_synchronize())
            _lib.check(lib.vdfgpu_set_stream(None))
        got = o.cpu().numpy().tobytes()
        del d, o, st
    checked += 1
    if got != want:
        mism += 1
        print(json.dumps({"mismatch": it, "mode": mode, "set": k, "m": m}), flush=True)
    if it % 97 == 96:
        _lib.check(lib.vdfgpu_trim())
torch.cuda.synchronize()
_lib.check(lib.vdfgpu_trim())
_lib.check(lib.vdfgpu_dropin_cache_clear())
torch.cuda.empty_cache()
free1 = torch.cuda.mem_get_info()[0]
print(json.dumps({"soak": "ok" if mism == 0 else "FAILED", "commitments": checked, "mismatches": mism,
                  "device_memory_delta_mib": (free0 - free1) / (1 << 20)}))
sys.exit(0 if mism == 0 else 1)
