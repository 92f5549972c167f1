#!/usr/bin/env python
"""Per-stage device times of the MSM pipeline over sizes / layouts / tuning knobs (CUDA events inside the
library).  Usage: python tools/sweep_msm.py [--log2n 16 18 20 22] [--plain] [--S 32 64 128] [--c 0]"""
import argparse
import ctypes
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from vdf_b200 import _lib, msm as G  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, nargs="+", default=[16, 18, 20, 22])
ap.add_argument("--n", type=int, nargs="+", default=[], help="explicit point counts (instead of --log2n)")
ap.add_argument("--layouts", nargs="+", default=["table", "plain"])
ap.add_argument("--S", type=int, nargs="+", default=[0])
ap.add_argument("--c", type=int, nargs="+", default=[0])
ap.add_argument("--affine", type=int, nargs="+", default=[-1], help="batched-affine halving rounds (-1 = library default)")
ap.add_argument("--K", type=int, nargs="+", default=[64], help="additions per thread and affine round")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()

lib = _lib.load()
_lib.check(lib.vdfgpu_init(0))
_stream = torch.cuda.Stream()
torch.cuda.set_stream(_stream)
_lib.check(lib.vdfgpu_set_stream(_stream.cuda_stream))
names = ["digits", "scan", "scatter", "accumulate", "records", "reduce", "final"]
for n in (args.n or [1 << lg for lg in args.log2n]):
    lg = n.bit_length() - 1
    scal = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda")
    scal[:, 3] &= (1 << 62) - 1
    out = torch.zeros(96, dtype=torch.uint8, device="cuda")
    for layout in args.layouts:
        for c in args.c:
            g = G.Generators.progression(0, 12345, 678, n, table=(layout == "table"), window_bits=c)
            for S, A, K in [(S, A, K) for S in args.S for A in args.affine for K in (args.K if A else args.K[:1])]:
                if A >= 0:
                    os.environ["VDFGPU_MSM_AFFINE"] = str(A)
                    os.environ["VDFGPU_MSM_AFFINE_K"] = str(K)
                else:
                    os.environ.pop("VDFGPU_MSM_AFFINE", None)
                    os.environ.pop("VDFGPU_MSM_AFFINE_K", None)
                if S:
                    os.environ["VDFGPU_MSM_S"] = str(S)
                else:
                    os.environ.pop("VDFGPU_MSM_S", None)
                for _ in range(2):
                    _lib.check(lib.vdfgpu_msm_dev(g._h, scal.data_ptr(), n, out.data_ptr()))
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.reps):
                    _lib.check(lib.vdfgpu_msm_dev(g._h, scal.data_ptr(), n, out.data_ptr()))
                e1.record()
                torch.cuda.synchronize()
                total = e0.elapsed_time(e1) / args.reps
                _lib.check(lib.vdfgpu_profile_enable(1))
                _lib.check(lib.vdfgpu_msm_dev(g._h, scal.data_ptr(), n, out.data_ptr()))
                buf = (ctypes.c_double * 7)()
                _lib.check(lib.vdfgpu_profile_read(buf, 7))
                _lib.check(lib.vdfgpu_profile_enable(0))
                rec = {"n": n, "log2n": lg, "layout": layout, "c": g.window_bits(n), "S": S, "affine": g.affine_rounds(n) if A < 0 else A, "K": K, "out": bytes(out.cpu().numpy()[:8]).hex(), "ms": round(total, 4),
                       "Gpts/s": round(n / total / 1e6, 4), "stages": {k: round(v, 4) for k, v in zip(names, buf)}}
                print(json.dumps(rec), flush=True)
            g.close()
