#!/usr/bin/env python
"""Large single-GPU MSMs (the 1-GPU sweep extended to 2^26): the table and plain layouts are two
independent code paths over the same points (different window size, bucket sets, reduction path), so equal
result bytes at sizes the Python oracle cannot reach is a strong parity check; timings are printed beside it."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from vdf_b200 import _lib, msm as G  # noqa: E402

lib = _lib.load()
_lib.check(lib.vdfgpu_init(0))
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
_lib.check(lib.vdfgpu_set_stream(st.cuda_stream))
sizes = [int(a) for a in sys.argv[1:]] or [24, 26]
for lg in sizes:
    n = 1 << lg
    gen = torch.Generator(device="cuda"); gen.manual_seed(lg)
    scal = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=gen)
    scal[:, 3] &= (1 << 62) - 1
    outs, times = {}, {}
    for layout in ("table", "plain"):
        g = G.Generators.progression(0, 0x1234567, 0x89ABCDEF01, n, table=(layout == "table"))
        out = torch.zeros(96, dtype=torch.uint8, device="cuda")
        _lib.check(lib.vdfgpu_msm_dev(g._h, scal.data_ptr(), n, out.data_ptr()))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.vdfgpu_msm_dev(g._h, scal.data_ptr(), n, out.data_ptr()))
        e1.record()
        torch.cuda.synchronize()
        outs[layout] = bytes(out.cpu().numpy().tobytes())
        times[layout] = e0.elapsed_time(e1)
        c = g.window_bits(n)
        g.close()
        torch.cuda.empty_cache()
        print(json.dumps({"log2n": lg, "layout": layout, "c": c, "ms": round(times[layout], 3),
                          "Gpts/s": round(n / times[layout] / 1e6, 4)}), flush=True)
    print(json.dumps({"log2n": lg, "table_equals_plain": outs["table"] == outs["plain"], "nonzero": any(outs["table"])}), flush=True)
