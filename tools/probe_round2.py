"""Round-2 A/B probes on the GPU (not a test): commitments in flight 1..4 (device-resident and host scalars), and the
Nova fold step with / without CUDA-graph replay of the small MSMs.  Prints one JSON line per measurement."""
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import bench as B  # noqa: E402
from vdf_b200 import _lib, encoding as E, msm as G, nova as N, synthetic as S  # noqa: E402

lib = _lib.load()
_lib.check(lib.vdfgpu_init(0))
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
_lib.check(lib.vdfgpu_set_stream(stream.cuda_stream))


def inflight_probe(log2n=22, steps=12):
    n = 1 << log2n
    r = B.MsmRunner(torch, lib, _lib, n)
    r.streams = [torch.cuda.Stream() for _ in range(4)]
    r.outs = [torch.zeros(96, dtype=torch.uint8, device="cuda") for _ in range(4)]
    for depth in (1, 2, 3, 4):
        def run(count):
            for k in range(count):
                r.launch(k % depth)
        run(depth * 2)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(r.streams[0])
        for s in r.streams[1:]:
            s.wait_event(e0)
        run(steps)
        for s in r.streams[1:]:
            ev = torch.cuda.Event()
            ev.record(s)
            r.streams[0].wait_event(ev)
        e1.record(r.streams[0])
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        print(json.dumps({"probe": "inflight_device", "log2n": log2n, "depth": depth, "ms": ms, "gpts": n / ms / 1e6}), flush=True)
    _lib.check(lib.vdfgpu_set_stream(stream.cuda_stream))
    host = torch.empty((n, 4), dtype=torch.int64, pin_memory=True)
    host.copy_(r.scal)
    outs = [torch.zeros(96, dtype=torch.uint8, pin_memory=True) for _ in range(4)]
    for depth in (1, 2, 3, 4):
        def pipelined(k_steps):
            for k in range(k_steps):
                if k >= depth:
                    _lib.check(lib.vdfgpu_msm_wait((k - depth) % depth))
                _lib.check(lib.vdfgpu_msm_submit(r.gens._h, host.data_ptr(), n, outs[k % depth].data_ptr(), k % depth))
            for k in range(max(0, k_steps - depth), k_steps):
                _lib.check(lib.vdfgpu_msm_wait(k % depth))
        pipelined(depth * 2)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pipelined(steps)
        dt = (time.perf_counter() - t0) / steps
        print(json.dumps({"probe": "inflight_e2e_pinned", "log2n": log2n, "depth": depth, "ms": dt * 1e3, "gpts": n / dt / 1e9}), flush=True)
    r.close()
    _lib.check(lib.vdfgpu_trim())


def nova_probe(ts=(1024, 16384)):
    for graph in ("0", "1"):
        os.environ["VDFGPU_GRAPH"] = graph
        res = B.nova_step_measurements(_lib, ts=ts)
        for t in ts:
            e = res[str(t)]
            print(json.dumps({"probe": "nova_step", "graph": graph, "t": t, "ms": e["ms"], "raw_ms": e["raw_jacobian_ms"],
                              "bank_ms": e["bank_ms"]}), flush=True)
    os.environ["VDFGPU_GRAPH"] = "1"


def small_msm_probe(sizes=(13904, 75344)):
    import numpy as np
    for graph in ("0", "1"):
        os.environ["VDFGPU_GRAPH"] = graph
        for n in sizes:
            g = G.Generators.progression(0, B.K0, B.D, n, table=True)
            rs = np.random.RandomState(1)
            raw = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
            raw[:, 7] &= 0x3FFFFFFF
            out = np.zeros(96, dtype=np.uint8)
            for _ in range(5):
                _lib.check(lib.vdfgpu_msm(g._h, raw.ctypes.data, n, out.ctypes.data))
            t0 = time.perf_counter()
            for _ in range(30):
                _lib.check(lib.vdfgpu_msm(g._h, raw.ctypes.data, n, out.ctypes.data))
            dt = (time.perf_counter() - t0) / 30
            print(json.dumps({"probe": "small_msm_sync_host", "graph": graph, "n": n, "ms": dt * 1e3, "out": out[:8].tobytes().hex()}), flush=True)
            g.close()
    os.environ["VDFGPU_GRAPH"] = "1"


if __name__ == "__main__":
    which = sys.argv[1:] or ["inflight", "nova", "small"]
    if "small" in which:
        small_msm_probe()
    if "nova" in which:
        nova_probe()
    if "inflight" in which:
        inflight_probe()
