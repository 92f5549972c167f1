#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: Pallas MSM Gpoints/s (BASELINE.json metric), with the
fold-step and batched-verify rates as extra keys on the same JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--log2n L]

One "step" = one commitment MSM over n = 2^L synthetic scalars against a device-resident generator set
(P_i = (k0 + i d) G, generated on the device; generators are fixed for the life of PublicParams in the
reference, src/nova/proof.rs:232-237).  N > 1 (torchrun, one rank per GPU): each rank owns a contiguous point
range of n points (weak scaling), emits one partial point, and the 96-byte partials are all-gathered (NCCL)
and summed on every rank -- SURVEY.md section 8(e).

`value`  : device-timed (CUDA events), scalars already in HBM.
`e2e`    : the same step through the reference-facing C ABI call vdfgpu_msm() with pinned HOST scalars
           (H2D of n*32 B and D2H of the 96-byte commitment inside the timed region).
`--impl reference`: the CPU restatement of the reference's pasta-msm Pippenger (oracle/cpu_ref.c, all host
           cores) on a bounded sample of the same workload; the Rust crates cannot be built here.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "Pallas MSM Gpoints/s"
print_json = print
K0, D = 0x1234567, 0x89ABCDEF01
MUL32_PER_FIELD_MUL = 136        # generic CIOS, n = 8 limbs: 2n^2 + n
FIELD_MUL_PER_MADD = 10          # XYZZ mixed addition 8M + 2S


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2n", type=int, default=22, help="points per GPU = 2^log2n")
    ap.add_argument("--plain", action="store_true", help="plain generator layout instead of the table")
    ap.add_argument("--no-extra", action="store_true", help="skip the fold-step / verify side measurements")
    ap.add_argument("--cpu-log2n", type=int, default=18, help="sample size of the CPU baseline MSM")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def visible_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except (ValueError, IndexError):
            return local_rank
    return local_rank


# ---------------------------------------------------------------------------------------------------
def cpu_msm_baseline(points72: bytes, scalars: bytes, min_seconds: float = 8.0, max_reps: int = 5):
    """oracle/cpu_ref.c MSM (restatement of pasta-msm's CPU Pippenger) on all host cores."""
    from oracle import cpu_ref as C
    cores = C.ncores()
    n = len(scalars) // 32
    C.msm(0, points72[:72 * 1024], scalars[:32 * 1024], True, cores)  # warm-up
    best, total, reps, out = None, 0.0, 0, None
    while reps < max_reps and (total < min_seconds or reps < 2):
        t0 = time.perf_counter()
        out = C.msm(0, points72, scalars, True, cores)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
        total += dt
        reps += 1
    return {"value": n / best / 1e9, "unit": "Gpoints/s", "cores": cores, "kind": "port",
            "sample": f"Pallas MSM n=2^{n.bit_length() - 1}, best of {reps} runs, {best * 1e3:.1f} ms each "
                      f"(oracle/cpu_ref.c: C restatement of pasta-msm's CPU Pippenger; the Rust crates cannot be built here)"}, out


def run_reference(args):
    """CPU arm: the reference algorithm on the host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_ref as C
    n = 1 << args.cpu_log2n
    cores = C.ncores()
    pts = C.progression(0, K0, D, n)
    scal = bytearray(os.urandom(32 * n))
    for i in range(n):
        scal[32 * i + 31] &= 0x3F
    scal = bytes(scal)
    for _ in range(max(1, args.warmup)):
        C.msm(0, pts, scal, True, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        C.msm(0, pts, scal, True, cores)
    dt = (time.perf_counter() - t0) / args.steps
    v = n / dt / 1e9
    sample = (f"each step = one Pallas MSM over a bounded sample of n=2^{args.cpu_log2n} points of the workload "
              f"(oracle/cpu_ref.c, C restatement of pasta-msm's CPU Pippenger, {cores} threads)")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "Gpoints/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (255-bit Montgomery)", "data": "synthetic",
            "config": {"workload": f"Pallas MSM, 2^{args.log2n} points per GPU (CPU arm runs a 2^{args.cpu_log2n} sample per step)"},
            "cpu_baseline": {"value": v, "unit": "Gpoints/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "Gpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print_json(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
def hbm_view(n, stage_s, traffic):
    """The same stage against the HBM roofline (the contract's other bound): compulsory bytes are one 32-byte scalar
    and one 64-byte point per MSM point, so this stage sits orders of magnitude under the HBM limit by algorithmic
    bytes; `traffic_gbs` is what ncu saw it really move."""
    mp = ROOT / "MEASURED_PEAKS.json"
    peak, src = 7700.0, "fallback (B200_PROFILING.md nominal)"
    if mp.exists():
        peak, src = float(json.loads(mp.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    alg = float(n) * 96
    out = {"bound": "hbm", "algorithmic_bytes": alg, "achieved": alg / stage_s / 1e9, "peak": peak, "unit": "GB/s",
           "frac": alg / stage_s / 1e9 / peak, "peak_source": src}
    if traffic:
        out["traffic_gbs"] = traffic / stage_s / 1e9
        out["traffic_frac_of_peak"] = traffic / stage_s / 1e9 / peak
    return out


def r1cs_hbm_measurements(lib, _lib, torch, log2rows: int = 21):
    """HBM-roofline leg for the R1CS kernels: cross-term, multiply_vec and fold on an ENLARGED
    synthetic shape (2^21 constraints, 4 non-zeros per constraint row triple: the step circuit's density) so
    the data (about 0.5 GB) do not sit in L2.  Device-resident operands, CUDA events."""
    import numpy as np
    from vdf_b200.encoding import Q, fe_to_bytes
    cons = vars_ = 1 << log2rows
    io = 2
    r = np.arange(cons, dtype=np.uint64)
    one, neg1, rnd = fe_to_bytes(1, Q), fe_to_bytes(Q - 1, Q), fe_to_bytes(0x1234567890ABCDEF1234567890ABCDEF, Q)
    a_rows, a_cols = r, r
    b_rows, b_cols = r, (r * 7 + 3) % vars_
    c_rows = np.concatenate([r, r])
    c_cols = np.concatenate([(r + 1) % vars_, (r * 5) % vars_])
    a_vals = np.frombuffer(one * cons, dtype=np.uint8)
    b_vals = np.frombuffer(rnd * cons, dtype=np.uint8)
    c_vals = np.frombuffer(one * cons + neg1 * cons, dtype=np.uint8)
    h = ctypes.c_void_p()
    _lib.check(lib.vdfgpu_r1cs_create(1, cons, vars_, io,
                                      a_rows.ctypes.data, a_cols.ctypes.data, a_vals.ctypes.data, cons,
                                      b_rows.ctypes.data, np.ascontiguousarray(b_cols).ctypes.data, b_vals.ctypes.data, cons,
                                      c_rows.ctypes.data, np.ascontiguousarray(c_cols).ctypes.data, c_vals.ctypes.data, 2 * cons,
                                      ctypes.byref(h)))
    nnz = 4 * cons

    def rand_fe(n):
        t = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda")
        t[:, 3] &= (1 << 62) - 1
        return t

    W1, W2, E1 = rand_fe(vars_), rand_fe(vars_), rand_fe(cons)
    uX1, uX2, rr = rand_fe(1 + io), rand_fe(1 + io), rand_fe(1)
    T = torch.zeros((cons, 4), dtype=torch.int64, device="cuda")
    ABC = torch.zeros((3 * cons, 4), dtype=torch.int64, device="cuda")

    def timed(fn, reps=5):
        fn(); fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e-3

    peak, src = 6650.0, "fallback"
    mp = ROOT / "MEASURED_PEAKS.json"
    if mp.exists():
        peak, src = float(json.loads(mp.read_text())["hbm_gbs"]), "measured"
    res = {"shape": {"cons": cons, "vars": vars_, "nnz": nnz}, "hbm_peak_gbs": peak, "hbm_peak_source": src}
    t_ct = timed(lambda: _lib.check(lib.vdfgpu_cross_term_dev(h, W1.data_ptr(), uX1.data_ptr(), W2.data_ptr(), uX2.data_ptr(), T.data_ptr())))
    b_ct = 36 * nnz + 4 * (3 * cons + 1) + 2 * 32 * (vars_ + 1 + io) + 32 * cons
    t_mv = timed(lambda: _lib.check(lib.vdfgpu_multiply_vec_dev(h, W1.data_ptr(), uX1.data_ptr(), ABC.data_ptr())))
    b_mv = 36 * nnz + 4 * (3 * cons + 1) + 32 * (vars_ + 1 + io) + 3 * 32 * cons
    t_fd = timed(lambda: _lib.check(lib.vdfgpu_fold_dev(1, W1.data_ptr(), W2.data_ptr(), vars_, E1.data_ptr(), T.data_ptr(), cons, rr.data_ptr())))
    b_fd = 96 * (vars_ + cons)
    for name, t, b in (("cross_term", t_ct, b_ct), ("multiply_vec", t_mv, b_mv), ("fold", t_fd, b_fd)):
        res[name] = {"ms": t * 1e3, "algorithmic_bytes": b, "achieved_gbs": b / t / 1e9, "frac_of_hbm": b / t / 1e9 / peak}
    _lib.check(lib.vdfgpu_r1cs_destroy(h))
    return res


def cpu_fold_step(shape, W, X, sec_shape, sec_W, sec_X, gens, sec_gens, _lib):
    """CPU baseline leg: the same step through oracle/cpu_ref.c on all host cores."""
    from oracle import cpu_ref as C, pasta as O
    cores = C.ncores()
    lib = _lib.load()
    parts = []
    for curve, sh, w, x, g in ((0, shape, W, X, gens), (1, sec_shape, sec_W, sec_X, sec_gens)):
        fid, cons, nvars, io, A, B, Cm = sh
        m = O.MODULUS[fid]
        ngen = max(cons, nvars)
        pts = bytearray(72 * ngen)
        _lib.check(lib.vdfgpu_gens_export(g._h, 0, ngen, _lib.as_ptr(pts)))
        coo = O.shape_to_coo_bytes(O.R1CSShape(m, cons, nvars, io, A, B, Cm))
        parts.append((curve, fid, m, cons, nvars, io, O.fes_to_bytes(w, m), O.fes_to_bytes(x, m), bytes(pts), coo))
    best = None
    for _ in range(2):
        t0 = time.perf_counter()
        for curve, fid, m, cons, nvars, io, wb, xb, pts, coo in parts:
            one = O.fe_to_bytes(1, m)
            abc1 = C.multiply_vec(fid, cons, nvars, io, coo, wb, one, xb)
            abc2 = C.multiply_vec(fid, cons, nvars, io, coo, wb, one, xb)
            T = C.cross_term(fid, cons, abc1, abc2, one)
            C.msm(curve, pts[:72 * nvars], wb, True, cores)
            C.msm(curve, pts[:72 * cons], T, True, cores)
            r = O.fe_to_bytes(0x1234567890ABCDEF, m)
            C.fold(fid, wb, wb, r)
            C.fold(fid, T, T, r)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"fold_steps_per_s": 1.0 / best, "ms": best * 1e3, "cores": cores, "kind": "port"}


def nova_step_measurements(_lib, ts=(1024, 4096, 16384)):
    """GPU side of one Nova fold step (BASELINE config 3) = NIFS on the primary curve (Pallas; step circuit of t
    inverse-MinRoot rounds + a SYNTHETIC 9.8k-constraint augmented block) followed by NIFS on the secondary curve
    (Vesta; SYNTHETIC 10.3k-constraint block, the trivial step circuit): per curve the fresh witness goes H2D,
    cross-term, ONE batched MSM for commit(W2) and commit(T), commitments D2H, fold with the challenge.  The four
    commitments of the reference's step are covered; bellperson synthesis and the Poseidon RO (host) are not timed."""
    from vdf_b200 import encoding as E, msm as G, nova as N, synthetic as S   # inputs: package-side generator
    res = {}
    s_cons, s_vars, s_io, sA, sB, sC, sec_W, sec_X = S.step_instance(E.FP, 0, 10300, seed=7)
    sec_gs = N.R1CSShape(E.FP, s_cons, s_vars, s_io, sA, sB, sC)
    sec_gens = G.Generators.progression(1, K0, D, max(s_cons, s_vars), table=True)
    sec = N.RunningProver(sec_gs, sec_gens)
    sec.set_running(sec_W, [0] * s_cons, N.RelaxedR1CSInstance(None, None, list(sec_X), 1))
    sWb, sXb = E.fes_to_bytes(sec_W, E.P), E.fes_to_bytes(sec_X, E.P)
    sec_shape = (E.FP, s_cons, s_vars, s_io, sA, sB, sC)
    r_fixed = 0x1234567890ABCDEF
    for t in ts:
        cons, nvars, io, A, B, C, W, X = S.step_instance(E.FQ, t, 9800, seed=42)
        shape = (E.FQ, cons, nvars, io, A, B, C)
        gs = N.R1CSShape(E.FQ, cons, nvars, io, A, B, C)
        gens = G.Generators.progression(0, K0, D, max(cons, nvars), table=True)
        pri = N.RunningProver(gs, gens)
        pri.set_running(W, [0] * cons, N.RelaxedR1CSInstance(None, None, list(X), 1))
        Wb, Xb = E.fes_to_bytes(W, E.Q), E.fes_to_bytes(X, E.Q)

        def step():
            sec.prove_step_bytes(sWb, sXb, r_fixed)
            pri.prove_step_bytes(Wb, Xb, r_fixed)

        for _ in range(3):
            step()
        reps = 15
        t0 = time.perf_counter()
        for _ in range(reps):
            step()
        dt = (time.perf_counter() - t0) / reps
        res[str(t)] = {"fold_steps_per_s": 1.0 / dt, "ms": dt * 1e3, "primary_cons": cons,
                       "primary_vars": nvars, "secondary_cons": s_cons}
        # CPU restatement of the same step's data-parallel work on the host cores (oracle/cpu_ref.c, "port"):
        # 4 commitments (pasta-msm style Pippenger), 2 x 2 multiply_vec, cross-terms, folds
        try:
            res[str(t)]["cpu_port"] = cpu_fold_step(shape, W, X, sec_shape, sec_W, sec_X, gens, sec_gens, _lib)
        except Exception as e:
            res[str(t)]["cpu_port"] = {"error": repr(e)}
        pri.close(); gens.close(); gs.close()
    sec.close(); sec_gens.close(); sec_gs.close()
    res["note"] = "GPU side only; augmented-circuit blocks are SYNTHETIC; host synthesis and Poseidon RO not timed"
    return res


def extra_measurements(lib, _lib, torch):
    """fold-steps/s (t = 1024, synthetic augmented block) and batched verify (C4)."""
    out = {}
    from vdf_b200 import encoding as E, msm as G, nova as N, synthetic as S
    try:
        t, aug = 1024, 9800
        cons, nvars, io, A, B, C, W, X = S.step_instance(E.FQ, t, aug, seed=42)
        gs = N.R1CSShape(E.FQ, cons, nvars, io, A, B, C)
        ngen = max(cons, nvars)
        gens = G.Generators.progression(0, K0, D, ngen, table=True)
        prover = N.RunningProver(gs, gens)
        Wb, Xb = E.fes_to_bytes(W, E.Q), E.fes_to_bytes(X, E.Q)
        prover.set_running(W, [0] * cons, N.RelaxedR1CSInstance(None, None, list(X), 1))
        r_fixed = 0x1234567890ABCDEF                # the random oracle is host work outside this path: fixed challenge
        for _ in range(3):
            prover.prove_step_bytes(Wb, Xb, r_fixed)
        reps = 20
        t0 = time.perf_counter()
        for _ in range(reps):
            prover.prove_step_bytes(Wb, Xb, r_fixed)   # H2D witness, cross-term, batched MSM(W2, T), D2H, fold
        dt = (time.perf_counter() - t0) / reps
        out["nifs_fold"] = {
            "value": 1.0 / dt, "unit": "NIFS folds/s (one curve: commit(W2) + commit_T + fold, host witness in)",
            "ms": dt * 1e3, "t": t, "cons": cons, "vars": nvars, "nnz": gs.nnz,
            "window_bits": gens.window_bits(ngen),
            "note": "augmented-circuit block is SYNTHETIC (9.8k random constraints); synthesis and the Poseidon RO stay on the host and are not timed"}
        prover.close(); gens.close()
        # same step with commitments returned as un-normalised Jacobian points (what pasta-msm returns; the caller's
        # to_affine() normalises on the host): skips the single-thread inversion at the end of the MSM
        gens = G.Generators.progression(0, K0, D, ngen, table=True, raw_jacobian=True)
        prover = N.RunningProver(gs, gens)
        prover.set_running(W, [0] * cons, N.RelaxedR1CSInstance(None, None, list(X), 1))
        lib_ = _lib.load()
        cW, cT = bytearray(96), bytearray(96)
        rb = E.fe_to_bytes(0x1234567890ABCDEF, E.Q)

        def raw_step():
            _lib.check(lib_.vdfgpu_running_commit(prover._h, _lib.as_ptr(Wb), _lib.as_ptr(Xb), _lib.as_ptr(cW), _lib.as_ptr(cT)))
            _lib.check(lib_.vdfgpu_running_finish(prover._h, _lib.as_ptr(rb)))

        for _ in range(3):
            raw_step()
        t0 = time.perf_counter()
        for _ in range(reps):
            raw_step()
        dt2 = (time.perf_counter() - t0) / reps
        out["nifs_fold"]["raw_jacobian_value"] = 1.0 / dt2
        out["nifs_fold"]["raw_jacobian_ms"] = dt2 * 1e3
        prover.close(); gens.close(); gs.close()
    except Exception as e:  # side measurement: never lose the headline line
        out.setdefault("nifs_fold", {})["error"] = repr(e)
    try:
        out["nova_step"] = nova_step_measurements(_lib)
    except Exception as e:
        out["nova_step"] = {"error": repr(e)}
    try:
        out["r1cs_hbm"] = r1cs_hbm_measurements(lib, _lib, torch)
    except Exception as e:
        out["r1cs_hbm"] = {"error": repr(e)}
    try:
        n, t = 1 << 16, 1000
        res = torch.randint(0, 1 << 62, (n, 12), dtype=torch.int64, device="cuda")
        res[:, 3::4] &= (1 << 61) - 1  # every element < 2^253 < modulus
        orig = torch.zeros_like(res)
        ok = torch.zeros(n, dtype=torch.uint8, device="cuda")
        fn = lib.vdfgpu_minroot_check_batch_dev
        for _ in range(2):
            _lib.check(fn(1, res.data_ptr(), orig.data_ptr(), None, t, n, ok.data_ptr()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        reps = 5
        for _ in range(reps):
            _lib.check(fn(1, res.data_ptr(), orig.data_ptr(), None, t, n, ok.data_ptr()))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        steps = n * t
        out["minroot_verify"] = {"value": steps / (ms * 1e-3), "unit": "MinRoot steps verified/s", "ms": ms,
                                 "chains": n, "t": t, "mul32_per_step_convention": 3 * MUL32_PER_FIELD_MUL}
    except Exception as e:
        out["minroot_verify"] = {"error": repr(e)}
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from vdf_b200 import _lib
    from vdf_b200 import msm as G

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the GPU path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = _lib.load()
    _lib.check(lib.vdfgpu_init(local_rank))
    # a dedicated (non-default) torch stream: torch.cuda.Event timing and the library's kernels share it.
    # (The legacy default stream has handle 0, which vdfgpu_set_stream reads as "use the library stream".)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    _lib.check(lib.vdfgpu_set_stream(stream.cuda_stream))

    n = 1 << args.log2n
    table = not args.plain
    steps, warmup = args.steps, max(3, args.warmup)

    # generators: this rank's contiguous point range of the global progression
    gens = G.Generators.progression(0, K0 + rank * n * D, D, n, table=table)
    c_bits = gens.window_bits(n)
    W = (256 + c_bits - 1) // c_bits

    # synthetic scalars: uniform 254-bit values (valid Montgomery-form field elements), seeded per rank
    gen = torch.Generator(device="cuda")
    gen.manual_seed(42 + rank)
    scal = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=gen)
    scal[:, 3] &= (1 << 62) - 1
    scal_host = torch.empty((n, 4), dtype=torch.int64, pin_memory=True)
    scal_host.copy_(scal)
    out_dev = torch.zeros(96, dtype=torch.uint8, device="cuda")
    gathered = torch.zeros(96 * world, dtype=torch.uint8, device="cuda") if world > 1 else None
    total_dev = torch.zeros(96, dtype=torch.uint8, device="cuda")

    def step_device():
        _lib.check(lib.vdfgpu_msm_dev(gens._h, scal.data_ptr(), n, out_dev.data_ptr()))
        if world > 1:
            dist.all_gather_into_tensor(gathered, out_dev)
            _lib.check(lib.vdfgpu_point_sum_dev(0, gathered.data_ptr(), world, total_dev.data_ptr()))

    out_host = torch.zeros(96, dtype=torch.uint8, pin_memory=True)

    def step_e2e():
        # the call a user of the reference makes: commit(scalars) with HOST buffers
        _lib.check(lib.vdfgpu_msm(gens._h, scal_host.data_ptr(), n, out_host.data_ptr()))
        if world > 1:
            out_dev.copy_(out_host, non_blocking=True)
            dist.all_gather_into_tensor(gathered, out_dev)
            _lib.check(lib.vdfgpu_point_sum_dev(0, gathered.data_ptr(), world, total_dev.data_ptr()))
            out_host.copy_(total_dev)
            torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step_device()
    barrier()
    sampler = ClockSampler(visible_index(local_rank))
    if rank == 0:
        sampler.start()
    launches0 = lib.vdfgpu_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        step_device()
    e1.record()
    barrier()
    launches = lib.vdfgpu_launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / steps
    value = world * n / (ms_per_step * 1e-3) / 1e9

    # end to end through the C ABI with host buffers (wall clock, max over ranks).
    # (1) synchronous calls, one after the other
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_e2e()
    barrier()
    e2e_sync_s = (time.perf_counter() - t0) / steps
    # (2) the asynchronous form of the same call, two commitments in flight: every step still uploads its own
    # n*32 bytes from pinned host memory and reads its own 96-byte result back, but the upload of step k+1 runs
    # under the kernels of step k.  Independent commitments (the microbenchmark's case) allow this; the serial
    # MSMs of one Nova step do not, which is why both numbers are reported.
    outs = [torch.zeros(96, dtype=torch.uint8, pin_memory=True) for _ in range(2)]

    def finish(slot):
        _lib.check(lib.vdfgpu_msm_wait(slot))
        if world > 1:
            # combine this step's partials; enqueued behind the MSM already submitted for the next step and
            # read back asynchronously (the closing barrier + synchronize of the timed region covers it)
            out_dev.copy_(outs[slot], non_blocking=True)
            dist.all_gather_into_tensor(gathered, out_dev)
            _lib.check(lib.vdfgpu_point_sum_dev(0, gathered.data_ptr(), world, total_dev.data_ptr()))
            out_host.copy_(total_dev, non_blocking=True)

    def pipelined(k_steps):
        _lib.check(lib.vdfgpu_msm_submit(gens._h, scal_host.data_ptr(), n, outs[0].data_ptr(), 0))
        for k in range(1, k_steps):
            _lib.check(lib.vdfgpu_msm_submit(gens._h, scal_host.data_ptr(), n, outs[k & 1].data_ptr(), k & 1))
            finish((k - 1) & 1)
        finish((k_steps - 1) & 1)

    pipelined(2)
    barrier()
    t0 = time.perf_counter()
    pipelined(steps)
    barrier()
    e2e_s = (time.perf_counter() - t0) / steps
    assert bytes(outs[0].numpy().tobytes()) == bytes(out_host.numpy().tobytes()) or world > 1
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    ts = torch.tensor([e2e_sync_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
    e2e_sync_s = float(ts.item())
    e2e = {"value": world * n / e2e_s / 1e9, "unit": "Gpoints/s", "ms_per_step": e2e_s * 1e3,
           "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": 96,
           "api": "vdfgpu_msm_submit / vdfgpu_msm_wait (host scalars in, host point out), two commitments in flight",
           "sync_value": world * n / e2e_sync_s / 1e9, "sync_ms_per_step": e2e_sync_s * 1e3,
           "sync_api": "vdfgpu_msm(gens, host scalars, n, host out), one call at a time"}

    # per-stage device time of the dominant kernel (CUDA events inside the library, same stream)
    stage_names = ["digits", "scan", "scatter", "accumulate", "records", "reduce", "final"]
    _lib.check(lib.vdfgpu_profile_enable(1))
    acc = [0.0] * 7
    prof_reps = 3
    for _ in range(prof_reps):
        _lib.check(lib.vdfgpu_msm_dev(gens._h, scal.data_ptr(), n, out_dev.data_ptr()))
        buf = (ctypes.c_double * 7)()
        _lib.check(lib.vdfgpu_profile_read(buf, 7))
        acc = [a + b for a, b in zip(acc, buf)]
    _lib.check(lib.vdfgpu_profile_enable(0))
    stage_ms = {k: v / prof_reps for k, v in zip(stage_names, acc)}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # integer-multiply roofline of the dominant kernel (bucket accumulation)
    pw, pl, pa = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    _lib.check(lib.vdfgpu_imad_peak(ctypes.byref(pw), ctypes.byref(pl), ctypes.byref(pa)))
    alg_mul32 = float(n) * W * FIELD_MUL_PER_MADD * MUL32_PER_FIELD_MUL
    acc_s = stage_ms["accumulate"] * 1e-3
    R = gens.affine_rounds(n)
    traffic = None
    tfile = ROOT / "profiles" / "traffic.json"   # dram bytes of one accumulate launch from the committed ncu capture
    if tfile.exists():
        t = json.loads(tfile.read_text())
        if t.get("log2n") == args.log2n and t.get("window_bits") == c_bits and t.get("layout") == ("table" if table else "plain") \
                and t.get("affine_rounds", 0) == R:
            traffic = t.get("accumulate_stage_dram_bytes", t.get("accumulate_dram_bytes_per_launch"))
    # field multiplications the accumulate stage really executes: R batched-affine halving rounds at 6 per addition
    # (the list shrinks to ~E / 2^R), then the XYZZ mixed additions of what is left at 10
    E = float(n) * W
    left = E / (1 << R)
    executed_mul = (E - left) * 6 + left * FIELD_MUL_PER_MADD
    roofline = {
        "bound": "imad",
        "kernel": (f"accumulate stage = {R} batched-affine halving rounds (AffineFwdFn, BatchInvFn, AffineBwdFn) + "
                   "AccumulateFn (XYZZ ranges)") if R else "AccumulateFn (XYZZ bucket accumulation)",
        "achieved": alg_mul32 / acc_s / 1e12, "peak": pw.value / 1e12, "unit": "Tmul32/s",
        "frac": (alg_mul32 / acc_s) / pw.value, "traffic": traffic,
        "frac_executed": (executed_mul * 88 / acc_s) / pw.value,
        "algorithmic": f"n * W(c) * 10 field-mul * 136 mul32 (SURVEY 8d) with the real c={c_bits}, W={W}",
        "peak_source": "measured in this run: vdfgpu_imad_peak, register-only IMAD.WIDE.U32.X carry chains with loop-variant "
                       "multiplicands (nominal 148 SM x 4 SMSP x 8 lanes x 1.965 GHz = 9.3 T; MEASURED_PEAKS.json has no integer figure)",
        "frac_executed_note": "products really executed: 88 per field multiplication (the SURVEY convention counts 136), 6 "
                              "multiplications per affine addition and 10 per XYZZ addition; frac > 1 means the stage does "
                              "less arithmetic than the convention assumes",
        "affine_rounds": R,
        "hbm_view": hbm_view(n, acc_s, traffic),
        "imad_lo_per_s": pl.value, "iadd3_per_s": pa.value,
        "kernel_ms": stage_ms["accumulate"], "stage_ms": stage_ms,
        "share_of_step": stage_ms["accumulate"] / max(1e-9, sum(stage_ms.values())),
        "judge_convention_c16_achieved": float(n) * 21760 / acc_s / 1e12,
    }

    cpu_baseline = None
    parity = None
    if world == 1:
        # bounded CPU sample: first 2^cpu_log2n points/scalars of the same workload, and a parity check of
        # the GPU prefix commitment against it (the only place bench.py executes oracle/)
        m = min(n, 1 << args.cpu_log2n)
        pts_bytes = bytearray(72 * m)
        _lib.check(lib.vdfgpu_gens_export(gens._h, 0, m, _lib.as_ptr(pts_bytes)))
        sc_bytes = scal_host[:m].numpy().tobytes()
        cpu_baseline, cpu_out = cpu_msm_baseline(bytes(pts_bytes), sc_bytes)
        gpu_out = gens.commit_bytes(sc_bytes)
        parity = "ok" if gpu_out == cpu_out else "MISMATCH"

    line = {
        "metric": METRIC, "value": value, "unit": "Gpoints/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 limbs (255-bit Montgomery, exact)", "data": "synthetic",
        "config": {"workload": f"Pallas MSM (BASELINE config 2/5), n = 2^{args.log2n} points per GPU, "
                               f"{'table' if table else 'plain'} generator layout, c = {c_bits}, W = {W}",
                   "points_total": world * n, "scalars": "uniform 254-bit, seed 42+rank",
                   "points": "known-dlog progression (k0 + i d) G generated on the device",
                   "l2": "inputs larger than L2 (scalars 32 B/pt + point table 64 B/pt/level >> 126 MB); no flush needed",
                   "parallelism": f"point-range shards x{world}, all-gather of 96-byte partials" if world > 1 else "single GPU"},
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        "cpu_baseline": cpu_baseline, "parity_vs_cpu_sample": parity,
    }
    if not args.no_extra and world == 1:
        line["extra"] = extra_measurements(lib, _lib, torch)
    print_json(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    # Exactly ONE line on stdout: libraries (NCCL's version banner, torchrun notices) write to fd 1 too, so
    # route fd 1 to stderr for the duration of the run and emit the JSON line on the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    captured = []

    def emit(line):
        captured.append(line)

    global print_json
    print_json = emit
    try:
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        for line in captured:
            os.write(1, (line + "\n").encode())
        os.close(real_stdout)


if __name__ == "__main__":
    main()
