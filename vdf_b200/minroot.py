"""MinRoot VDF host interface, mirroring src/minroot.rs: State, MinRootVDF (PallasVDF / VestaVDF),
Evaluation.  The sequential fifth-root chain (`eval`, minroot.rs:348-359) stays on the host by design;
the fast direction -- `check`, `Evaluation::verify`, `Evaluation::append` (minroot.rs:363-371, :424-438)
-- runs batched on the GPU, one independent chain per thread."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

from . import _lib
from .encoding import FP, FQ, MODULUS, STATE_BYTES, fe_from_bytes, fe_to_bytes

# src/minroot.rs:273-285
FP_RESCUE_INVALPHA = 0x33333333333333333333333333333333_4E9EE0C9A10A60E2_E0F0F3F0CCCCCCCD
FQ_RESCUE_INVALPHA = 0x33333333333333333333333333333333_4E9EE0C9A143BA4A_D69F2280CCCCCCCD


@dataclass(frozen=True)
class State:  # minroot.rs:267-272
    x: int
    y: int
    i: int

    def to_bytes(self, m: int) -> bytes:
        return fe_to_bytes(self.x, m) + fe_to_bytes(self.y, m) + fe_to_bytes(self.i, m)

    @staticmethod
    def from_bytes(b: bytes, m: int) -> "State":
        return State(fe_from_bytes(b[0:32], m), fe_from_bytes(b[32:64], m), fe_from_bytes(b[64:96], m))


class MinRootVDF:
    """trait MinRootVDF<G> (minroot.rs:287-374) for one scalar field."""

    field_id: int = FQ
    exponent_value: int = FQ_RESCUE_INVALPHA

    def __init__(self):
        self.m = MODULUS[self.field_id]

    @classmethod
    def inverse_exponent(cls) -> int:  # minroot.rs:68-70
        return 5

    @classmethod
    def exponent(cls) -> int:  # minroot.rs:64-66
        return cls.exponent_value

    def element(self, n: int) -> int:  # minroot.rs:60-62
        return n % self.m

    # -- slow direction: host, sequential (north_star: cannot be parallelised) --------------------
    def forward_step(self, x: int) -> int:  # minroot.rs:312-320
        return pow(x, self.exponent_value, self.m)

    def round(self, s: State) -> State:  # minroot.rs:329-335
        m = self.m
        return State(self.forward_step((s.x + s.y) % m), (s.x + s.i) % m, (s.i + 1) % m)

    def eval(self, x: State, t: int) -> State:  # minroot.rs:348-359
        for _ in range(t):
            x = self.round(x)
        return x

    # -- fast direction: GPU, batched -------------------------------------------------------------
    def inverse_eval_batch(self, results: Sequence[State], t: int) -> List[State]:
        n = len(results)
        if n == 0:
            return []
        inp = b"".join(s.to_bytes(self.m) for s in results)
        out = bytearray(n * STATE_BYTES)
        _lib.check(_lib.load().vdfgpu_minroot_inverse_eval_batch(self.field_id, _lib.as_ptr(inp), t, n, _lib.as_ptr(out)))
        return [State.from_bytes(bytes(out[k:k + STATE_BYTES]), self.m) for k in range(0, len(out), STATE_BYTES)]

    def inverse_eval(self, x: State, t: int) -> State:  # minroot.rs:363-365
        return self.inverse_eval_batch([x], t)[0]

    def step_witness_batch(self, results: Sequence[State], t: int) -> List[List[int]]:
        """The 4t+1 step-circuit variables of each step (src/nova/proof.rs:107-126, :162-189), generated on the
        GPU in allocation order: per round new_x, tmp1, tmp2, new_y; then final_i."""
        n = len(results)
        if n == 0:
            return []
        per = 4 * t + 1
        inp = b"".join(s.to_bytes(self.m) for s in results)
        out = bytearray(n * per * 32)
        _lib.check(_lib.load().vdfgpu_minroot_witness_batch(self.field_id, _lib.as_ptr(inp), t, n, _lib.as_ptr(out)))
        from .encoding import fes_from_bytes
        vals = fes_from_bytes(bytes(out), self.m)
        return [vals[k * per:(k + 1) * per] for k in range(n)]

    def check_batch_bytes(self, results: bytes, originals: bytes, t) -> List[bool]:
        n = len(results) // STATE_BYTES
        if len(originals) != len(results):
            raise ValueError("results and originals differ in length")
        if n == 0:
            return []
        ok = bytearray(n)
        lib = _lib.load()
        if isinstance(t, int):
            _lib.check(lib.vdfgpu_minroot_check_batch(self.field_id, _lib.as_ptr(results), _lib.as_ptr(originals), None, t, n, _lib.as_ptr(ok)))
        else:
            import struct
            if len(t) != n:
                raise ValueError("one t per chain expected")
            tb = struct.pack("<%dQ" % n, *t)
            _lib.check(lib.vdfgpu_minroot_check_batch(self.field_id, _lib.as_ptr(results), _lib.as_ptr(originals), _lib.as_ptr(tb), 0, n, _lib.as_ptr(ok)))
        return [bool(b) for b in ok]

    def check_batch(self, results: Sequence[State], t, originals: Sequence[State]) -> List[bool]:
        """check() for many independent (result, t, original) triples; t is an int or one int per chain."""
        return self.check_batch_bytes(b"".join(s.to_bytes(self.m) for s in results),
                                      b"".join(s.to_bytes(self.m) for s in originals), t)

    def check(self, result: State, t: int, original: State) -> bool:  # minroot.rs:369-371
        return self.check_batch([result], t, [original])[0]


class PallasVDF(MinRootVDF):  # minroot.rs:38-85: modulus of Fq
    field_id = FQ
    exponent_value = FQ_RESCUE_INVALPHA


class VestaVDF(MinRootVDF):  # minroot.rs:199-262: modulus of Fp
    field_id = FP
    exponent_value = FP_RESCUE_INVALPHA


TargetVDF = PallasVDF  # minroot.rs:265


@dataclass
class Evaluation:  # minroot.rs:376-439
    vdf: MinRootVDF
    result: State
    t: int

    @classmethod
    def eval(cls, vdf: MinRootVDF, x: State, t: int) -> Tuple[List[int], "Evaluation"]:  # :394-408
        result = vdf.eval(x, t)
        return [result.x, result.y, result.i], cls(vdf, result, t)

    def verify(self, original: State) -> bool:  # :424-426
        return self.vdf.check(self.result, self.t, original)

    def append(self, other: "Evaluation") -> Optional["Evaluation"]:  # :428-438
        if other.verify(self.result):
            return Evaluation(self.vdf, other.result, self.t + other.t)
        return None

    @staticmethod
    def verify_batch(evals: Sequence["Evaluation"], originals: Sequence[State]) -> List[bool]:
        """Many independent Evaluation::verify calls in one launch (all on the same field)."""
        if not evals:
            return []
        vdf = evals[0].vdf
        return vdf.check_batch([e.result for e in evals], [e.t for e in evals], originals)
