"""Host-side mirror of the sum-check building blocks of CompressedSNARK::prove (src/nova/proof.rs:360-368 -> nova-snark
0.8 spartan_with_ipa_pc: EqPolynomial::evals, SumcheckProof::{prove_cubic_with_additive_term, prove_quad},
MultilinearPolynomial::{bound_poly_var_top, evaluate}).  The arithmetic runs on the GPU through the C ABI; the
transcript (round polynomial in, challenge out) is the caller's, passed as a Python callable."""
from __future__ import annotations

import ctypes
from typing import Callable, List, Sequence, Tuple

from . import _lib
from .encoding import MODULUS, fe_from_bytes, fe_to_bytes, fes_from_bytes, fes_to_bytes

Challenge = Callable[[int, Tuple[int, ...]], int]


def eq_evals(field_id: int, r: Sequence[int]) -> List[int]:
    m = MODULUS[field_id]
    out = bytearray(32 << len(r))
    _lib.check(_lib.load().vdfgpu_eq_evals(field_id, _lib.as_ptr(fes_to_bytes(r, m) or bytes(32)), len(r), _lib.as_ptr(out)))
    return fes_from_bytes(bytes(out), m)


def poly_evaluate(field_id: int, poly: Sequence[int], r: Sequence[int]) -> int:
    m = MODULUS[field_id]
    if len(poly) != 1 << len(r):
        raise ValueError("poly_evaluate: the table must have 2^len(r) entries")
    out = bytearray(32)
    _lib.check(_lib.load().vdfgpu_poly_evaluate(field_id, _lib.as_ptr(fes_to_bytes(poly, m)), _lib.as_ptr(fes_to_bytes(r, m) or bytes(32)),
                                                len(r), _lib.as_ptr(out)))
    return fe_from_bytes(bytes(out), m)


def _round_trampoline(m: int, challenge: Challenge, log: list):
    def fn(_user, rnd, evals_ptr, n_evals, r_out):
        try:
            raw = ctypes.string_at(evals_ptr, 32 * n_evals)
            evals = tuple(fes_from_bytes(raw, m))
            r = challenge(int(rnd), evals) % m
            log.append((evals, r))
            ctypes.memmove(r_out, fe_to_bytes(r, m), 32)
            return 0
        except Exception:   # surfaces as VDFGPU_ERR_STATE
            return 1
    return _lib.ROUND_FN(fn)


def sumcheck(field_id: int, tables: Sequence[Sequence[int]], challenge: Challenge):
    """Cubic-with-additive-term sum-check over 4 tables (comb = A (B C - D)) or quadratic over 2 (comb = A B).
    Returns (per-round evaluations, challenges, final evaluations)."""
    m = MODULUS[field_id]
    n = len(tables[0])
    ell = n.bit_length() - 1
    if n != 1 << ell or any(len(t) != n for t in tables) or len(tables) not in (2, 4):
        raise ValueError("sumcheck: 2 or 4 tables of the same power-of-two length")
    lib = _lib.load()
    bufs = [fes_to_bytes(t, m) for t in tables]
    log: list = []
    cb = _round_trampoline(m, challenge, log)
    final = bytearray(32 * len(tables))
    cbp = ctypes.cast(cb, ctypes.c_void_p)
    if len(tables) == 4:
        rc = lib.vdfgpu_sumcheck_cubic(field_id, *[_lib.as_ptr(b) for b in bufs], ell, cbp, None, _lib.as_ptr(final))
    else:
        rc = lib.vdfgpu_sumcheck_quad(field_id, *[_lib.as_ptr(b) for b in bufs], ell, cbp, None, _lib.as_ptr(final))
    _lib.check(rc)
    return [e for e, _ in log], [r for _, r in log], fes_from_bytes(bytes(final), m)


# ---- inner-product-argument building blocks -----------------------------------------------------------------
def vec_lincomb(field_id: int, a: Sequence[int], b: Sequence[int], x: int, y: int) -> List[int]:
    """out[i] = x a[i] + y b[i] (the vector folds of one IPA round)."""
    m = MODULUS[field_id]
    if len(a) != len(b):
        raise ValueError("vec_lincomb: length mismatch")
    out = bytearray(32 * len(a))
    _lib.check(_lib.load().vdfgpu_vec_lincomb(field_id, _lib.as_ptr(fes_to_bytes(a, m)), _lib.as_ptr(fes_to_bytes(b, m)), len(a),
                                              _lib.as_ptr(fe_to_bytes(x, m)), _lib.as_ptr(fe_to_bytes(y, m)), _lib.as_ptr(out)))
    return fes_from_bytes(bytes(out), m)


def inner_product(field_id: int, a: Sequence[int], b: Sequence[int]) -> int:
    m = MODULUS[field_id]
    if len(a) != len(b):
        raise ValueError("inner_product: length mismatch")
    out = bytearray(32)
    _lib.check(_lib.load().vdfgpu_inner_product(field_id, _lib.as_ptr(fes_to_bytes(a, m) or bytes(32)),
                                                _lib.as_ptr(fes_to_bytes(b, m) or bytes(32)), len(a), _lib.as_ptr(out)))
    return fe_from_bytes(bytes(out), m)


def points_lincomb(curve: int, P: bytes, Q: bytes, w1: int, w2: int) -> bytes:
    """out[i] = w1 P[i] + w2 Q[i] on 72-byte affine points (CommitGens::fold of one IPA round)."""
    from .encoding import CURVE_ORDER
    if len(P) != len(Q) or len(P) % 72:
        raise ValueError("points_lincomb: two affine arrays of the same length")
    order = CURVE_ORDER[curve]
    out = bytearray(len(P))
    _lib.check(_lib.load().vdfgpu_points_lincomb(curve, _lib.as_ptr(P), _lib.as_ptr(Q), len(P) // 72,
                                                 _lib.as_ptr(fe_to_bytes(w1, order)), _lib.as_ptr(fe_to_bytes(w2, order)), _lib.as_ptr(out)))
    return bytes(out)

