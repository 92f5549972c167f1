"""Byte layouts at the C ABI (pasta_curves with feature repr-c, reference Cargo.toml:17).

Field elements are 32 bytes: four little-endian u64 limbs of value * 2**256 mod m (Montgomery form).
These helpers convert Python integers for the host-side mirror; they do no heavy arithmetic.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

# Pallas base field / Vesta scalar field, and Pallas scalar field / Vesta base field
P = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001
Q = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001
R = 1 << 256

FP, FQ = 0, 1
PALLAS, VESTA = 0, 1
MODULUS = {FP: P, FQ: Q}
CURVE_BASE = {PALLAS: P, VESTA: Q}       # coordinate field modulus
CURVE_ORDER = {PALLAS: Q, VESTA: P}      # scalar field modulus == group order
CURVE_SCALAR_FIELD = {PALLAS: FQ, VESTA: FP}
CURVE_BASE_FIELD = {PALLAS: FP, VESTA: FQ}

_RINV = {m: pow(R, -1, m) for m in (P, Q)}
_RMOD = {m: R % m for m in (P, Q)}

AFFINE_BYTES = 72
POINT_BYTES = 96
STATE_BYTES = 96


def fe_to_bytes(v: int, m: int) -> bytes:
    return ((v % m) * _RMOD[m] % m).to_bytes(32, "little")


def fe_from_bytes(b: bytes, m: int) -> int:
    return int.from_bytes(b[:32], "little") * _RINV[m] % m


def fes_to_bytes(vs: Iterable[int], m: int) -> bytes:
    rm = _RMOD[m]
    return b"".join(((v % m) * rm % m).to_bytes(32, "little") for v in vs)


def fes_from_bytes(b: bytes, m: int) -> List[int]:
    ri = _RINV[m]
    return [int.from_bytes(b[k:k + 32], "little") * ri % m for k in range(0, len(b), 32)]


Affine = Optional[Tuple[int, int]]


def affine_to_bytes(pt: Affine, base: int) -> bytes:
    if pt is None:
        return bytes(64) + b"\x01" + bytes(7)
    return fe_to_bytes(pt[0], base) + fe_to_bytes(pt[1], base) + bytes(8)


def affines_to_bytes(pts: Sequence[Affine], base: int) -> bytes:
    return b"".join(affine_to_bytes(p, base) for p in pts)


def affine_from_bytes(b: bytes, base: int) -> Affine:
    if b[64] != 0:
        return None
    return (fe_from_bytes(b[0:32], base), fe_from_bytes(b[32:64], base))


def point_from_bytes(b: bytes, base: int) -> Affine:
    """96-byte Jacobian (X, Y, Z) -> affine tuple or None (identity)."""
    X, Y, Z = (fe_from_bytes(b[k:k + 32], base) for k in (0, 32, 64))
    if Z == 0:
        return None
    zi = pow(Z, -1, base)
    return (X * zi * zi % base, Y * zi * zi * zi % base)


def known_dlog_scalar(raw, k0: int, d: int, first: int = 0) -> int:
    """sum_i s_i * (k0 + (first + i) d) over the 256-bit little-endian integers s_i = raw[i, 0..7] (u32 limbs),
    in O(n) numpy work: 2^15-element chunks keep every partial sum below 2^63."""
    import numpy as np
    n = raw.shape[0]
    CH = 1 << 15
    pad = (-n) % CH
    j = np.arange(CH, dtype=np.uint64)
    s_tot = s_idx = 0
    for limb in range(8):
        col = raw[:, limb].astype(np.uint64)
        if pad:
            col = np.concatenate([col, np.zeros(pad, dtype=np.uint64)])
        A = col.reshape(-1, CH)
        csum = A.sum(axis=1)                      # < 2^47 each
        cjsum = (A * j).sum(axis=1)               # < 2^62 each
        tot = idx = 0
        for c in range(A.shape[0]):
            cs = int(csum[c])
            tot += cs
            idx += (first + c * CH) * cs + int(cjsum[c])
        s_tot += tot << (32 * limb)
        s_idx += idx << (32 * limb)
    return k0 * s_tot + d * s_idx
