"""Commitment generators and MSM: host-side mirror of nova-snark's CommitGens / commit() and of
pasta_msm::{pallas,vesta}, which the reference reaches through RecursiveSNARK::prove_step
(src/nova/proof.rs:342-349).  All group arithmetic runs on the GPU through the C ABI."""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

from . import _lib
from .encoding import (AFFINE_BYTES, CURVE_BASE, CURVE_ORDER, POINT_BYTES, Affine, affine_from_bytes,
                       affines_to_bytes, fes_to_bytes, point_from_bytes)

GENS_TABLE = 1
GENS_RAW_JACOBIAN = 2   # results un-normalised (X, Y, Z) as pasta-msm returns them; decode with point_from_bytes


class Generators:
    """Device-resident generator set (nova CommitGens).  `table=True` precomputes the window levels."""

    def __init__(self, curve: int, handle: int):
        self.curve = curve
        self._h = ctypes.c_void_p(handle)

    @classmethod
    def from_affine_bytes(cls, curve: int, data: bytes, table: bool = False, window_bits: int = 0,
                          raw_jacobian: bool = False) -> "Generators":
        if len(data) % AFFINE_BYTES:
            raise ValueError("affine point array must be a multiple of 72 bytes")
        lib = _lib.load()
        h = ctypes.c_void_p()
        flags = (GENS_TABLE if table else 0) | (GENS_RAW_JACOBIAN if raw_jacobian else 0)
        _lib.check(lib.vdfgpu_gens_create(curve, _lib.as_ptr(data), len(data) // AFFINE_BYTES,
                                          flags, window_bits, ctypes.byref(h)))
        return cls(curve, h.value)

    @classmethod
    def from_points(cls, curve: int, pts: Sequence[Affine], **kw) -> "Generators":
        return cls.from_affine_bytes(curve, affines_to_bytes(pts, CURVE_BASE[curve]), **kw)

    @classmethod
    def progression(cls, curve: int, k0: int, d: int, n: int, table: bool = False, window_bits: int = 0,
                    raw_jacobian: bool = False) -> "Generators":
        """Synthetic known-discrete-log set P_i = (k0 + i d) G, generated on the device."""
        lib = _lib.load()
        h = ctypes.c_void_p()
        flags = (GENS_TABLE if table else 0) | (GENS_RAW_JACOBIAN if raw_jacobian else 0)
        _lib.check(lib.vdfgpu_gens_progression(curve, _lib.as_ptr(k0.to_bytes(32, "little")),
                                               _lib.as_ptr(d.to_bytes(32, "little")), n,
                                               flags, window_bits, ctypes.byref(h)))
        return cls(curve, h.value)

    def __len__(self) -> int:
        return _lib.load().vdfgpu_gens_len(self._h)

    def window_bits(self, n: int) -> int:
        return _lib.load().vdfgpu_gens_window_bits(self._h, n)

    def affine_rounds(self, n: int) -> int:
        return _lib.load().vdfgpu_gens_affine_rounds(self._h, n)

    def export(self, first: int = 0, count: Optional[int] = None) -> List[Affine]:
        count = len(self) - first if count is None else count
        buf = bytearray(count * AFFINE_BYTES)
        _lib.check(_lib.load().vdfgpu_gens_export(self._h, first, count, _lib.as_ptr(buf)))
        base = CURVE_BASE[self.curve]
        return [affine_from_bytes(bytes(buf[k:k + AFFINE_BYTES]), base) for k in range(0, len(buf), AFFINE_BYTES)]

    def commit_bytes(self, scalars_mont: bytes) -> bytes:
        """MSM over the first len(scalars) generators; scalars as 32-byte Montgomery elements."""
        out = bytearray(POINT_BYTES)
        _lib.check(_lib.load().vdfgpu_msm(self._h, _lib.as_ptr(scalars_mont), len(scalars_mont) // 32, _lib.as_ptr(out)))
        return bytes(out)

    def commit(self, scalars: Sequence[int]) -> Affine:
        """nova commit(): sum_i scalars[i] * gens[i], returned as an affine tuple (None = identity)."""
        raw = self.commit_bytes(fes_to_bytes(scalars, CURVE_ORDER[self.curve]))
        return point_from_bytes(raw, CURVE_BASE[self.curve])

    def msm_dev(self, scalars_dev_ptr: int, n: int, out_dev_ptr: int, first: int = 0) -> None:
        """Enqueue an MSM on device-resident scalars (no synchronisation)."""
        lib = _lib.load()
        if first:
            _lib.check(lib.vdfgpu_msm_range_dev(self._h, first, scalars_dev_ptr, n, out_dev_ptr))
        else:
            _lib.check(lib.vdfgpu_msm_dev(self._h, scalars_dev_ptr, n, out_dev_ptr))

    def close(self) -> None:
        if self._h:
            _lib.load().vdfgpu_gens_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def mult_pippenger(curve: int, points_affine72: bytes, scalars32: bytes, is_mont: bool = True) -> bytes:
    """pasta-msm's one-shot entry point (points travel with the call)."""
    lib = _lib.load()
    out = bytearray(POINT_BYTES)
    fn = lib.mult_pippenger_pallas if curve == 0 else lib.mult_pippenger_vesta
    fn(_lib.as_ptr(out), _lib.as_ptr(points_affine72), len(points_affine72) // AFFINE_BYTES, _lib.as_ptr(scalars32), is_mont)
    return bytes(out)


def point_sum(curve: int, points96: bytes) -> bytes:
    out = bytearray(POINT_BYTES)
    _lib.check(_lib.load().vdfgpu_point_sum(curve, _lib.as_ptr(points96), len(points96) // POINT_BYTES, _lib.as_ptr(out)))
    return bytes(out)


def sharded_msm(gens: Generators, scalars_mont: bytes, rank: int, world: int, all_gather=None) -> bytes:
    """Point-range-sharded MSM: `gens` holds THIS rank's contiguous point range and
    `scalars_mont` the matching scalars.  Each rank computes one partial point; the 96-byte partials are
    exchanged with `all_gather(bytes) -> list[bytes]` (torch.distributed wrapper) and summed on the GPU."""
    part = gens.commit_bytes(scalars_mont)
    if world == 1 or all_gather is None:
        return part
    parts = all_gather(part)
    return point_sum(gens.curve, b"".join(parts))
