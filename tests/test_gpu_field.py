"""GPU parity: the PTX Montgomery multiplication (field.cuh) against Python integers, through the C ABI
(vdfgpu_field_mul_batch).  Bit-exact; both fields; edge values + seeded random; iterated products."""
import pytest

from oracle import pasta as O
from vdf_b200 import _lib
from tests.util import edge_field_values, rand_scalars

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fid", [O.FIELD_FP, O.FIELD_FQ])
def test_field_mul_matches_python(gpu_lib, fid):
    m = O.MODULUS[fid]
    rng = O.XorShiftRng()
    edge = edge_field_values(m)
    A = [a for a in edge for _ in edge] + rand_scalars(rng, m, 4096)
    B = [b for _ in edge for b in edge] + rand_scalars(rng, m, 4096)
    n = len(A)
    out = bytearray(n * 32)
    _lib.check(gpu_lib.vdfgpu_field_mul_batch(fid, _lib.as_ptr(O.fes_to_bytes(A, m)), _lib.as_ptr(O.fes_to_bytes(B, m)),
                                              n, 1, _lib.as_ptr(out)))
    got = O.fes_from_bytes(bytes(out), m)
    assert got == [a * b % m for a, b in zip(A, B)]
    # raw limbs must be canonical (< m)
    assert all(int.from_bytes(out[k:k + 32], "little") < m for k in range(0, len(out), 32))


@pytest.mark.parametrize("fid", [O.FIELD_FP, O.FIELD_FQ])
def test_field_mul_iterated(gpu_lib, fid):
    m = O.MODULUS[fid]
    rng = O.XorShiftRng()
    n, iters = 512, 37
    A, B = rand_scalars(rng, m, n), rand_scalars(rng, m, n)
    out = bytearray(n * 32)
    _lib.check(gpu_lib.vdfgpu_field_mul_batch(fid, _lib.as_ptr(O.fes_to_bytes(A, m)), _lib.as_ptr(O.fes_to_bytes(B, m)),
                                              n, iters, _lib.as_ptr(out)))
    assert O.fes_from_bytes(bytes(out), m) == [a * pow(b, iters, m) % m for a, b in zip(A, B)]


def test_imad_probe_runs(gpu_lib):
    import ctypes
    w, l, a = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    _lib.check(gpu_lib.vdfgpu_imad_peak(ctypes.byref(w), ctypes.byref(l), ctypes.byref(a)))
    assert w.value > 1e12 and l.value > 1e12 and a.value > 1e12


@pytest.mark.parametrize("fid", [O.FIELD_FP, O.FIELD_FQ])
def test_field_sqr_matches_python(gpu_lib, fid):
    """Dedicated squaring (36 products, field.cuh) against Python integers: chosen RAW limb patterns (the
    Montgomery representatives themselves: all-ones limbs, m - 1 with limbs cleared, powers of two) + random,
    single and iterated."""
    import random
    m = O.MODULUS[fid]
    py = random.Random(5 + fid)
    rinv = pow(1 << 256, -1, m)
    mask = 0xFFFFFFFF
    raw = [0, 1, 2, m - 1, m - 2, m >> 1, (1 << 254) - 1, 1 << 254, (1 << 254) + 1, mask, (1 << 254) | mask]
    raw += [(m - 1) & ~(mask << (32 * k)) for k in range(8)]
    raw += [(m - 1) & ~(1 << (32 * k)) for k in range(8)]
    raw += [sum((mask if py.random() < .5 else py.randrange(1 << 32)) << (32 * k) for k in range(8)) % m for _ in range(3000)]
    raw += [m - 1 - py.randrange(1 << 64) for _ in range(500)]
    raw += [py.randrange(m) for _ in range(4096)]
    A = [r * rinv % m for r in raw]          # field values whose Montgomery form is exactly `raw`
    enc = O.fes_to_bytes(A, m)
    assert [int.from_bytes(enc[k:k + 32], "little") for k in range(0, 32 * 40, 32)] == raw[:40]
    n = len(A)
    for iters in (1, 2, 9):
        out = bytearray(n * 32)
        _lib.check(gpu_lib.vdfgpu_field_mul_batch(fid, _lib.as_ptr(enc), _lib.as_ptr(enc), n, iters | (1 << 30), _lib.as_ptr(out)))
        assert O.fes_from_bytes(bytes(out), m) == [pow(a, 1 << iters, m) for a in A], iters
        assert all(int.from_bytes(out[k:k + 32], "little") < m for k in range(0, len(out), 32))
