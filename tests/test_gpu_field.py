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
