"""CPU: the carry-exact model of the PTX multiplier / squaring (tools/ptx_model.py) against big integers.
Checks the ALGORITHM field.cuh implements (row structure, specialised reduction, which carries are propagated)
without a GPU; tests/test_gpu_field.py checks the compiled kernels."""
import random

import pytest

from oracle import pasta as O
from tools.ptx_model import MASK, mont_mul, mont_sqr


def _patterns(m, py, n_rand):
    raw = [0, 1, 2, m - 1, m - 2, m >> 1, (1 << 254) - 1, 1 << 254, (1 << 254) + 1, MASK, (1 << 254) | MASK]
    raw += [(m - 1) & ~(MASK << (32 * k)) for k in range(8)]
    raw += [(m - 1) & ~(1 << (32 * k)) for k in range(8)]
    raw += [sum((MASK if py.random() < .5 else py.randrange(1 << 32)) << (32 * k) for k in range(8)) % m for _ in range(n_rand)]
    raw += [m - 1 - py.randrange(1 << 64) for _ in range(n_rand // 4)]
    raw += [py.randrange(m) for _ in range(n_rand)]
    return raw


@pytest.mark.parametrize("fid", [O.FIELD_FP, O.FIELD_FQ])
def test_ptx_model_mul(fid):
    m = O.MODULUS[fid]
    py = random.Random(fid)
    rinv = pow(1 << 256, -1, m)
    vals = _patterns(m, py, 1500)
    for a in vals[:60]:
        for b in vals[:60]:
            assert mont_mul(a, b, m)[0] == a * b * rinv % m
    for a in vals:
        b = vals[py.randrange(len(vals))]
        got, products = mont_mul(a, b, m)
        assert got == a * b * rinv % m, (hex(a), hex(b))
    assert products == 64 + 32      # 8 x 8 operand products + 4 reduction products per row


@pytest.mark.parametrize("fid", [O.FIELD_FP, O.FIELD_FQ])
def test_ptx_model_sqr(fid):
    m = O.MODULUS[fid]
    py = random.Random(10 + fid)
    rinv = pow(1 << 256, -1, m)
    for a in _patterns(m, py, 4000):
        got, products = mont_sqr(a, m)
        assert got == a * a * rinv % m, hex(a)
    assert products == 36 + 32
