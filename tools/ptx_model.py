"""Carry-flag-exact Python model of the PTX Montgomery multiplication and of the dedicated squaring in
vdf_b200/csrc/field.cuh (even/odd accumulator rows, modulus limbs [1, M1, M2, M3, 0, 0, 0, 2^30], q = -t0).
Every PTX carry chain of the CUDA code is replayed instruction by instruction; an instruction that would drop a
carry the CUDA code does not propagate raises AssertionError.  tests/test_ptx_model.py runs it against big-integer
arithmetic on the CPU, so the algorithm is checked without a GPU; tests/test_gpu_field.py checks the real kernels."""
MASK = 0xffffffff


class Flags:
    """The CC.CF carry flag + the PTX instructions the multiplier uses."""

    def __init__(s):
        s.cf = 0
        s.products = 0

    def add_cc(s, a, b): t = a + b; s.cf = t >> 32; return t & MASK
    def addc_cc(s, a, b): t = a + b + s.cf; s.cf = t >> 32; return t & MASK

    def addc(s, a, b):
        t = a + b + s.cf
        assert t >> 32 == 0, "lost carry (addc)"
        return t & MASK

    def mad_lo_cc(s, a, b, c): s.products += 1; t = ((a * b) & MASK) + c; s.cf = t >> 32; return t & MASK
    def madc_lo_cc(s, a, b, c): s.products += 1; t = ((a * b) & MASK) + c + s.cf; s.cf = t >> 32; return t & MASK
    def madc_hi_cc(s, a, b, c): t = ((a * b) >> 32) + c + s.cf; s.cf = t >> 32; return t & MASK

    def madc_hi(s, a, b, c):
        t = ((a * b) >> 32) + c + s.cf
        assert t >> 32 == 0, "lost carry (madc.hi)"
        return t & MASK


def limbs(x, n=8):
    return [(x >> (32 * i)) & MASK for i in range(n)]


def _mul_n(m, acc, a, off, bi):
    for j in range(0, 8, 2):
        p = a[off + j] * bi if off + j < 8 else 0
        m.products += 1
        acc[j] = p & MASK
        acc[j + 1] = p >> 32


def _redc_row(m, even, odd, M1, M2, M3, M7):
    mi = (-even[0]) & MASK
    odd[0] = m.mad_lo_cc(mi, M1, odd[0]); odd[1] = m.madc_hi_cc(mi, M1, odd[1])
    odd[2] = m.madc_lo_cc(mi, M3, odd[2]); odd[3] = m.madc_hi_cc(mi, M3, odd[3])
    odd[4] = m.addc_cc(odd[4], 0); odd[5] = m.addc_cc(odd[5], 0)
    odd[6] = m.madc_lo_cc(mi, M7, odd[6]); odd[7] = m.madc_hi(mi, M7, odd[7])
    even[0] = m.add_cc(even[0], mi)
    assert even[0] == 0
    even[1] = m.addc_cc(even[1], 0)
    even[2] = m.madc_lo_cc(mi, M2, even[2]); even[3] = m.madc_hi_cc(mi, M2, even[3])
    for k in range(4, 8):
        even[k] = m.addc_cc(even[k], 0)
    odd[7] = m.addc(odd[7], 0)


# first += v[0,2,4,6] * bi, chain starting at the first non-zero limb (ze leading zero even limbs); carry -> top[7]
def _cmad_top(m, acc, v, bi, top, ze=0):
    if ze >= 4:
        return
    for k in range(ze, 4):
        j = 2 * k
        acc[j] = m.mad_lo_cc(v[j], bi, acc[j]) if k == ze else m.madc_lo_cc(v[j], bi, acc[j])
        acc[j + 1] = m.madc_hi_cc(v[j], bi, acc[j + 1])
    top[7] = m.addc(top[7], 0)


# even[0] += odd[1]; odd = (odd >> 64) + v[1,3,5,7] * bi; the zo lowest odd limbs are zero: plain carry adds
def _shift_mad(m, even, odd, v, bi, zo=0):
    even[0] = m.add_cc(even[0], odd[1])
    for k in range(4):
        j, limb = 2 * k, 2 * k + 1
        c_lo = odd[j + 2] if j + 2 < 8 else 0
        c_hi = odd[j + 3] if j + 3 < 8 else 0
        if k < zo:
            assert v[limb] == 0
            odd[j] = m.addc_cc(c_lo, 0); odd[j + 1] = m.addc_cc(c_hi, 0)
        else:
            odd[j] = m.madc_lo_cc(v[limb], bi, c_lo)
            odd[j + 1] = m.madc_hi(v[limb], bi, c_hi) if k == 3 else m.madc_hi_cc(v[limb], bi, c_hi)


def _finish(m, even, odd, mod):
    even[0] = m.add_cc(even[0], odd[1])
    for j in range(1, 7):
        even[j] = m.addc_cc(even[j], odd[j + 1])
    even[7] = m.addc(even[7], 0)
    r = sum(even[j] << (32 * j) for j in range(8))
    if r >= mod:
        r -= mod
    assert r < mod, "result not reduced by one subtraction"
    return r


def mont_mul(a, b, mod):
    """(a * b / 2^256) mod `mod` the way Field::mul computes it; returns (result, lo/hi product pairs)."""
    ml = limbs(mod)
    M1, M2, M3, M7 = ml[1], ml[2], ml[3], ml[7]
    m = Flags()
    A, B = limbs(a), limbs(b)
    even, odd = [0] * 8, [0] * 8

    def row(first, second, bi, is_first):
        if is_first:
            _mul_n(m, second, A, 1, bi); _mul_n(m, first, A, 0, bi)
        else:
            _shift_mad(m, first, second, A, bi)
            _cmad_top(m, first, A, bi, second)
        _redc_row(m, first, second, M1, M2, M3, M7)
    for i in range(0, 8, 2):
        row(even, odd, B[i], i == 0)
        row(odd, even, B[i + 1], False)
    return _finish(m, even, odd, mod), m.products


def mont_sqr(a, mod):
    """(a * a / 2^256) mod `mod` the way Field::sqr computes it: row i multiplies a_i by
    V_i = [0 (j < i), a_i, (2a)_(i+1) & ~1, (2a)_(i+2), ...] and skips the products of the zero limbs."""
    ml = limbs(mod)
    M1, M2, M3, M7 = ml[1], ml[2], ml[3], ml[7]
    m = Flags()
    assert 2 * a < (1 << 256)
    A, D = limbs(a), limbs(2 * a)
    even, odd = [0] * 8, [0] * 8

    def row(first, second, i):
        v = [0] * 8
        v[i] = A[i]
        if i + 1 < 8:
            v[i + 1] = D[i + 1] & ~1 & MASK
        for j in range(i + 2, 8):
            v[j] = D[j]
        if i == 0:
            _mul_n(m, second, v, 1, A[0]); _mul_n(m, first, v, 0, A[0])
        else:
            _shift_mad(m, first, second, v, A[i], zo=i // 2)
            _cmad_top(m, first, v, A[i], second, ze=(i + 1) // 2)
        _redc_row(m, first, second, M1, M2, M3, M7)
    for i in range(0, 8, 2):
        row(even, odd, i)
        row(odd, even, i + 1)
    return _finish(m, even, odd, mod), m.products
