"""GPU parity for the sum-check building blocks (CompressedSNARK::prove's non-MSM work) through the
C ABI against the oracle: eq tables, every round's evaluations, the bound tables' final evaluations, polynomial
evaluation; both fields; edge sizes."""
import random

import pytest

from oracle import pasta as O
from vdf_b200 import spartan as SP

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fid", [O.FIELD_FP, O.FIELD_FQ])
def test_eq_evals_and_evaluate(gpu_lib, fid):
    m = O.MODULUS[fid]
    py = random.Random(fid)
    for ell in (0, 1, 2, 5, 11):
        r = [py.randrange(m) for _ in range(ell)]
        assert SP.eq_evals(fid, r) == O.eq_evals(r, m)
        poly = [py.randrange(m) for _ in range(1 << ell)]
        assert SP.poly_evaluate(fid, poly, r) == O.poly_evaluate(poly, r, m)
    # boolean points select single entries
    assert SP.eq_evals(fid, [1, 0, 1]) == [0, 0, 0, 0, 0, 1, 0, 0]


@pytest.mark.parametrize("fid", [O.FIELD_FP, O.FIELD_FQ])
@pytest.mark.parametrize("ntab,ell", [(4, 1), (4, 7), (4, 12), (2, 1), (2, 9), (2, 13)])
def test_sumcheck_rounds_match_oracle(gpu_lib, fid, ntab, ell):
    m = O.MODULUS[fid]
    py = random.Random(ell * 10 + ntab)
    tabs = [[py.randrange(m) for _ in range(1 << ell)] for _ in range(ntab)]
    tabs[0][3 % (1 << ell)] = 0
    tabs[1][0] = m - 1

    def challenge(rnd, evals):          # deterministic, depends on the transcript so far
        return (sum(evals) * 0x9E3779B97F4A7C15 + rnd + 12345) % m

    want = O.sumcheck_prove(tabs, m, challenge)
    got = SP.sumcheck(fid, tabs, challenge)
    assert got[0] == [tuple(e) for e in want[0]]
    assert got[1] == want[1] and got[2] == want[2]
    # the final evaluations are the tables evaluated at the challenge point
    assert got[2] == [O.poly_evaluate(t, got[1], m) for t in tabs]


def test_sumcheck_callback_failure_and_arguments(gpu_lib):
    from vdf_b200 import VdfGpuError
    m = O.Q
    tabs = [[1, 2, 3, 4]] * 2

    def boom(rnd, evals):
        raise RuntimeError("transcript failed")

    with pytest.raises(VdfGpuError):
        SP.sumcheck(O.FIELD_FQ, tabs, boom)
    with pytest.raises(ValueError):
        SP.sumcheck(O.FIELD_FQ, [[1, 2, 3]] * 2, lambda r, e: 1)
    # the library is usable afterwards
    assert SP.eq_evals(O.FIELD_FQ, [5]) == [(1 - 5) % m, 5]


@pytest.mark.parametrize("cid", [O.CURVE_PALLAS, O.CURVE_VESTA])
def test_ipa_round_building_blocks(gpu_lib, cid):
    """One round of the inner-product argument (CommitGens::fold, the vector folds, the cross inner products) against
    Python integers, including identity generators, equal points (doubling inside P + Q) and P = -Q."""
    cv = O.CURVES[cid]
    fid = 1 if cid == 0 else 0          # scalar field of the curve
    m = cv.order
    py = random.Random(40 + cid)
    n = 37
    gens = [cv.mul(py.randrange(1, m), cv.gen) for _ in range(2 * n)]
    gens[3] = None                       # identity on the left
    gens[n + 5] = None                   # identity on the right
    gens[n + 7] = gens[7]                # P == Q
    gens[n + 9] = cv.neg(gens[9])        # P == -Q
    r = py.randrange(1, m)
    r_inv = pow(r, -1, m)
    L, R = gens[:n], gens[n:]
    got = SP.points_lincomb(cid, O.affines_to_bytes(cv, L), O.affines_to_bytes(cv, R), r_inv, r)
    want = [cv.add(cv.mul(r_inv, a), cv.mul(r, b)) for a, b in zip(L, R)]
    assert [O.affine_from_bytes(cv, got[72 * i:72 * i + 72]) for i in range(n)] == want
    # folding preserves the commitment: <a', G'> = r^-1... check the algebra the IPA relies on
    a = [py.randrange(m) for _ in range(2 * n)]
    b = [py.randrange(m) for _ in range(2 * n)]
    a2 = SP.vec_lincomb(fid, a[:n], a[n:], r, r_inv)
    b2 = SP.vec_lincomb(fid, b[:n], b[n:], r_inv, r)
    assert a2 == [(r * x + r_inv * y) % m for x, y in zip(a[:n], a[n:])]
    cL, cR = SP.inner_product(fid, a[:n], b[n:]), SP.inner_product(fid, a[n:], b[:n])
    assert cL == sum(x * y for x, y in zip(a[:n], b[n:])) % m
    c = sum(x * y for x, y in zip(a, b)) % m
    assert SP.inner_product(fid, a2, b2) == (c + r * r * cL + r_inv * r_inv * cR) % m
    assert SP.inner_product(fid, [], []) == 0
    # special scalars
    assert SP.points_lincomb(cid, O.affines_to_bytes(cv, L[:4]), O.affines_to_bytes(cv, R[:4]), 0, 1) == O.affines_to_bytes(cv, R[:4])
    assert SP.points_lincomb(cid, O.affines_to_bytes(cv, L[:4]), O.affines_to_bytes(cv, R[:4]), 1, 0) == O.affines_to_bytes(cv, L[:4])

