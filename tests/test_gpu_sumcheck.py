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
