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



def _gpu_shape(shape):
    from vdf_b200 import nova as N
    fid = O.FIELD_FP if shape.m == O.P else O.FIELD_FQ
    return N.R1CSShape(fid, shape.num_cons, shape.num_vars, shape.num_io, shape.A, shape.B, shape.C)


@pytest.mark.parametrize("fid", [O.FIELD_FP, O.FIELD_FQ])
def test_bind_rows_matches_oracle(gpu_lib, fid):
    """The inner sum-check's table: transposed sparse product over the column view of the CSR; the step circuit's
    constant column holds an entry per round (the heavy column a warp sums cooperatively)."""
    m = O.MODULUS[fid]
    py = random.Random(60 + fid)
    for t, aug in ((1, 0), (100, 0), (10, 300)):   # t = 100: the constant column has > 64 entries (BindHeavyFn)
        shape, W, X, _ = O.make_step_instance(fid, t, O.State(py.randrange(m), py.randrange(m), t), aug_cons=aug)
        gs = _gpu_shape(shape)
        eq = [py.randrange(m) for _ in range(shape.num_cons)]
        for r_abc in ([py.randrange(m) for _ in range(3)], [1, 0, 0], [0, 0, m - 1]):
            assert gs.bind_rows(eq, r_abc) == shape.bind_rows(eq, r_abc)
        with pytest.raises(ValueError):
            gs.bind_rows(eq[:-1], [1, 2, 3])
        gs.close()


def _interp(points, r, m):
    """value at r of the polynomial through (x_k, y_k)"""
    acc = 0
    for k, (xk, yk) in enumerate(points):
        num, den = 1, 1
        for j, (xj, _) in enumerate(points):
            if j != k:
                num = num * (r - xj) % m
                den = den * (xk - xj) % m
        acc = (acc + yk * num * pow(den, -1, m)) % m
    return acc


def _verify_rounds(claim, evals, rs, m, cubic):
    """The sum-check verifier: e(0) + e(1) = claim, next claim = e(r); e(1) is implied by the claim."""
    for e, r in zip(evals, rs):
        e1 = (claim - e[0]) % m
        pts = [(0, e[0]), (1, e1), (2, e[1])] + ([(3, e[2])] if cubic else [])
        claim = _interp(pts, r, m)
    return claim


def test_spartan_sumcheck_phases_on_a_folded_relaxed_instance(gpu_lib):
    """Both sum-check phases of RelaxedR1CSSNARK::prove (nova-snark spartan_with_ipa_pc [R], from
    CompressedSNARK::prove, src/nova/proof.rs:363) on the GPU building blocks, checked by a verifier written here:
    outer  sum_x eq(tau, x) (Az(x) Bz(x) - (u Cz(x) + E(x))) = 0,
    inner  sum_y (rA A + rB B + rC C)(rx, y) z(y) = rA Az(rx) + rB Bz(rx) + rC Cz(rx),
    on an instance with u != 1 and E != 0 (two step instances folded by the oracle)."""
    from vdf_b200 import nova as N
    fid, m = O.FIELD_FQ, O.Q
    py = random.Random(77)
    shape, W1, X1, _ = O.make_step_instance(fid, 12, O.State(5, 6, 12), aug_cons=100)
    _, W2, X2, _ = O.make_step_instance(fid, 12, O.State(9, 1, 12), aug_cons=100)
    r_fold = py.randrange(1 << 128)
    T = shape.cross_term(W1, 1, X1, W2, X2)
    W, E = O.fold_vec(W1, W2, r_fold, m), [r_fold * t % m for t in T]
    X, u = O.fold_vec(X1, X2, r_fold, m), (1 + r_fold) % m
    assert shape.is_sat_relaxed(W, E, u, X) and any(E)
    z = shape.z_of(W, u, X)
    gs = _gpu_shape(shape)
    Az, Bz, Cz = gs.multiply_vec(z)

    def pad(v, ell):
        return list(v) + [0] * ((1 << ell) - len(v))

    def challenge(rnd, evals):
        return (sum(evals) * 0x9E3779B97F4A7C15 + rnd + 999) % m

    # outer phase
    ell_x = max(1, (shape.num_cons - 1).bit_length())
    tau = [py.randrange(m) for _ in range(ell_x)]
    D = [(u * c + e) % m for c, e in zip(Cz, E)]
    tabs = [SP.eq_evals(fid, tau), pad(Az, ell_x), pad(Bz, ell_x), pad(D, ell_x)]
    evals, rx, fin = SP.sumcheck(fid, tabs, challenge)
    last = _verify_rounds(0, evals, rx, m, cubic=True)
    assert last == fin[0] * (fin[1] * fin[2] - fin[3]) % m
    assert fin[0] == O.poly_evaluate(O.eq_evals(tau, m), rx, m)
    claim_Az, claim_Bz = fin[1], fin[2]
    claim_Cz = SP.poly_evaluate(fid, pad(Cz, ell_x), rx)
    assert fin[3] == (u * claim_Cz + SP.poly_evaluate(fid, pad(E, ell_x), rx)) % m

    # inner phase
    r_abc = [py.randrange(m) for _ in range(3)]
    eq_rx = SP.eq_evals(fid, rx)[:shape.num_cons]
    Mt = gs.bind_rows(eq_rx, r_abc)
    ell_y = max(1, (len(z) - 1).bit_length())
    claim = (r_abc[0] * claim_Az + r_abc[1] * claim_Bz + r_abc[2] * claim_Cz) % m
    evals, ry, fin = SP.sumcheck(fid, [pad(Mt, ell_y), pad(z, ell_y)], challenge)
    last = _verify_rounds(claim, evals, ry, m, cubic=False)
    assert last == fin[0] * fin[1] % m
    # what the verifier recomputes itself: the sparse matrices at (rx, ry), and z(ry)
    eq_ry = O.eq_evals(ry, m)
    eq_rx_full = O.eq_evals(rx, m)
    want = 0
    for M, rm in zip((shape.A, shape.B, shape.C), r_abc):
        want = (want + rm * sum(v * eq_rx_full[r] % m * eq_ry[c] for r, c, v in M)) % m
    assert fin[0] == want
    assert fin[1] == O.poly_evaluate(pad(z, ell_y), ry, m)
    gs.close()


@pytest.mark.parametrize("cid", [O.CURVE_PALLAS, O.CURVE_VESTA])
def test_inner_product_argument_end_to_end(gpu_lib, cid):
    """All log2(n) rounds of an inner-product argument (the polynomial-evaluation argument of spartan_with_ipa_pc [R],
    from CompressedSNARK::prove, src/nova/proof.rs:363) on the GPU blocks: per round two cross inner products, two
    commitments over generator halves (the drop-in MSM symbol), the folds of a, b and of the generators.  The verifier
    below uses Python integers only: P' = r^2 L + P + r^-2 R per round, and at the end P_final = a (G_final + b U)."""
    from vdf_b200 import msm as G
    from vdf_b200.encoding import CURVE_BASE, point_from_bytes
    cv = O.CURVES[cid]
    fid = 1 if cid == 0 else 0
    m = cv.order
    py = random.Random(90 + cid)
    n = 64
    gens = [cv.mul(py.randrange(1, m), cv.gen) for _ in range(n)]
    U = cv.mul(py.randrange(1, m), cv.gen)
    a = [py.randrange(m) for _ in range(n)]
    b = O.eq_evals([py.randrange(m) for _ in range(6)], m)      # the evaluation point's eq table
    c = SP.inner_product(fid, a, b)

    def commit(points_bytes, scalars):
        raw = G.mult_pippenger(cid, points_bytes, O.fes_to_bytes(scalars, m), True)
        return point_from_bytes(raw, CURVE_BASE[cid])

    Gb = O.affines_to_bytes(cv, gens)
    P = cv.add(commit(Gb, a), cv.mul(c, U))                     # what the verifier holds: commitment + claimed value
    assert commit(Gb, a) == cv.msm(a, gens)
    while n > 1:
        h = n // 2
        a_lo, a_hi, b_lo, b_hi = a[:h], a[h:], b[:h], b[h:]
        G_lo, G_hi = Gb[:72 * h], Gb[72 * h:]
        cL, cR = SP.inner_product(fid, a_lo, b_hi), SP.inner_product(fid, a_hi, b_lo)
        L = cv.add(commit(G_hi, a_lo), cv.mul(cL, U))
        R = cv.add(commit(G_lo, a_hi), cv.mul(cR, U))
        r = py.randrange(1, m)                                   # the transcript's challenge
        r_inv = pow(r, -1, m)
        a = SP.vec_lincomb(fid, a_lo, a_hi, r, r_inv)
        b = SP.vec_lincomb(fid, b_lo, b_hi, r_inv, r)
        Gb = SP.points_lincomb(cid, G_lo, G_hi, r_inv, r)
        # verifier
        P = cv.add(cv.add(cv.mul(r * r % m, L), P), cv.mul(r_inv * r_inv % m, R))
        n = h
    G_final = O.affine_from_bytes(cv, Gb[:72])
    assert P == cv.mul(a[0], cv.add(G_final, cv.mul(b[0], U)))
