"""CPU: pins the C restatement (oracle/cpu_ref.c, the CPU baseline) against the Python-integer oracle."""
import random

import pytest

from oracle import cpu_ref as C
from oracle import pasta as O
from tests.util import edge_field_values, nova_like_scalars, rand_scalars


@pytest.mark.parametrize("fid", [O.FIELD_FP, O.FIELD_FQ])
def test_field_mul(fid):
    m = O.MODULUS[fid]
    rng = O.XorShiftRng()
    edge = edge_field_values(m)
    A = [a for a in edge for _ in edge] + rand_scalars(rng, m, 500)
    B = [b for _ in edge for b in edge] + rand_scalars(rng, m, 500)
    got = O.fes_from_bytes(C.field_mul(fid, O.fes_to_bytes(A, m), O.fes_to_bytes(B, m)), m)
    assert got == [a * b % m for a, b in zip(A, B)]


@pytest.mark.parametrize("cid", [O.CURVE_PALLAS, O.CURVE_VESTA])
def test_progression_and_msm(cid):
    cv = O.CURVES[cid]
    rng, py = O.XorShiftRng(), random.Random(2)
    n = 1500
    k0, d = 11, 29
    pb = C.progression(cid, k0, d, n)
    assert pb[:72 * 40] == O.affines_to_bytes(cv, cv.progression(k0, d, 40))
    for sc in (rand_scalars(rng, cv.order, n), nova_like_scalars(py, rng, cv.order, n)):
        sc[0], sc[1] = 0, cv.order - 1
        want = O.jac_to_bytes(cv, cv.msm_known_dlog(sc, k0, d))
        for threads in (1, 4):
            assert C.msm(cid, pb, O.fes_to_bytes(sc, cv.order), True, threads) == want
        assert C.msm(cid, pb, b"".join(s.to_bytes(32, "little") for s in sc), False, 2) == want
    assert C.msm(cid, b"", b"", True, 1) == bytes(96)
    # small n (single window slice), identity point, P / -P
    pts = cv.progression(3, 1, 10)
    pts[2] = None
    pts[5] = cv.neg(pts[4])
    sc = [7, 0, 9, 1, 5, 5, 2, 3, cv.order - 2, 1 << 200]
    assert C.msm(cid, O.affines_to_bytes(cv, pts), O.fes_to_bytes(sc, cv.order), True, 2) == O.jac_to_bytes(cv, cv.msm_naive(sc, pts))


def test_msm_2_16_known_dlog():
    cv = O.PALLAS
    py = random.Random(7)
    n, k0, d = 1 << 16, 5, 3
    # progression via the oracle's O(n) generator is slow in Python; cpu_ref's own generator was pinned above
    pb = C.progression(cv.cid, k0, d, n)
    sc = [py.randrange(cv.order) for _ in range(n)]
    assert C.msm(cv.cid, pb, O.fes_to_bytes(sc, cv.order)) == O.jac_to_bytes(cv, cv.msm_known_dlog(sc, k0, d))


@pytest.mark.parametrize("fid,mk", [(O.FIELD_FQ, O.PallasVDF), (O.FIELD_FP, O.VestaVDF)])
def test_minroot_check(fid, mk):
    vdf = mk()
    rng = O.XorShiftRng()
    res, orig, ts = [], [], []
    for k in range(200):
        r = O.State(*rand_scalars(rng, vdf.m, 3))
        t = [0, 1, 5, 10][k % 4]
        o = vdf.inverse_eval(r, t)
        if k % 7 == 3:
            o = O.State((o.x + 1) % vdf.m, o.y, o.i)
        res.append(r); orig.append(o); ts.append(t)
    rb = b"".join(O.state_to_bytes(s, vdf.m) for s in res)
    ob = b"".join(O.state_to_bytes(s, vdf.m) for s in orig)
    assert [bool(b) for b in C.minroot_check(fid, rb, ob, ts, 3)] == [vdf.check(r, t, o) for r, t, o in zip(res, ts, orig)]


def test_r1cs_and_fold():
    fid = O.FIELD_FQ
    vdf = O.PallasVDF()
    m = vdf.m
    rng = O.XorShiftRng()
    s = vdf.eval(O.State(O.field_random(rng, m), 0, 1), 8)
    shape, W1, X1, _ = O.make_step_instance(fid, 8, s, aug_cons=60)
    _, W2, X2, _ = O.make_step_instance(fid, 8, vdf.eval(s, 8), aug_cons=60)
    coo = O.shape_to_coo_bytes(shape)
    u1 = 0xDEADBEEF
    abc1 = C.multiply_vec(fid, shape.num_cons, shape.num_vars, shape.num_io, coo, O.fes_to_bytes(W1, m),
                          O.fe_to_bytes(u1, m), O.fes_to_bytes(X1, m))
    Az, Bz, Cz = shape.multiply_vec(shape.z_of(W1, u1, X1))
    assert O.fes_from_bytes(abc1, m) == Az + Bz + Cz
    abc2 = C.multiply_vec(fid, shape.num_cons, shape.num_vars, shape.num_io, coo, O.fes_to_bytes(W2, m),
                          O.fe_to_bytes(1, m), O.fes_to_bytes(X2, m))
    T = C.cross_term(fid, shape.num_cons, abc1, abc2, O.fe_to_bytes(u1, m))
    assert O.fes_from_bytes(T, m) == shape.cross_term(W1, u1, X1, W2, X2)
    r = O.field_random(rng, m) >> 127
    assert O.fes_from_bytes(C.fold(fid, O.fes_to_bytes(W1, m), O.fes_to_bytes(W2, m), O.fe_to_bytes(r, m)), m) == O.fold_vec(W1, W2, r, m)
