"""GPU parity for the R1CS kernels (shape source S) through the C ABI: multiply_vec,
cross-term T, fold, and a chain of device-resident NIFS steps whose folded instance must stay satisfied."""
import functools

import pytest

from oracle import pasta as O
from vdf_b200 import nova as N
from vdf_b200 import msm as G

pytestmark = pytest.mark.gpu


@functools.lru_cache(maxsize=None)
def _chain_state(fid, t, k):
    """State of the seed-42 chain after k + 1 segments of t rounds (the slow fifth-root direction, oracle)."""
    ovdf = O.MinRootVDF(fid)
    if k < 0:
        return O.State(O.field_random(O.XorShiftRng(), ovdf.m), 0, 1)
    return ovdf.eval(_chain_state(fid, t, k - 1), t)


def _instance(fid, t, k, aug):
    """k-th fresh instance of the step shape: result of a seed-42 chain advanced k + 1 segments."""
    return O.make_step_instance(fid, t, _chain_state(fid, t, k), aug_cons=aug)


@pytest.mark.parametrize("fid", [O.FIELD_FQ, O.FIELD_FP])
@pytest.mark.parametrize("t,aug", [(5, 0), (10, 200), (100, 500)])
def test_multiply_vec_and_cross_term(gpu_lib, fid, t, aug):
    shape, W1, X1, _ = _instance(fid, t, 0, aug)
    _, W2, X2, _ = _instance(fid, t, 1, aug)
    assert shape.num_cons == 3 * t + 1 + aug
    gs = N.R1CSShape(fid, shape.num_cons, shape.num_vars, shape.num_io, shape.A, shape.B, shape.C)
    z = shape.z_of(W1, 1, X1)
    assert list(gs.multiply_vec(z)) == shape.multiply_vec(z)
    # relaxed running instance with u != 1
    u1 = 0x1234567890ABCDEF
    T, _ = gs.commit_T(None, W1, u1, X1, W2, X2)
    assert T == shape.cross_term(W1, u1, X1, W2, X2)
    with pytest.raises(ValueError):
        gs.multiply_vec(z[:-1])


def test_fold_vectors(gpu_lib):
    m = O.Q
    rng = O.XorShiftRng()
    W1, W2 = [O.field_random(rng, m) for _ in range(1000)], [O.field_random(rng, m) for _ in range(1000)]
    E1, T = [O.field_random(rng, m) for _ in range(777)], [O.field_random(rng, m) for _ in range(777)]
    r = O.field_random(rng, m) >> 127
    W, E = N.fold_vectors(O.FIELD_FQ, W1, W2, E1, T, r)
    assert W == O.fold_vec(W1, W2, r, m) and E == O.fold_vec(E1, T, r, m)


def test_commit_T_commitment(gpu_lib):
    fid, cv = O.FIELD_FQ, O.PALLAS
    shape, W1, X1, _ = _instance(fid, 10, 0, 100)
    _, W2, X2, _ = _instance(fid, 10, 1, 100)
    gs = N.R1CSShape(fid, shape.num_cons, shape.num_vars, shape.num_io, shape.A, shape.B, shape.C)
    k0, d = 77, 13
    gens = G.Generators.progression(cv.cid, k0, d, max(shape.num_cons, shape.num_vars), table=True)
    T, cT = gs.commit_T(gens, W1, 1, X1, W2, X2)
    assert T == shape.cross_term(W1, 1, X1, W2, X2)
    assert cT == cv.msm_known_dlog(T, k0, d)


@pytest.mark.parametrize("fid,cid", [(O.FIELD_FQ, O.CURVE_PALLAS), (O.FIELD_FP, O.CURVE_VESTA)])
def test_running_prover_chain(gpu_lib, fid, cid):
    """Three NIFS-style folds with W, E resident on the device (test_nova_proof's t = 5, 3 steps,
    src/nova/proof.rs:403-451).  After every fold: Az.Bz = u.Cz + E (is_sat_relaxed) and the folded
    commitments equal fresh commitments of the folded vectors."""
    cv = O.CURVES[cid]
    t, aug = 5, 64
    shape, W0, X0, _ = _instance(fid, t, 0, aug)
    gs = N.R1CSShape(fid, shape.num_cons, shape.num_vars, shape.num_io, shape.A, shape.B, shape.C)
    k0, d = 4242, 17
    ngen = max(shape.num_cons, shape.num_vars)
    gens = G.Generators.progression(cid, k0, d, ngen, table=True)
    prover = N.RunningProver(gs, gens)
    E0 = [0] * shape.num_cons
    U = N.RelaxedR1CSInstance(comm_W=gens.commit(W0), comm_E=None, X=list(X0), u=1)
    assert U.comm_W == cv.msm_known_dlog(W0, k0, d)
    prover.set_running(W0, E0, U)
    Wr, Er, ur, Xr = list(W0), list(E0), 1, list(X0)
    for k in range(1, 4):
        _, W2, X2, _ = _instance(fid, t, k, aug)
        T_want = shape.cross_term(Wr, ur, Xr, W2, X2)
        comm_T, r = prover.prove_step(W2, X2)
        assert comm_T == cv.msm_known_dlog(T_want, k0, d)
        assert 0 < r < (1 << 128)
        Wr, Er = O.fold_vec(Wr, W2, r, shape.m), O.fold_vec(Er, T_want, r, shape.m)
        ur, Xr = (ur + r) % shape.m, [(a + r * b) % shape.m for a, b in zip(Xr, X2)]
        Wg, Eg, ug, Xg = prover.get_running()
        assert (Wg, Eg, ug, Xg) == (Wr, Er, ur, Xr)
        assert shape.is_sat_relaxed(Wg, Eg, ug, Xg)
        assert prover.U.u == ur and prover.U.X == Xr
        assert prover.U.comm_W == cv.msm_known_dlog(Wr, k0, d)
        assert prover.U.comm_E == cv.msm_known_dlog(Er, k0, d)


@pytest.mark.parametrize("fid,cid,t", [(O.FIELD_FQ, O.CURVE_PALLAS, 1024), (O.FIELD_FP, O.CURVE_VESTA, 1024),
                                       (O.FIELD_FQ, O.CURVE_PALLAS, 4096), (O.FIELD_FQ, O.CURVE_PALLAS, 16384)])
def test_fold_step_at_baseline_sizes(gpu_lib, fid, cid, t):
    """BASELINE config 3 sizes (t = 1024 / 4096 / 16384 MinRoot rounds per fold step + the ~10k-constraint augmented block):
    one NIFS fold against a RELAXED running instance (u != 1, E != 0), every output compared with the oracle --
    multiply_vec, the cross-term T, both commitments by the known-discrete-log identity, the folded W, E, u, X,
    and is_sat_relaxed of the result."""
    cv = O.CURVES[cid]
    aug = 9800
    shape, W0, X0, _ = _instance(fid, t, 0, aug)
    assert shape.num_cons == 3 * t + 1 + aug
    gs = N.R1CSShape(fid, shape.num_cons, shape.num_vars, shape.num_io, shape.A, shape.B, shape.C)
    z = shape.z_of(W0, 1, X0)
    assert list(gs.multiply_vec(z)) == shape.multiply_vec(z)
    k0, d = 991, 5
    gens = G.Generators.progression(cid, k0, d, max(shape.num_cons, shape.num_vars), table=True)
    prover = N.RunningProver(gs, gens)
    # make the running instance properly relaxed first: fold instance 1 into instance 0 with the oracle
    _, W1, X1, _ = _instance(fid, t, 1, aug)
    r0 = 0x1F2E3D4C5B6A79880123456789ABCDEF
    T0 = shape.cross_term(W0, 1, X0, W1, X1)
    Wr, Er = O.fold_vec(W0, W1, r0, shape.m), O.fold_vec([0] * shape.num_cons, T0, r0, shape.m)
    ur, Xr = (1 + r0) % shape.m, [(a + r0 * b) % shape.m for a, b in zip(X0, X1)]
    assert shape.is_sat_relaxed(Wr, Er, ur, Xr)
    U = N.RelaxedR1CSInstance(comm_W=cv.msm_known_dlog(Wr, k0, d), comm_E=cv.msm_known_dlog(Er, k0, d), X=list(Xr), u=ur)
    prover.set_running(Wr, Er, U)
    # the GPU fold of instance 2
    _, W2, X2, _ = _instance(fid, t, 2, aug)
    T_want = shape.cross_term(Wr, ur, Xr, W2, X2)
    comm_T, r = prover.prove_step(W2, X2)
    assert comm_T == cv.msm_known_dlog(T_want, k0, d)
    Wn, En = O.fold_vec(Wr, W2, r, shape.m), O.fold_vec(Er, T_want, r, shape.m)
    un, Xn = (ur + r) % shape.m, [(a + r * b) % shape.m for a, b in zip(Xr, X2)]
    assert prover.get_running() == (Wn, En, un, Xn)
    assert shape.is_sat_relaxed(Wn, En, un, Xn)
    assert prover.U.comm_W == cv.msm_known_dlog(Wn, k0, d)
    assert prover.U.comm_E == cv.msm_known_dlog(En, k0, d)


def test_r1cs_argument_errors(gpu_lib):
    from vdf_b200 import VdfGpuError
    with pytest.raises(VdfGpuError):
        N.R1CSShape(O.FIELD_FQ, 2, 2, 0, [(5, 0, 1)], [], [])   # row out of range
    with pytest.raises(VdfGpuError):
        N.R1CSShape(O.FIELD_FQ, 2, 2, 0, [(0, 3, 1)], [], [])   # column out of range
