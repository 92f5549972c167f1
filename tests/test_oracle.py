"""CPU: pins the oracle.  The reference holds no golden values for this path, so the pins
are (1) the reference's own constants and addition chains, (2) its round-trip tests restated, (3) algebraic
identities between independent oracle routines, (4) the frozen fixtures under tests/golden."""
import json
from pathlib import Path

import pytest

from oracle import pasta as O

GOLDEN = Path(__file__).parent / "golden" / "vectors.json"


def test_moduli_and_reference_constants():
    # src/minroot.rs:273-285: 5 * INVALPHA == 1 mod (m - 1), and INVALPHA == (4(m-1)+1)/5
    for limbs, m in ((O.FP_RESCUE_INVALPHA, O.P), (O.FQ_RESCUE_INVALPHA, O.Q)):
        e = O.limbs_to_int(limbs)
        assert 5 * e % (m - 1) == 1
        assert e == (4 * (m - 1) + 1) // 5
        assert m.bit_length() == 255 and (m - 1) % 5 == 1
        assert (-pow(m, -1, 1 << 32)) % (1 << 32) == 0xFFFFFFFF
    assert O.P % (1 << 128) != O.Q % (1 << 128) and O.P >> 128 == O.Q >> 128 == 1 << 126


def test_exponents():  # src/minroot.rs:449-458
    assert O.PallasVDF().inverse_exponent == 5 and O.VestaVDF().inverse_exponent == 5


@pytest.mark.parametrize("mk", [O.PallasVDF, O.VestaVDF])
def test_steps(mk):  # src/minroot.rs:460-477, plus the reference's addition chains (:88-127, :223-261)
    rng = O.XorShiftRng()
    vdf = mk()
    for _ in range(100):
        x = O.field_random(rng, vdf.m)
        y = vdf.forward_step(x)
        assert vdf.inverse_step(y) == x
        assert vdf.forward_step_addition_chain(x) == y


def test_eval():  # src/minroot.rs:479-510
    rng = O.XorShiftRng()
    vdf = O.PallasVDF()
    for _ in range(10):
        x = O.State(O.field_random(rng, vdf.m), O.field_random(rng, vdf.m), 0)
        result = vdf.eval(x, 10)
        assert vdf.inverse_eval(result, 10) == x
        assert vdf.check(result, 10, x)


@pytest.mark.parametrize("mk", [O.PallasVDF, O.VestaVDF])
def test_vanilla_proof(mk):  # src/minroot.rs:512-542
    rng = O.XorShiftRng()
    vdf = mk()
    x = O.State(O.field_random(rng, vdf.m), 0, 0)
    t, n = 4, 3
    _, final = O.Evaluation.eval(vdf, x, t)
    for _ in range(1, n):
        _, new = O.Evaluation.eval(vdf, final.result, t)
        final = final.append(new)
        assert final is not None
    assert vdf.element(final.t) == final.result.i and final.t == n * t and final.verify(x)


def test_xorshift_generator():
    # rand_xorshift: state words are the seed's four LE u32s; first output from the documented recurrence
    rng = O.XorShiftRng()
    x = 0x2A2A2A2A
    t = (x ^ (x << 11)) & 0xFFFFFFFF
    assert rng.next_u32() == (x ^ (x >> 19) ^ (t ^ (t >> 8))) & 0xFFFFFFFF


@pytest.mark.parametrize("cv", [O.PALLAS, O.VESTA])
def test_curve_group_law(cv):
    g = cv.gen
    assert cv.on_curve(g)
    assert cv.mul(cv.order, g) is None and cv.mul(cv.order - 1, g) == cv.neg(g)
    a, b = cv.mul(12345, g), cv.mul(67890, g)
    assert cv.add(a, b) == cv.mul(12345 + 67890, g) and cv.add(a, cv.neg(a)) is None
    assert cv.add(a, a) == cv.mul(2 * 12345, g)
    # byte layouts round-trip
    assert O.affine_from_bytes(cv, O.affine_to_bytes(cv, a)) == a
    assert O.affine_from_bytes(cv, O.affine_to_bytes(cv, None)) is None
    assert O.jac_from_bytes(cv, O.jac_to_bytes(cv, a)) == a and O.jac_from_bytes(cv, bytes(96)) is None


@pytest.mark.parametrize("cv", [O.PALLAS, O.VESTA])
def test_msm_three_ways(cv):
    rng = O.XorShiftRng()
    n = 48
    pts = cv.progression(9, 4, n)
    sc = [O.field_random(rng, cv.order) for _ in range(n)]
    a = cv.msm_naive(sc, pts)
    assert a == cv.msm(sc, pts) == cv.msm(sc, pts, c=5) == cv.msm_known_dlog(sc, 9, 4)


@pytest.mark.parametrize("t", [5, 10, 100])
def test_step_circuit_shape(t):  # src/nova/proof.rs:87-230; dimensions of SURVEY.md section 8
    vdf = O.PallasVDF()
    rng = O.XorShiftRng()
    s0 = O.State(O.field_random(rng, vdf.m), 0, 1)
    res = vdf.eval(s0, t)
    shape, W, X, outs = O.make_step_instance(O.FIELD_FQ, t, res)
    assert shape.num_cons == 3 * t + 1
    assert shape.num_vars == 4 * t + 1 + 3            # + the three z_in variables
    assert (len(shape.A), len(shape.B), len(shape.C)) == (3 * t + 1, 3 * t + 1, 6 * t + 2)
    assert outs == s0                                  # circuit output == inverse_eval(result, t)
    assert shape.is_sat_relaxed(W, [0] * shape.num_cons, 1, X)
    Wbad = list(W)
    Wbad[5] = (Wbad[5] + 1) % vdf.m
    assert not shape.is_sat_relaxed(Wbad, [0] * shape.num_cons, 1, X)


def test_fold_preserves_relaxed_satisfiability():
    vdf = O.PallasVDF()
    rng = O.XorShiftRng()
    s = vdf.eval(O.State(O.field_random(rng, vdf.m), 0, 1), 10)
    shape, W, X, _ = O.make_step_instance(O.FIELD_FQ, 10, s, aug_cons=40)
    E, u = [0] * shape.num_cons, 1
    for k in range(2):
        s = vdf.eval(s, 10)
        _, W2, X2, _ = O.make_step_instance(O.FIELD_FQ, 10, s, aug_cons=40)
        T = shape.cross_term(W, u, X, W2, X2)
        r = O.field_random(rng, vdf.m) >> 127
        W, E = O.fold_vec(W, W2, r, vdf.m), O.fold_vec(E, T, r, vdf.m)
        u, X = (u + r) % vdf.m, [(a + r * b) % vdf.m for a, b in zip(X, X2)]
        assert shape.is_sat_relaxed(W, E, u, X)


def test_golden_vectors():
    """Frozen known-answer vectors (tests/golden/make_golden.py): any drift in the oracle shows here."""
    g = json.loads(GOLDEN.read_text())
    for fname, m in (("fp", O.P), ("fq", O.Q)):
        for a, b, ab_mont in g["field_mul_mont"][fname]:
            a, b = int(a, 16), int(b, 16)
            assert O.fe_to_bytes(a * b % m, m).hex() == ab_mont
    for name, mk in (("pallas", O.PallasVDF), ("vesta", O.VestaVDF)):
        vdf = mk()
        for rec in g["minroot"][name]:
            s = O.State(*[int(v, 16) for v in rec["start"]])
            r = vdf.eval(s, rec["t"])
            assert [hex(r.x), hex(r.y), hex(r.i)] == rec["result"]
            assert vdf.check(r, rec["t"], s)
    for cname, cv in (("pallas", O.PALLAS), ("vesta", O.VESTA)):
        rec = g["msm"][cname]
        rng = O.XorShiftRng()
        sc = [O.field_random(rng, cv.order) for _ in range(rec["n"])]
        pts = cv.progression(rec["k0"], rec["d"], rec["n"])
        assert O.jac_to_bytes(cv, cv.msm(sc, pts)).hex() == rec["result_point96"]
    rec = g["cross_term"]
    vdf = O.PallasVDF()
    rng = O.XorShiftRng()
    s = vdf.eval(O.State(O.field_random(rng, vdf.m), 0, 1), rec["t"])
    shape, W1, X1, _ = O.make_step_instance(O.FIELD_FQ, rec["t"], s, aug_cons=rec["aug"])
    _, W2, X2, _ = O.make_step_instance(O.FIELD_FQ, rec["t"], vdf.eval(s, rec["t"]), aug_cons=rec["aug"])
    T = shape.cross_term(W1, int(rec["u1"], 16), X1, W2, X2)
    import hashlib
    assert hashlib.sha256(O.fes_to_bytes(T, vdf.m)).hexdigest() == rec["T_sha256"]


def test_package_synthetic_generator_matches_oracle_semantics():
    """vdf_b200.synthetic (bench.py's input generator, kept outside oracle/) builds instances the oracle accepts."""
    from vdf_b200 import synthetic as S
    for fid in (O.FIELD_FP, O.FIELD_FQ):
        cons, nv, io, A, B, C, W, X = S.step_instance(fid, 12, 60)
        sh = O.R1CSShape(O.MODULUS[fid], cons, nv, io, A, B, C)
        assert cons == 3 * 12 + 1 + 60 and sh.is_sat_relaxed(W, [0] * cons, 1, X)
        vdf = O.MinRootVDF(fid)
        s = vdf.inverse_eval(O.State(W[-4 * 12 - 4], W[-4 * 12 - 3], W[-4 * 12 - 2]), 12)
        assert (W[-5], W[-2], W[-1]) == (s.x, s.y, s.i)      # last round's new_x, new_y and final_i


def test_known_dlog_numpy_helper():
    """tests/util.known_dlog_scalar (the O(n) numpy side of the 2^24 / 2^26 GPU parity tests) against Python ints."""
    import numpy as np
    from tests.util import known_dlog_scalar
    rs = np.random.RandomState(1)
    raw = rs.randint(0, 1 << 32, size=(40000, 8), dtype=np.uint64).astype(np.uint32)
    want = sum(int.from_bytes(raw[i].tobytes(), "little") * (7 + (5 + i) * 11) for i in range(raw.shape[0]))
    assert known_dlog_scalar(raw, 7, 11, first=5) == want


def test_sumcheck_oracle_is_a_sound_sumcheck():
    """Self-consistency of the oracle's sum-check restatement: every round e0 + e1 = claim with the
    cubic through (e0, e1, e2, e3), the final claim equals comb of the final evaluations, eq tables evaluate right."""
    import random
    m = O.Q
    py = random.Random(3)
    ell = 5
    tabs = [[py.randrange(m) for _ in range(1 << ell)] for _ in range(4)]
    claim = sum(a * (b * c - d) for a, b, c, d in zip(*tabs)) % m

    def interp_eval(ys, x):      # Lagrange through x = 0, 1, 2, 3
        tot = 0
        for i, yi in enumerate(ys):
            num = den = 1
            for j in range(4):
                if j != i:
                    num = num * (x - j) % m
                    den = den * (i - j) % m
            tot += yi * num * pow(den, -1, m)
        return tot % m

    state = {"claim": claim}

    def challenge(rnd, e):
        e1 = (state["claim"] - e[0]) % m
        r = py.randrange(m)
        state["claim"] = interp_eval([e[0], e1, e[1], e[2]], r)
        return r

    evals, rs, final = O.sumcheck_prove(tabs, m, challenge)
    assert len(evals) == ell and len(rs) == ell
    a, b, c, d = final
    assert state["claim"] == a * (b * c - d) % m
    assert final == [O.poly_evaluate(t, rs, m) for t in tabs]
    # eq table: eq(r, idx) and sum to one
    r = [py.randrange(m) for _ in range(4)]
    eq = O.eq_evals(r, m)
    assert sum(eq) % m == 1
    assert eq[0b1010] == r[0] * (1 - r[1]) * r[2] * (1 - r[3]) % m
    assert O.eq_evals([], m) == [1]


def test_bind_rows_is_the_transposed_product():
    """sum_y bind_rows(eq, r)[y] z[y] = sum_m r_m <eq, M z>: the identity Spartan's inner sum-check rests on."""
    import random
    py = random.Random(3)
    for fid in (O.FIELD_FP, O.FIELD_FQ):
        m = O.MODULUS[fid]
        shape, W, X, _ = O.make_step_instance(fid, 7, O.State(1, 2, 7), aug_cons=40)
        z = shape.z_of(W, 1, X)
        eq = [py.randrange(m) for _ in range(shape.num_cons)]
        r = [py.randrange(m) for _ in range(3)]
        lhs = sum(a * b for a, b in zip(shape.bind_rows(eq, r), z)) % m
        rhs = sum(rm * sum(e * v for e, v in zip(eq, Mz)) for rm, Mz in zip(r, shape.multiply_vec(z))) % m
        assert lhs == rhs
