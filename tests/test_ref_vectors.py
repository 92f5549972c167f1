"""Consumer of tests/golden/ref_vectors.json -- golden vectors dumped from the REFERENCE stack (protocol/vdf +
pasta_curves 0.4 + pasta-msm 0.1 [+ nova-snark 0.8]) by rust/golden (see rust/README.md).

No Rust toolchain exists in this build environment, so the file is absent today and parity stays "unpinned"
(DESIGN.md section 5): the parity tests below then SKIP with that reason.  The day the file is committed they pin,
with no further code, (CPU) the oracle and (GPU, `-m gpu`) the CUDA path against the reference's own bytes:
field encodings and arithmetic, point encodings, pasta-msm's MSM result, MinRoot chains (test_eval's inputs,
src/minroot.rs:497-516) and -- with the generator's `r1cs` feature -- multiply_vec / commit_T / fold on the real
step-circuit shape.

`test_consumer_plumbing_on_oracle_document` runs the same checks on a document of the same format produced by the
oracle itself: it proves the consumer works, not parity."""
import json
from pathlib import Path

import pytest

from oracle import pasta as O

REF = Path(__file__).resolve().parent / "golden" / "ref_vectors.json"
FIELDS = {"fp": O.P, "fq": O.Q}
CURVES = {"pallas": O.PALLAS, "vesta": O.VESTA}
SKIP = "tests/golden/ref_vectors.json absent: no Rust toolchain here to run rust/golden (parity unpinned, DESIGN.md section 5)"


def _h(b: bytes) -> str:
    return b.hex()


def oracle_document() -> dict:
    """Same structure as rust/golden/src/main.rs writes, values from the oracle (plumbing check only)."""
    rng = O.XorShiftRng()
    doc = {"generator": "oracle (plumbing check)", "fields": {}, "curves": {}, "minroot": {}}
    for name, m in FIELDS.items():
        elems = [{"canonical_le": _h(v.to_bytes(32, "little")), "mont": _h(O.fe_to_bytes(v, m))}
                 for v in (0, 1, m - 1, 2, (1 << 64) - 1)]
        ops = []
        for _ in range(4):
            a, b = O.field_random(rng, m), O.field_random(rng, m)
            f = lambda v: _h(O.fe_to_bytes(v % m, m))  # noqa: E731
            ops.append({"a": f(a), "b": f(b), "add": f(a + b), "sub": f(a - b), "mul": f(a * b), "square": f(a * a),
                        "neg": f(-a), "invert": f(pow(a, -1, m))})
        doc["fields"][name] = {"elements": elems, "ops": ops}
    for name, cv in CURVES.items():
        n = 40
        pts = [cv.mul(O.field_random(rng, cv.order), cv.gen) for _ in range(n)]
        sc = [O.field_random(rng, cv.order) for _ in range(n)]
        sc[0], sc[1], sc[2] = 0, 1, cv.order - 1
        doc["curves"][name] = {
            "generator_affine72": _h(O.affine_to_bytes(cv, cv.gen)),
            "multiples": [{"k": k, "affine72": _h(O.affine_to_bytes(cv, cv.mul(k, cv.gen)))} for k in (0, 1, 2, 3, 0xDEADBEEF)],
            "msm": {"n": n, "points_affine72": _h(O.affines_to_bytes(cv, pts)), "scalars_mont": _h(O.fes_to_bytes(sc, cv.order)),
                    "result_affine72": _h(O.affine_to_bytes(cv, cv.msm_naive(sc, pts)))}}
    for name, fid in (("pallas", O.FIELD_FQ), ("vesta", O.FIELD_FP)):
        vdf = O.MinRootVDF(fid)
        chains = []
        for _ in range(3):
            x = O.State(O.field_random(rng, vdf.m), O.field_random(rng, vdf.m), 0)
            chains.append({"t": 10, "original": _h(O.state_to_bytes(x, vdf.m)), "result": _h(O.state_to_bytes(vdf.eval(x, 10), vdf.m))})
        doc["minroot"][name] = {"chains": chains}
    # r1cs section on the step-circuit shape (t = 5), as the generator's `r1cs` feature writes it
    fid, cv = O.FIELD_FQ, O.PALLAS
    vdf = O.MinRootVDF(fid)
    s1 = vdf.eval(O.State(O.field_random(rng, vdf.m), 0, 0), 5)
    shape, W1, X1, _ = O.make_step_instance(fid, 5, s1)
    _, W2, X2, _ = O.make_step_instance(fid, 5, vdf.eval(s1, 5))
    m = shape.m
    gens = cv.progression(3, 5, max(shape.num_cons, shape.num_vars))
    Az, Bz, Cz = shape.multiply_vec(shape.z_of(W1, 1, X1))
    T = shape.cross_term(W1, 1, X1, W2, X2)
    r = 0x123456789ABCDEF0
    coo = lambda M: [[a, b, _h(O.fe_to_bytes(v, m))] for a, b, v in M]  # noqa: E731
    doc["r1cs"] = {"t": 5, "num_cons": shape.num_cons, "num_vars": shape.num_vars, "num_io": shape.num_io,
                   "A": coo(shape.A), "B": coo(shape.B), "C": coo(shape.C), "gens_affine72": _h(O.affines_to_bytes(cv, gens)),
                   "W1": _h(O.fes_to_bytes(W1, m)), "X1": _h(O.fes_to_bytes(X1, m)), "W2": _h(O.fes_to_bytes(W2, m)),
                   "X2": _h(O.fes_to_bytes(X2, m)), "Az1": _h(O.fes_to_bytes(Az, m)), "Bz1": _h(O.fes_to_bytes(Bz, m)),
                   "Cz1": _h(O.fes_to_bytes(Cz, m)), "T": _h(O.fes_to_bytes(T, m)),
                   "comm_W1_affine72": _h(O.affine_to_bytes(cv, cv.msm(W1, gens[:len(W1)]))),
                   "comm_T_affine72": _h(O.affine_to_bytes(cv, cv.msm(T, gens[:len(T)]))),
                   "r": _h(O.fe_to_bytes(r, m)), "W_folded": _h(O.fes_to_bytes(O.fold_vec(W1, W2, r, m), m)),
                   "E_folded": _h(O.fes_to_bytes(O.fold_vec([0] * shape.num_cons, T, r, m), m))}
    return doc


# ---- the checks: `doc` is the reference's document ---------------------------------------------------------
def check_oracle_against(doc: dict) -> None:
    for name, m in FIELDS.items():
        sec = doc["fields"][name]
        for e in sec["elements"]:
            v = int.from_bytes(bytes.fromhex(e["canonical_le"]), "little")
            assert O.fe_to_bytes(v, m) == bytes.fromhex(e["mont"])           # Montgomery R = 2^256, 4 x u64 LE
        for o in sec["ops"]:
            a, b = (O.fe_from_bytes(bytes.fromhex(o[k]), m) for k in ("a", "b"))
            want = {"add": a + b, "sub": a - b, "mul": a * b, "square": a * a, "neg": -a, "invert": pow(a, -1, m)}
            for k, v in want.items():
                assert O.fe_to_bytes(v % m, m) == bytes.fromhex(o[k]), (name, k)
    for name, cv in CURVES.items():
        sec = doc["curves"][name]
        assert O.affine_to_bytes(cv, cv.gen)[:65] == bytes.fromhex(sec["generator_affine72"])[:65]   # G = (-1, 2)
        for p in sec["multiples"]:
            assert O.affine_to_bytes(cv, cv.mul(p["k"], cv.gen))[:65] == bytes.fromhex(p["affine72"])[:65]
        msm = sec["msm"]
        pb, sb = bytes.fromhex(msm["points_affine72"]), bytes.fromhex(msm["scalars_mont"])
        pts = [O.affine_from_bytes(cv, pb[72 * i:72 * i + 72]) for i in range(msm["n"])]
        sc = O.fes_from_bytes(sb, cv.order)
        assert O.affine_to_bytes(cv, cv.msm(sc, pts))[:65] == bytes.fromhex(msm["result_affine72"])[:65]
    for name, fid in (("pallas", O.FIELD_FQ), ("vesta", O.FIELD_FP)):
        vdf = O.MinRootVDF(fid)
        for c in doc["minroot"][name]["chains"]:
            ob, rb = bytes.fromhex(c["original"]), bytes.fromhex(c["result"])
            orig = O.State(*O.fes_from_bytes(ob, vdf.m))
            res = O.State(*O.fes_from_bytes(rb, vdf.m))
            assert vdf.eval(orig, c["t"]) == res and vdf.check(res, c["t"], orig)
    if "r1cs" in doc:
        r, cv = doc["r1cs"], O.PALLAS
        m = O.Q
        dec = lambda k: O.fes_from_bytes(bytes.fromhex(r[k]), m)  # noqa: E731
        mats = [[(a, b, O.fe_from_bytes(bytes.fromhex(v), m)) for a, b, v in r[k]] for k in "ABC"]
        shape = O.R1CSShape(m, r["num_cons"], r["num_vars"], r["num_io"], *mats)
        # the oracle's own restatement of InverseMinRootCircuit::synthesize must give the reference's shape
        want_shape, _, _, _ = O.make_step_instance(O.FIELD_FQ, r["t"], O.State(1, 2, 3))
        assert (shape.num_cons, shape.num_vars) == (want_shape.num_cons, want_shape.num_vars)
        assert [sorted(M) for M in (shape.A, shape.B, shape.C)] == [sorted(M) for M in (want_shape.A, want_shape.B, want_shape.C)]
        W1, X1, W2, X2 = dec("W1"), dec("X1"), dec("W2"), dec("X2")
        assert list(shape.multiply_vec(shape.z_of(W1, 1, X1))) == [dec("Az1"), dec("Bz1"), dec("Cz1")]
        T = shape.cross_term(W1, 1, X1, W2, X2)
        assert T == dec("T")
        rr = dec("r")[0]
        assert O.fold_vec(W1, W2, rr, m) == dec("W_folded") and O.fold_vec([0] * shape.num_cons, T, rr, m) == dec("E_folded")
        gb = bytes.fromhex(r["gens_affine72"])
        gens = [O.affine_from_bytes(cv, gb[72 * i:72 * i + 72]) for i in range(len(gb) // 72)]
        assert O.affine_to_bytes(cv, cv.msm(T, gens[:len(T)]))[:65] == bytes.fromhex(r["comm_T_affine72"])[:65]
        assert O.affine_to_bytes(cv, cv.msm(W1, gens[:len(W1)]))[:65] == bytes.fromhex(r["comm_W1_affine72"])[:65]


def check_gpu_against(doc: dict) -> None:
    from vdf_b200 import _lib, minroot as M, msm as G, nova as N
    lib = _lib.load()
    _lib.check(lib.vdfgpu_init(0))
    for name, fid in (("fp", 0), ("fq", 1)):
        ops = doc["fields"][name]["ops"]
        a = b"".join(bytes.fromhex(o["a"]) for o in ops)
        b = b"".join(bytes.fromhex(o["b"]) for o in ops)
        out = bytearray(len(a))
        _lib.check(lib.vdfgpu_field_mul_batch(fid, _lib.as_ptr(a), _lib.as_ptr(b), len(ops), 1, _lib.as_ptr(out)))
        assert bytes(out) == b"".join(bytes.fromhex(o["mul"]) for o in ops)
        _lib.check(lib.vdfgpu_field_mul_batch(fid, _lib.as_ptr(a), _lib.as_ptr(b), len(ops), 1 | 0x40000000, _lib.as_ptr(out)))
        assert bytes(out) == b"".join(bytes.fromhex(o["square"]) for o in ops)
    for name, cv in CURVES.items():
        msm = doc["curves"][name]["msm"]
        pb, sb = bytes.fromhex(msm["points_affine72"]), bytes.fromhex(msm["scalars_mont"])
        want = O.affine_from_bytes(cv, bytes.fromhex(msm["result_affine72"]))
        assert O.jac_from_bytes(cv, G.mult_pippenger(cv.cid, pb, sb, True)) == want          # the literal pasta-msm symbol
        for table in (False, True):
            g = G.Generators.from_affine_bytes(cv.cid, pb, table=table)
            assert O.jac_from_bytes(cv, g.commit_bytes(sb)) == want
            g.close()
    for name, fid in (("pallas", 1), ("vesta", 0)):
        chains = doc["minroot"][name]["chains"]
        res = b"".join(bytes.fromhex(c["result"]) for c in chains)
        orig = b"".join(bytes.fromhex(c["original"]) for c in chains)
        ok = bytearray(len(chains))
        _lib.check(lib.vdfgpu_minroot_check_batch(fid, _lib.as_ptr(res), _lib.as_ptr(orig), None, chains[0]["t"], len(chains), _lib.as_ptr(ok)))
        assert list(ok) == [1] * len(chains)
        back = bytearray(len(res))
        _lib.check(lib.vdfgpu_minroot_inverse_eval_batch(fid, _lib.as_ptr(res), chains[0]["t"], len(chains), _lib.as_ptr(back)))
        assert bytes(back) == orig
    if "r1cs" in doc:
        r, cv, m = doc["r1cs"], O.PALLAS, O.Q
        dec = lambda k: O.fes_from_bytes(bytes.fromhex(r[k]), m)  # noqa: E731
        mats = [[(a, b, O.fe_from_bytes(bytes.fromhex(v), m)) for a, b, v in r[k]] for k in "ABC"]
        gs = N.R1CSShape(O.FIELD_FQ, r["num_cons"], r["num_vars"], r["num_io"], *mats)
        W1, X1, W2, X2 = dec("W1"), dec("X1"), dec("W2"), dec("X2")
        assert list(gs.multiply_vec(W1 + [1] + X1)) == [dec("Az1"), dec("Bz1"), dec("Cz1")]
        gens = G.Generators.from_affine_bytes(cv.cid, bytes.fromhex(r["gens_affine72"]), table=True)
        T, cT = gs.commit_T(gens, W1, 1, X1, W2, X2)
        assert T == dec("T") and cT == O.affine_from_bytes(cv, bytes.fromhex(r["comm_T_affine72"]))
        assert gens.commit(W1) == O.affine_from_bytes(cv, bytes.fromhex(r["comm_W1_affine72"]))
        Wf, Ef = N.fold_vectors(O.FIELD_FQ, W1, W2, [0] * r["num_cons"], T, dec("r")[0])
        assert Wf == dec("W_folded") and Ef == dec("E_folded")
        gens.close(); gs.close()
    assert M is not None


def _load():
    if not REF.exists():
        pytest.skip(SKIP)
    return json.loads(REF.read_text())


def test_oracle_matches_reference_vectors():
    check_oracle_against(_load())


@pytest.mark.gpu
def test_gpu_matches_reference_vectors(gpu_lib):
    check_gpu_against(_load())


def test_consumer_plumbing_on_oracle_document():
    """NOT parity: the consumer's checks run green on a document of the generator's format built by the oracle."""
    doc = oracle_document()
    json.loads(json.dumps(doc))
    check_oracle_against(doc)


@pytest.mark.gpu
def test_gpu_consumer_plumbing_on_oracle_document(gpu_lib):
    """The CUDA path against the same oracle-built document (doubles as an encoding-level GPU parity test)."""
    check_gpu_against(oracle_document())
