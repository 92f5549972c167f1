"""Generates tests/golden/vectors.json from the oracle (seed 42 inputs).  The reference itself cannot be
run or imported here (Rust, no toolchain), so these are frozen oracle outputs: they pin the
oracle and the CUDA path against drift, not against the Rust crates.  Run: python tests/golden/make_golden.py"""
import hashlib
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from oracle import pasta as O  # noqa: E402


def main():
    out = {"field_mul_mont": {}, "minroot": {}, "msm": {}}
    for fname, m in (("fp", O.P), ("fq", O.Q)):
        rng = O.XorShiftRng()
        vals = [0, 1, m - 1, 1 << 254, (1 << 128) - 1] + [O.field_random(rng, m) for _ in range(11)]
        out["field_mul_mont"][fname] = [[hex(a), hex(b), O.fe_to_bytes(a * b % m, m).hex()]
                                       for a, b in zip(vals, vals[1:] + vals[:1])]
    for name, mk in (("pallas", O.PallasVDF), ("vesta", O.VestaVDF)):
        vdf = mk()
        rng = O.XorShiftRng()
        recs = []
        for t in (1, 4, 10, 25):
            s = O.State(O.field_random(rng, vdf.m), O.field_random(rng, vdf.m), 0)
            r = vdf.eval(s, t)
            recs.append({"t": t, "start": [hex(s.x), hex(s.y), hex(s.i)], "result": [hex(r.x), hex(r.y), hex(r.i)]})
        out["minroot"][name] = recs
    for cname, cv in (("pallas", O.PALLAS), ("vesta", O.VESTA)):
        rng = O.XorShiftRng()
        n, k0, d = 64, 7, 3
        sc = [O.field_random(rng, cv.order) for _ in range(n)]
        pts = cv.progression(k0, d, n)
        out["msm"][cname] = {"n": n, "k0": k0, "d": d, "result_point96": O.jac_to_bytes(cv, cv.msm_naive(sc, pts)).hex()}
    vdf = O.PallasVDF()
    rng = O.XorShiftRng()
    t, aug, u1 = 10, 50, 0xABCDEF0123456789
    s = vdf.eval(O.State(O.field_random(rng, vdf.m), 0, 1), t)
    shape, W1, X1, _ = O.make_step_instance(O.FIELD_FQ, t, s, aug_cons=aug)
    _, W2, X2, _ = O.make_step_instance(O.FIELD_FQ, t, vdf.eval(s, t), aug_cons=aug)
    T = shape.cross_term(W1, u1, X1, W2, X2)
    out["cross_term"] = {"t": t, "aug": aug, "u1": hex(u1), "T_sha256": hashlib.sha256(O.fes_to_bytes(T, vdf.m)).hexdigest()}
    path = Path(__file__).parent / "vectors.json"
    path.write_text(json.dumps(out, indent=1))
    print("wrote", path)


if __name__ == "__main__":
    main()
