"""GPU parity for hot path (4): batched MinRoot check (src/minroot.rs:338-371, :424-438), written to
read like the reference's own tests (minroot.rs:449-542) with the GPU in place of the CPU `check`."""
import pytest

from oracle import pasta as O
from vdf_b200 import minroot as M

pytestmark = pytest.mark.gpu

VDFS = [(M.PallasVDF, O.PallasVDF), (M.VestaVDF, O.VestaVDF)]


def test_exponents():  # minroot.rs:449-458
    for V, _ in VDFS:
        assert V.inverse_exponent() == 5


@pytest.mark.parametrize("V,OV", VDFS)
def test_eval(gpu_lib, V, OV):  # minroot.rs:479-510 (both fields here; the reference runs Pallas only)
    rng = O.XorShiftRng()
    vdf, ovdf = V(), OV()
    t = 10
    for _ in range(10):
        x = M.State(O.field_random(rng, vdf.m), O.field_random(rng, vdf.m), 0)
        result = vdf.eval(x, t)
        assert result == M.State(*vars(ovdf.eval(O.State(x.x, x.y, x.i), t)).values())
        again = vdf.inverse_eval(result, t)          # GPU
        assert x == again
        assert vdf.check(result, t, x)               # GPU
        assert not vdf.check(result, t, M.State(x.x, x.y, 1))
        assert not vdf.check(result, t + 1, x)


@pytest.mark.parametrize("V,OV", VDFS)
def test_vanilla_proof(gpu_lib, V, OV):  # minroot.rs:512-542
    rng = O.XorShiftRng()
    vdf = V()
    x = M.State(O.field_random(rng, vdf.m), 0, 0)
    t, n = 4, 3
    _z0, first = M.Evaluation.eval(vdf, x, t)
    final = first
    for _ in range(1, n):
        _, new = M.Evaluation.eval(vdf, final.result, t)
        final = final.append(new)
        assert final is not None, "failed to append proof"
    assert vdf.element(final.t) == final.result.i
    assert n * t == final.t
    assert final.verify(x)


@pytest.mark.parametrize("V,OV", VDFS)
def test_check_batch_against_oracle(gpu_lib, V, OV):
    """2^12 independent chains with ragged t, 1 % corrupted, vs the oracle's check."""
    rng = O.XorShiftRng()
    vdf, ovdf = V(), OV()
    n = 4096
    results, originals, ts = [], [], []
    for k in range(n):
        r = O.State(O.field_random(rng, vdf.m), O.field_random(rng, vdf.m), O.field_random(rng, vdf.m))
        t = [0, 1, 2, 7, 10, 33][k % 6]
        o = ovdf.inverse_eval(r, t)
        if k % 100 == 17:
            o = O.State(o.x, (o.y + 1) % vdf.m, o.i)
        results.append(M.State(r.x, r.y, r.i))
        originals.append(M.State(o.x, o.y, o.i))
        ts.append(t)
    got = vdf.check_batch(results, ts, originals)
    want = [ovdf.check(O.State(r.x, r.y, r.i), t, O.State(o.x, o.y, o.i)) for r, t, o in zip(results, ts, originals)]
    assert got == want
    assert got.count(False) == len([k for k in range(n) if k % 100 == 17])
    assert vdf.check_batch([], 5, []) == []


def test_check_batch_full_size_property(gpu_lib):
    """BASELINE config 4 size: 2^16 chains, t = 1000.  Size-independent property: chains built by the GPU's
    own inverse_eval verify, every corrupted one is rejected; a 64-chain sample is compared with the oracle."""
    rng = O.XorShiftRng()
    vdf, ovdf = M.PallasVDF(), O.PallasVDF()
    n, t = 1 << 16, 1000
    base = [M.State(O.field_random(rng, vdf.m), O.field_random(rng, vdf.m), (k + t) % vdf.m) for k in range(256)]
    results = [base[k % 256] if k < 256 else M.State((base[k % 256].x + k) % vdf.m, base[k % 256].y, base[k % 256].i)
               for k in range(n)]
    originals = vdf.inverse_eval_batch(results, t)
    for k in range(0, n, 1024):
        o = ovdf.inverse_eval(O.State(results[k].x, results[k].y, results[k].i), t)
        assert originals[k] == M.State(o.x, o.y, o.i)
    bad = set(range(5, n, 100))
    originals = [M.State(o.x, o.y, (o.i + 1) % vdf.m) if k in bad else o for k, o in enumerate(originals)]
    ok = vdf.check_batch(results, t, originals)
    assert all(ok[k] == (k not in bad) for k in range(n))


@pytest.mark.parametrize("V,OV,fid", [(M.PallasVDF, O.PallasVDF, O.FIELD_FQ), (M.VestaVDF, O.VestaVDF, O.FIELD_FP)])
def test_step_witness_batch(gpu_lib, V, OV, fid):
    """the step part of W generated on the device equals the oracle's synthesis of
    InverseMinRootCircuit (src/nova/proof.rs:87-230) and satisfies the step shape."""
    vdf, ovdf = V(), OV()
    rng = O.XorShiftRng()
    t, n = 25, 7
    states = [ovdf.eval(O.State(O.field_random(rng, vdf.m), 0, 1), t) for _ in range(n)]
    got = vdf.step_witness_batch([M.State(s.x, s.y, s.i) for s in states], t)
    for s, w in zip(states, got):
        shape, W, X, outs = O.make_step_instance(fid, t, s)
        assert w == W[3:]
        assert shape.is_sat_relaxed([s.x, s.y, s.i] + w, [0] * shape.num_cons, 1, X)
    assert vdf.step_witness_batch([], t) == []
