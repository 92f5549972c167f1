"""The C++ host layer (include/vdf_host.hpp, reference-shaped names over the C ABI): compiled with g++ and linked
against libvdfgpu.so.  CPU: it builds and fails loudly without a GPU.  GPU: it replays oracle fixtures bit-exactly."""
import struct
import subprocess
from pathlib import Path

import pytest

from oracle import pasta as O

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "tests" / "cpp" / "host_check"


def _build():
    from vdf_b200 import _lib
    _lib.load()
    src = ROOT / "tests" / "cpp" / "host_check.cpp"
    deps = [src, ROOT / "include" / "vdf_host.hpp", ROOT / "include" / "vdfgpu.h"]
    if not EXE.exists() or any(d.stat().st_mtime > EXE.stat().st_mtime for d in deps):
        libdir = ROOT / "vdf_b200" / "lib"
        subprocess.run(["g++", "-O1", "-std=c++17", "-o", str(EXE), str(src), f"-L{libdir}", "-lvdfgpu",
                        f"-Wl,-rpath,{libdir}"], check=True)
    return EXE


def _fixtures(d: Path):
    m = O.Q
    vdf = O.PallasVDF()
    rng = O.XorShiftRng()
    res, orig, ts, ok = [], [], [], []
    for k in range(300):
        r = O.State(*[O.field_random(rng, m) for _ in range(3)])
        t = [0, 1, 3, 9][k % 4]
        o = vdf.inverse_eval(r, t)
        if k % 11 == 5:
            o = O.State(o.x, (o.y + 1) % m, o.i)
        res.append(r); orig.append(o); ts.append(t); ok.append(vdf.check(r, t, o))
    (d / "mr_results.bin").write_bytes(b"".join(O.state_to_bytes(s, m) for s in res))
    (d / "mr_originals.bin").write_bytes(b"".join(O.state_to_bytes(s, m) for s in orig))
    (d / "mr_t.bin").write_bytes(struct.pack("<%dQ" % len(ts), *ts))
    (d / "mr_ok.bin").write_bytes(bytes(int(b) for b in ok))
    (d / "mr_inverse6.bin").write_bytes(b"".join(O.state_to_bytes(vdf.inverse_eval(s, 6), m) for s in res))
    s0 = O.State(O.field_random(rng, m), 0, 0)
    s1 = vdf.eval(s0, 4)
    s2 = vdf.eval(s1, 4)
    (d / "mr_segments.bin").write_bytes(b"".join(O.state_to_bytes(s, m) for s in (s0, s1, s2)))
    cv = O.PALLAS
    n, k0, dd = 400, 21, 5
    sc = [O.field_random(rng, cv.order) for _ in range(n)]
    sc[0], sc[1] = 0, cv.order - 1
    (d / "msm_k0_d.bin").write_bytes(k0.to_bytes(32, "little") + dd.to_bytes(32, "little"))
    (d / "msm_scalars.bin").write_bytes(O.fes_to_bytes(sc, cv.order))
    (d / "msm_points.bin").write_bytes(O.affines_to_bytes(cv, cv.progression(k0, dd, n)))
    (d / "msm_point.bin").write_bytes(O.jac_to_bytes(cv, cv.msm_known_dlog(sc, k0, dd)))
    t, aug = 7, 40
    s = vdf.eval(O.State(O.field_random(rng, m), 0, 1), t)
    shape, W1, X1, _ = O.make_step_instance(O.FIELD_FQ, t, s, aug_cons=aug)
    _, W2, X2, _ = O.make_step_instance(O.FIELD_FQ, t, vdf.eval(s, t), aug_cons=aug)
    coo = O.shape_to_coo_bytes(shape)
    (d / "r1cs_dims.bin").write_bytes(struct.pack("<6Q", shape.num_cons, shape.num_vars, shape.num_io, *[c[3] for c in coo]))
    for name, (rows, cols, vals, _) in zip("abc", coo):
        (d / f"r1cs_{name}_rows.bin").write_bytes(rows)
        (d / f"r1cs_{name}_cols.bin").write_bytes(cols)
        (d / f"r1cs_{name}_vals.bin").write_bytes(vals)
    u1, r = 1, O.field_random(rng, m) >> 127   # running instance fresh enough for E1 = 0 to be consistent
    for name, v in (("W1", W1), ("W2", W2), ("X1", X1), ("X2", X2), ("u1", [u1]), ("r", [r])):
        (d / f"r1cs_{name}.bin").write_bytes(O.fes_to_bytes(v, m))
    Az, Bz, Cz = shape.multiply_vec(shape.z_of(W1, u1, X1))
    (d / "r1cs_ABC.bin").write_bytes(O.fes_to_bytes(Az + Bz + Cz, m))
    T = shape.cross_term(W1, u1, X1, W2, X2)
    (d / "r1cs_T.bin").write_bytes(O.fes_to_bytes(T, m))
    (d / "r1cs_commT.bin").write_bytes(O.jac_to_bytes(cv, cv.msm_known_dlog(T, k0, dd)))
    (d / "r1cs_Wfold.bin").write_bytes(O.fes_to_bytes(O.fold_vec(W1, W2, r, m), m))
    (d / "r1cs_Efold.bin").write_bytes(O.fes_to_bytes(O.fold_vec([0] * shape.num_cons, T, r, m), m))


def test_cpp_host_builds_and_has_no_cpu_fallback(tmp_path):
    import torch
    exe = _build()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    _fixtures(tmp_path)
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_host_parity(tmp_path):
    exe = _build()
    _fixtures(tmp_path)
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout + r.stderr
