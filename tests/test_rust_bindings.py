"""CPU: the Rust side (rust/, uncompiled in this environment) stays in step with the C header: every entry point
include/vdfgpu.h declares is declared in rust/vdfgpu-sys/src/lib.rs with the same number of parameters, the glue
module only calls symbols that exist, and the patched pasta-msm build script links the library."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "vdfgpu.h"
SYS = ROOT / "rust" / "vdfgpu-sys" / "src" / "lib.rs"


def _c_decls():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    out = {}
    for m in re.finditer(r"\b((?:vdfgpu_|mult_pippenger_)\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def _rust_decls():
    text = re.sub(r"//.*", "", SYS.read_text())
    out = {}
    for m in re.finditer(r"pub fn ((?:vdfgpu_|mult_pippenger_)\w+)\s*\(([^)]*)\)", text, flags=re.S):
        args = m.group(2).strip().rstrip(",")
        out[m.group(1)] = 0 if not args else args.count(",") + 1
    return out


def test_sys_crate_declares_every_header_symbol_with_the_same_arity():
    c, r = _c_decls(), _rust_decls()
    assert len(c) >= 45
    assert sorted(c) == sorted(r), (sorted(set(c) - set(r)), sorted(set(r) - set(c)))
    assert {k: v for k, v in c.items() if r[k] != v} == {}


def test_glue_module_uses_only_declared_symbols():
    r = _rust_decls()
    glue = (ROOT / "rust" / "src_nova_gpu.rs").read_text()
    used = set(re.findall(r"sys::((?:vdfgpu_|mult_pippenger_)\w+)\s*\(", glue))          # calls
    assert used and used <= set(r), used - set(r)
    types = set(re.findall(r"sys::(vdfgpu_\w+)\b(?!\s*\()", glue))                      # opaque handle types
    assert types <= set(re.findall(r"pub struct (vdfgpu_\w+)", SYS.read_text())), types
    # the constants it passes exist too
    consts = set(re.findall(r"sys::(VDFGPU_\w+)", glue))
    declared = set(re.findall(r"pub const (VDFGPU_\w+)", SYS.read_text()))
    assert consts <= declared, consts - declared


def test_pasta_msm_patch_links_the_library_and_symbols_match_the_wrapper():
    b = (ROOT / "rust" / "patches" / "pasta-msm-build.rs").read_text()
    assert "rustc-link-lib=dylib=vdfgpu" in b
    c = _c_decls()
    assert c["mult_pippenger_pallas"] == 5 and c["mult_pippenger_vesta"] == 5   # (out, points, npoints, scalars, is_mont)
    build = (ROOT / "rust" / "vdfgpu-sys" / "build.rs").read_text()
    assert "arch=compute_100a,code=sm_100a" in build and "api_core.cu" in build and "api_r1cs.cu" in build and "api_sumcheck.cu" in build
