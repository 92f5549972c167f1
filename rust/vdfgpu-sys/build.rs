// build.rs of vdfgpu-sys.  NOT COMPILED in the build environment of this repository (no cargo).
//
// Two ways to get the native code:
//   VDFGPU_LIB_DIR=/path/to/vdf_b200/lib   link the libvdfgpu.so that `python -m vdf_b200._build` produced;
//   otherwise                              compile vdf_b200/csrc/api_{core,r1cs,sumcheck}.cu with nvcc through `cc`,
//                                          exactly the flags of vdf_b200/_build.py (sm_100a only, no fallback arch).
use std::{env, path::PathBuf};

fn main() {
    println!("cargo:rerun-if-env-changed=VDFGPU_LIB_DIR");
    if let Ok(dir) = env::var("VDFGPU_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=vdfgpu");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
        return;
    }
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("vdf_b200/csrc");
    for f in ["api_core.cu", "api_r1cs.cu", "api_sumcheck.cu"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include/vdfgpu.h").display());
    cc::Build::new()
        .cuda(true)
        .cudart("static")
        .flag("-gencode")
        .flag("arch=compute_100a,code=sm_100a")
        .flag("-O3")
        .flag("-lineinfo")
        .flag("-std=c++17")
        .flag("-Xcompiler")
        .flag("-fvisibility=hidden")
        .file(csrc.join("api_core.cu"))
        .file(csrc.join("api_r1cs.cu"))
        .file(csrc.join("api_sumcheck.cu"))
        .compile("vdfgpu");
    println!("cargo:rustc-link-lib=dylib=stdc++");
}
