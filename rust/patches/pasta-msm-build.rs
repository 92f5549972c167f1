// Replacement build.rs for a vendored pasta-msm 0.1.x.  NOT COMPILED in the build environment of this repository.
//
// pasta-msm's Rust wrapper (src/lib.rs) declares
//     extern "C" { fn mult_pippenger_pallas(out: *mut pallas::Point, points: *const pallas::Affine, npoints: usize,
//                                           scalars: *const pallas::Scalar, is_mont: bool); /* + _vesta */ }
// and its stock build.rs compiles the C++/assembly (sppark + semolina) that defines them.  libvdfgpu.so exports the
// same two symbols with the same signature (include/vdfgpu.h, section a4), so this build.rs only has to link it:
// every Group::vartime_multiscalar_mul of nova-snark -- commit(W), commit(T) inside RecursiveSNARK::prove_step
// (reference src/nova/proof.rs:342-349), the Spartan/IPA commitments of compress (:360-368) -- then runs on the
// B200 with no change to nova-snark or to the wrapper.  The library keeps the generator sets it has seen resident
// (keyed by the slice's address; see "drop-in cache" in vdf_b200/csrc/api_core.cu), so after the first call only the
// scalars cross PCIe.
use std::env;

fn main() {
    println!("cargo:rerun-if-env-changed=VDFGPU_LIB_DIR");
    let dir = env::var("VDFGPU_LIB_DIR").expect("set VDFGPU_LIB_DIR to the directory holding libvdfgpu.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=vdfgpu");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
}
