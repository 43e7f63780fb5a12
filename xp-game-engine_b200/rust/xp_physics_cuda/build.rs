// UNVERIFIED (no Rust toolchain in the build image).
// Compiles the CUDA sources for sm_100a with nvcc and links the resulting shared library.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = manifest.join("../../csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let so = out.join("libxp_physics_cuda.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let status = Command::new(nvcc)
        .args(&["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                "--prec-div=true", "--prec-sqrt=true", "--ftz=false", "-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&so)
        .args(["xp_api.cu", "xp_build.cu", "xp_query.cu"].iter().map(|f| csrc.join(f)))
        .status()
        .expect("nvcc not found");
    assert!(status.success(), "nvcc failed");
    for f in &["xp_api.cu", "xp_build.cu", "xp_query.cu", "xp_narrowphase.cuh", "xp_device.cuh", "xp_internal.h"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=xp_physics_cuda");
}
