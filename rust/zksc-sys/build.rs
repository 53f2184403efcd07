// Builds zk_cryptography_b200/libzksc.so with nvcc (sm_100a only) through the repository's Makefile and links it.
// ZKSC_ROOT may point at another checkout; by default the crate sits in <root>/rust/zksc-sys.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = env::var("ZKSC_ROOT")
        .map(PathBuf::from)
        .unwrap_or_else(|_| PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../.."));
    let csrc = root.join("zk_cryptography_b200/csrc");
    let status = Command::new("make").arg("-s").arg("-j8").arg("-C").arg(&csrc).status().expect("running make");
    assert!(status.success(), "nvcc build of libzksc.so failed (needs CUDA 12.9+, targets sm_100a only)");
    let libdir = root.join("zk_cryptography_b200");
    println!("cargo:rustc-link-search=native={}", libdir.display());
    println!("cargo:rustc-link-lib=dylib=zksc");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", libdir.display());
    for f in ["zksc.cu", "kernels.cuh", "resident_kernel.cuh", "res_inst.cu", "round_inst.cu", "fr.cuh", "aux_kernels.cuh", "host_field.hpp"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include/zksc.h").display());
}
