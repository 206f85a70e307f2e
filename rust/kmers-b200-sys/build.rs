// UNVERIFIED SOURCE (no Rust toolchain in the build image).
// Compiles the translation units (kmers_b200/csrc/*.cu for sm_100a, *.cpp = the host packer, by nvcc's host compiler) into
// one shared library and links it.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let mut srcs: Vec<PathBuf> = std::fs::read_dir(root.join("kmers_b200/csrc")).unwrap()
        .map(|e| e.unwrap().path()).filter(|p| p.extension().map_or(false, |x| x == "cu" || x == "cpp")).collect();
    srcs.sort();
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libkmers_b200.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let status = Command::new(nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&lib)
        .args(&srcs)
        .args(["-ldl", "-lpthread"])
        .status()
        .expect("nvcc not runnable");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=kmers_b200");
    println!("cargo:rerun-if-changed={}", root.join("kmers_b200/csrc").display());
    println!("cargo:rerun-if-changed={}", root.join("include/kmers_b200.h").display());
}
