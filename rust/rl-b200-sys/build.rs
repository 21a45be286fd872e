//! Compiles the hand-written CUDA sources for sm_100a and links them.  No cc / cmake crates: one nvcc call,
//! the same command `__graft_entry__.build()` runs.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = manifest.join("../../rendering_learning_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let sources = ["api.cu", "lbvh.cu", "rtc_kernels.cu", "ow_kernels.cu", "peaks.cu", "flatten.cpp"];
    let lib = out.join("librl_b200.so");
    let mut cmd = Command::new(&nvcc);
    cmd.args(["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"])
        .args(["-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&lib);
    for s in sources {
        cmd.arg(csrc.join(s));
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
    }
    for h in ["scene.h", "device.cuh", "kernels.h", "lbvh.h", "../../include/rl_b200.h"] {
        println!("cargo:rerun-if-changed={}", csrc.join(h).display());
    }
    let status = cmd.status().expect("nvcc not found: there is no CPU fallback for this crate");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=rl_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
}
