// build.rs of the `gpu` feature of bam3d: compiles the CUDA sources for sm_100a with nvcc into a static library and
// links it together with the CUDA runtime. SOURCE ONLY in this repository — no Rust toolchain exists in the build image
// (see INTEGRATION.md); the same four translation units are what bam3d_b200/build.py compiles.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../bam3d_b200/csrc");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let mut objs = Vec::new();
    for unit in ["narrow.cu", "broad.cu", "capi.cu", "comm.cu"] {
        let obj = out.join(unit.replace(".cu", ".o"));
        let ok = Command::new(&nvcc)
            .args(["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"])
            // numerics contract: no FMA contraction, IEEE div / sqrt (bit parity with glam's scalar f32)
            .args(["-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false", "-Xcompiler", "-fPIC"])
            .arg("-c").arg(csrc.join(unit)).arg("-o").arg(&obj)
            .status().expect("nvcc not found").success();
        assert!(ok, "nvcc failed on {unit}");
        objs.push(obj);
        println!("cargo:rerun-if-changed={}", csrc.join(unit).display());
    }
    let lib = out.join("libbam3d_gpu.a");
    assert!(Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=bam3d_gpu");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=static=cudart_static");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");
    println!("cargo:rustc-link-lib=dylib=rt");
}
