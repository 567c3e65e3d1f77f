// Builds liblzfse_b200.so from the CUDA sources with nvcc (sm_100a only) and links it.
//
//   LZFSE_B200_LIB_DIR  use a prebuilt liblzfse_b200.so from this directory instead of compiling
//   NVCC                nvcc binary (default: nvcc on PATH, then /usr/local/cuda/bin/nvcc)
//   CUDA_HOME           where libcudart lives (default /usr/local/cuda)
use std::env;
use std::path::{Path, PathBuf};
use std::process::Command;

const SOURCES: [&str; 5] = ["decode.cu", "expand.cu", "expand_long.cu", "encode.cu", "api.cu"];

fn main() {
    let cuda_home = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    println!("cargo:rerun-if-env-changed=LZFSE_B200_LIB_DIR");
    if let Ok(dir) = env::var("LZFSE_B200_LIB_DIR") {
        link(Path::new(&dir), &cuda_home);
        return;
    }
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = manifest.join("../../lzfse_rust_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| {
        if Command::new("nvcc").arg("--version").output().is_ok() { "nvcc".into() } else { format!("{}/bin/nvcc", cuda_home) }
    });
    let mut objs = Vec::new();
    for src in SOURCES.iter() {
        let s = csrc.join(src);
        println!("cargo:rerun-if-changed={}", s.display());
        let o = out.join(src.replace(".cu", ".o"));
        let st = Command::new(&nvcc)
            .args(&["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math", "-Xcompiler", "-fPIC,-O2", "-c"])
            .arg(&s).arg("-o").arg(&o)
            .status().expect("nvcc not found: the B200 path has no CPU fallback");
        assert!(st.success(), "nvcc failed on {}", src);
        objs.push(o);
    }
    for h in ["common.cuh", "lz_blocks.cuh", "encode_long.cuh", "host_util.h", "../../include/lzfse_b200.h"].iter() {
        println!("cargo:rerun-if-changed={}", csrc.join(h).display());
    }
    let so = out.join("liblzfse_b200.so");
    let st = Command::new(&nvcc).arg("-shared").arg("-o").arg(&so).args(&objs)
        .args(&["-gencode", "arch=compute_100a,code=sm_100a"]).status().expect("nvcc link");
    assert!(st.success(), "linking liblzfse_b200.so failed");
    link(&out, &cuda_home);
}

fn link(dir: &Path, cuda_home: &str) {
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=lzfse_b200");
    println!("cargo:rustc-link-search=native={}/lib64", cuda_home);
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-env=LZFSE_B200_LIB_DIR={}", dir.display());
}
