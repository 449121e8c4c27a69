// Links the prebuilt sm_100a library.  TFHE_AES_CUDA_LIB_DIR points at the directory that holds libtfhe_aes_cuda.so
// (tfhe-aes-2_b200/csrc after `python -c 'import __graft_entry__ as g; g.build()'`).
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("TFHE_AES_CUDA_LIB_DIR")
        .map(PathBuf::from)
        .unwrap_or_else(|_| PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../tfhe-aes-2_b200/csrc"));
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=tfhe_aes_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=TFHE_AES_CUDA_LIB_DIR");
}
