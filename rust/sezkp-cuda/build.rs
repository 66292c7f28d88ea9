// Link against libsezkp_cuda.so; SEZKP_CUDA_LIB_DIR points at the directory that holds it
// (streaming-zero-knowledge-proofs_b200/ of this repository after `python __graft_entry__.py`).
fn main() {
    let dir = std::env::var("SEZKP_CUDA_LIB_DIR").unwrap_or_else(|_| "/usr/local/lib".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=sezkp_cuda");
    println!("cargo:rerun-if-env-changed=SEZKP_CUDA_LIB_DIR");
}
