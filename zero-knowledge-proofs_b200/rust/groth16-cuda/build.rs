// Link against libg16cuda.so.  G16_CUDA_LIB_DIR points at zero-knowledge-proofs_b200/lib.
fn main() {
    let dir = std::env::var("G16_CUDA_LIB_DIR").unwrap_or_else(|_| "../../lib".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=g16cuda");
    println!("cargo:rerun-if-env-changed=G16_CUDA_LIB_DIR");
}
