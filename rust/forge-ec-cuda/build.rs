// Links libforge_ec_cuda.so (built by `make -C forge-ec_b200/csrc`).  FORGE_EC_CUDA_LIB_DIR points at it.
fn main() {
    let dir = std::env::var("FORGE_EC_CUDA_LIB_DIR").unwrap_or_else(|_| "../../forge-ec_b200/lib".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=forge_ec_cuda");
    println!("cargo:rerun-if-env-changed=FORGE_EC_CUDA_LIB_DIR");
}
