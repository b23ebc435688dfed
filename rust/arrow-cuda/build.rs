// Links the C-ABI product library built by `python -c "import __graft_entry__ as g; g.build()"`.
fn main() {
    let dir = std::env::var("ARROW_CUDA_LIB_DIR").unwrap_or_else(|_| "../../arrow-rs_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=arrow_cuda");
    println!("cargo:rerun-if-env-changed=ARROW_CUDA_LIB_DIR");
}
