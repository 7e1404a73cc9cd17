// Link against the in-tree library: MANTAPROVER_LIB_DIR=<repo>/manta-rs_b200 (the library links the CUDA runtime statically).
fn main() {
    let dir = std::env::var("MANTAPROVER_LIB_DIR").unwrap_or_else(|_| "../../../manta-rs_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=mantaprover");
    println!("cargo:rerun-if-env-changed=MANTAPROVER_LIB_DIR");
}
