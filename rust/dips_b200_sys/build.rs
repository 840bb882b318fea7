// Links against the prebuilt libdips_b200.so (built by `python -m dips_b200._build` / nvcc).
// DIPS_B200_LIB_DIR points at the directory holding it (default: ../../dips_b200 relative to this crate).
fn main() {
    let dir = std::env::var("DIPS_B200_LIB_DIR").unwrap_or_else(|_| {
        let manifest = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{}/../../dips_b200", manifest)
    });
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=dips_b200");
    println!("cargo:rerun-if-env-changed=DIPS_B200_LIB_DIR");
}
