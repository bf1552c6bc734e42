// Links against the in-tree librf_b200.so built by `python -m retrofire_b200.build`.
fn main() {
    let dir = std::env::var("RF_B200_LIB_DIR").unwrap_or_else(|_| "../../retrofire_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=rf_b200");
    println!("cargo:rerun-if-env-changed=RF_B200_LIB_DIR");
}
