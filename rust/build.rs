fn main() {
    let dir = std::env::var("FMX_LIB_DIR").unwrap_or_else(|_| "../fm-index_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=fmx_b200");
    println!("cargo:rerun-if-env-changed=FMX_LIB_DIR");
}
