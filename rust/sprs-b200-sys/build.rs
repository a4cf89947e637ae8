// Same shape as suitesparse_umfpack_sys/build.rs:1-4: link an already-built library.
fn main() {
    let dir = std::env::var("SPRS_B200_LIB_DIR").unwrap_or_else(|_| "../../sprs_b200".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=sprs_b200");
    println!("cargo:rerun-if-env-changed=SPRS_B200_LIB_DIR");
}
