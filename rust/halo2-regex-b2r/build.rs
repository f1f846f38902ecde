// Links against the prebuilt shared library of this repository (halo2_regex_b200/libb2r.so; `make -C halo2_regex_b200/csrc`).
fn main() {
    let dir = std::env::var("B2R_LIB_DIR").expect("set B2R_LIB_DIR to the directory that holds libb2r.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=b2r");
    println!("cargo:rerun-if-env-changed=B2R_LIB_DIR");
}
