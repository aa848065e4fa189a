// Link against the in-tree shared library: TECDSA_B200_LIB_DIR=<repo>/multi-party-ecdsa_b200 (where __graft_entry__.build() puts it).
fn main() {
    let dir = std::env::var("TECDSA_B200_LIB_DIR").unwrap_or_else(|_| "../../multi-party-ecdsa_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=tecdsa_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=TECDSA_B200_LIB_DIR");
}
