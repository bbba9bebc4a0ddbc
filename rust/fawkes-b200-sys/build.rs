// Links libfawkes_b200.so (built by `make -C fawkes-crypto_b200/csrc`).  FAWKES_B200_LIB_DIR names the directory
// that holds it; without the variable the in-tree location is used.
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("FAWKES_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../fawkes-crypto_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=fawkes_b200");
    println!("cargo:rerun-if-env-changed=FAWKES_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/fawkes_b200.h");
}
