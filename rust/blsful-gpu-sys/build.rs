// Links libblsgpu.so (built by `python __graft_entry__.py` into agora-blsful_b200/).  BLSGPU_LIB_DIR names its directory.
use std::env;

fn main() {
    println!("cargo:rerun-if-env-changed=BLSGPU_LIB_DIR");
    if let Ok(dir) = env::var("BLSGPU_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=blsgpu");
}
