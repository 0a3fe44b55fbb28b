// Links libazb.so.  AZB_LIB_DIR names the directory that holds it (azdopt_b200/lib in this repository).
fn main() {
    println!("cargo:rerun-if-env-changed=AZB_LIB_DIR");
    if let Ok(dir) = std::env::var("AZB_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=azb");
}
