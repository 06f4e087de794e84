// Links libswirl_b200.so (built by `make -C stark-backend_b200`); SWIRL_B200_LIB_DIR points at the directory holding it.
fn main() {
    if let Ok(dir) = std::env::var("SWIRL_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
    }
    println!("cargo:rustc-link-lib=dylib=swirl_b200");
    println!("cargo:rerun-if-env-changed=SWIRL_B200_LIB_DIR");
}
