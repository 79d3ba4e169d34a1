// Link against libwrach_cuda.so.  WRACH_CUDA_LIB_DIR points at the directory holding it
// (<repo>/wrach_b200/lib after `make -C wrach_b200/csrc`); the same directory must be on
// LD_LIBRARY_PATH (or in the rpath set below) at run time.
fn main() {
    let dir = std::env::var("WRACH_CUDA_LIB_DIR").unwrap_or_else(|_| {
        let manifest = std::path::PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap());
        manifest.join("../../../wrach_b200/lib").to_string_lossy().into_owned()
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=wrach_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=WRACH_CUDA_LIB_DIR");
}
