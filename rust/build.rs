// build.rs -- links the reference against libslime_b200.so (built by `python -m slime_mold_b200.build`, nvcc, sm_100a).
// No bindgen / cc: the ABI (include/slime_b200.h) is 33 plain extern "C" functions, declared by hand in src/ffi.rs.
fn main() {
    let dir = std::env::var("SLIME_B200_LIB_DIR")
        .expect("set SLIME_B200_LIB_DIR to the directory that holds libslime_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=slime_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=SLIME_B200_LIB_DIR");
}
