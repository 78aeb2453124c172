// Points the linker at libbls12381_b200.so: set BLS12381_B200_LIB_DIR to the directory that holds it
// (in this repository: bls12_381_b200/, built by `python -m bls12_381_b200.build`).
fn main() {
    if let Ok(dir) = std::env::var("BLS12381_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=bls12381_b200");
    println!("cargo:rerun-if-env-changed=BLS12381_B200_LIB_DIR");
}
