fn main() {
    // RIPP_B200_LIB_DIR = directory holding libripp_b200.so (ripp_b200/ of the ripp-b200 checkout)
    if let Ok(dir) = std::env::var("RIPP_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=ripp_b200");
}
