// Links the prebuilt engine; set RVC_B200_LIB_DIR to the directory holding librvc_b200.so.
fn main() {
    if let Ok(dir) = std::env::var("RVC_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=rvc_b200");
}
