// links libpvw_b200.so; PVW_B200_LIB_DIR = the directory that holds it (python pvw-rs_b200/build.py puts it into pvw-rs_b200/)
fn main() {
    let dir = std::env::var("PVW_B200_LIB_DIR").unwrap_or_else(|_| "../../pvw-rs_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=pvw_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=PVW_B200_LIB_DIR");
}
