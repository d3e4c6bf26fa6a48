// Links libvibrato_b200.so from <repo>/vibrato_b200 (override with VIBRATO_B200_LIB_DIR).
fn main() {
    let dir = std::env::var("VIBRATO_B200_LIB_DIR").unwrap_or_else(|_| {
        let here = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{here}/../vibrato_b200")
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=vibrato_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=VIBRATO_B200_LIB_DIR");
}
