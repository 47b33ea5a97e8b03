// Links the prebuilt libligero_b200.so (python -m ligero_b200.build, or __graft_entry__.build()).
// LIGERO_B200_LIB_DIR = the directory that holds it (default: ../../ligero_b200 relative to this crate).
fn main() {
    let dir = std::env::var("LIGERO_B200_LIB_DIR").unwrap_or_else(|_| {
        let here = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{here}/../../ligero_b200")
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=ligero_b200");
    println!("cargo:rerun-if-env-changed=LIGERO_B200_LIB_DIR");
}
