// build.rs for the reference crate with the B200 engine linked in.
// Keeps the reference's protobuf codegen (build.rs:1-10 of the reference) and adds the link
// directives for libflechasdb_b200.so.  NOT COMPILED HERE: this environment has no rustc/cargo.
fn main() {
    // --- unchanged from the reference ---
    protobuf_codegen::Codegen::new()
        .protoc()
        .protoc_path(&protoc_bin_vendored::protoc_bin_path().unwrap())
        .includes(&["src/protos"])
        .input("src/protos/database.proto")
        .cargo_out_dir("protos")
        .run_from_script();

    // --- new: the CUDA engine ---
    // FLECHASDB_B200_DIR points at the directory that holds libflechasdb_b200.so
    // (built by `python -m flechasdb_b200.build`, i.e. nvcc -gencode arch=compute_100a,code=sm_100a).
    let dir = std::env::var("FLECHASDB_B200_DIR").expect("set FLECHASDB_B200_DIR");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=flechasdb_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=FLECHASDB_B200_DIR");
}
