// Link against constraint_solver_b200/libcs_b200.so (set CS_B200_LIB_DIR to its directory).
fn main() {
    let dir = std::env::var("CS_B200_LIB_DIR").unwrap_or_else(|_| "../../../constraint_solver_b200".into());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=cs_b200");
}
