// Link against libphylo2vec_b200.so built by `python -m phylo2vec_b200.build`.
fn main() {
    let dir = std::env::var("PHYLO2VEC_B200_LIB_DIR").unwrap_or_else(|_| "../../../phylo2vec_b200/lib".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=phylo2vec_b200");
}
