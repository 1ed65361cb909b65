//! C-ABI bindings of libphylo2vec_b200 (include/phylo2vec_b200.h) and safe wrappers
//! with the signatures of the functions they replace in the `phylo2vec` core crate:
//! `vector::convert::{to_pairs,to_ancestry,to_edges,from_pairs,from_ancestry,from_edges}`
//! (phylo2vec/src/vector/convert.rs) and `vector::graph::cophenetic_distances` /
//! `matrix::graph::cophenetic_distances` (vector/graph.rs:107-109, matrix/graph.rs:23-26).
//! SOURCE ONLY: not compiled in this repository's image (no cargo/rustc).
use std::os::raw::{c_char, c_int, c_void};

pub type Stream = *mut c_void; // cudaStream_t

extern "C" {
    pub fn p2v_version() -> c_int;
    pub fn p2v_last_error() -> *const c_char;
    pub fn p2v_to_pairs_host(v: *const c_void, idx_bytes: c_int, b: i64, k: i64, out: *mut c_void, status: *mut i32) -> c_int;
    pub fn p2v_to_ancestry_host(v: *const c_void, idx_bytes: c_int, b: i64, k: i64, out: *mut c_void, status: *mut i32) -> c_int;
    pub fn p2v_to_edges_host(v: *const c_void, idx_bytes: c_int, b: i64, k: i64, out: *mut c_void, status: *mut i32) -> c_int;
    pub fn p2v_from_pairs_host(x: *const c_void, idx_bytes: c_int, b: i64, k: i64, v: *mut c_void, status: *mut i32) -> c_int;
    pub fn p2v_from_ancestry_host(x: *const c_void, idx_bytes: c_int, b: i64, k: i64, v: *mut c_void, status: *mut i32) -> c_int;
    pub fn p2v_from_edges_host(x: *const c_void, idx_bytes: c_int, b: i64, k: i64, v: *mut c_void, status: *mut i32) -> c_int;
    pub fn p2v_cophenetic_host(v: *const c_void, idx_bytes: c_int, bls: *const f64, b: i64, k: i64, unrooted: c_int,
                               row_begin: i64, row_end: i64, out: *mut f64, status: *mut i32) -> c_int;
    pub fn p2v_cophenetic_matrix_host(m: *const f64, b: i64, k: i64, unrooted: c_int, row_begin: i64, row_end: i64,
                                      out: *mut f64, status: *mut i32) -> c_int;
    pub fn p2v_vcv_host(v: *const c_void, idx_bytes: c_int, bls: *const f64, b: i64, k: i64, row_begin: i64, row_end: i64,
                        out: *mut f64, status: *mut i32) -> c_int;
    pub fn p2v_vcv_matrix_host(m: *const f64, b: i64, k: i64, row_begin: i64, row_end: i64, out: *mut f64,
                               status: *mut i32) -> c_int;
    pub fn p2v_node_depths_host(v_or_m: *const c_void, idx_bytes: c_int, bls: *const f64, b: i64, k: i64,
                                out: *mut f64, status: *mut i32) -> c_int;
    pub fn p2v_balance_host(v: *const c_void, idx_bytes: c_int, b: i64, k: i64, out: *mut f64, status: *mut i32) -> c_int;
    pub fn p2v_newick_max_bytes(k: i64) -> usize;
    pub fn p2v_to_newick_host(v: *const c_void, idx_bytes: c_int, b: i64, k: i64, out: *mut c_char, out_stride: i64,
                              lengths: *mut i64, status: *mut i32) -> c_int;
    pub fn p2v_from_newick_host(newick: *const c_char, in_stride: i64, b: i64, k: i64, v: *mut c_void, idx_bytes: c_int,
                                status: *mut i32) -> c_int;
    // device-pointer entry points (asynchronous on `stream`, caller-owned buffers)
    pub fn p2v_decode_workspace_bytes(b: i64, k: i64) -> usize;
    pub fn p2v_to_ancestry_batch(v: *const c_void, idx_bytes: c_int, b: i64, k: i64, out: *mut c_void, status: *mut i32,
                                 ws: *mut c_void, ws_bytes: usize, stream: Stream) -> c_int;
    pub fn p2v_cophenetic_workspace_bytes(b: i64, k: i64) -> usize;
    pub fn p2v_cophenetic_batch(v: *const c_void, idx_bytes: c_int, bls: *const f64, b: i64, k: i64, unrooted: c_int,
                                row_begin: i64, row_end: i64, out: *mut f64, status: *mut i32, ws: *mut c_void,
                                ws_bytes: usize, stream: Stream) -> c_int;
}

fn check(rc: c_int, status: i32, v: &[usize]) {
    if rc != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(p2v_last_error()) }.to_string_lossy().into_owned();
        panic!("phylo2vec_b200: {msg} (code {rc})");
    }
    if status > 0 {
        // same message as vector::base::check_max (phylo2vec/src/vector/base.rs:60-67)
        let idx = (status - 1) as usize;
        panic!("Validation failed: v[{idx}] = {} is out of bounds (max = {})", v[idx], 2 * idx);
    }
    assert!(status == 0, "malformed tree input");
}

/// Drop-in for `phylo2vec::vector::convert::to_ancestry` (convert.rs:167-188).
pub fn to_ancestry(v: &[usize]) -> Vec<[usize; 3]> {
    let k = v.len();
    let mut out = vec![[0usize; 3]; k];
    let mut status = 0i32;
    let rc = unsafe { p2v_to_ancestry_host(v.as_ptr().cast(), 8, 1, k as i64, out.as_mut_ptr().cast(), &mut status) };
    check(rc, status, v);
    out
}

/// Drop-in for `phylo2vec::vector::convert::to_pairs` (convert.rs:103-109).
pub fn to_pairs(v: &[usize]) -> Vec<(usize, usize)> {
    let k = v.len();
    let mut out = vec![(0usize, 0usize); k]; // (usize, usize) is two consecutive words, like the (k,2) layout
    let mut status = 0i32;
    let rc = unsafe { p2v_to_pairs_host(v.as_ptr().cast(), 8, 1, k as i64, out.as_mut_ptr().cast(), &mut status) };
    check(rc, status, v);
    out
}

/// Drop-in for `phylo2vec::vector::convert::from_ancestry` (convert.rs:129-136).
pub fn from_ancestry(ancestry: &[[usize; 3]]) -> Vec<usize> {
    let k = ancestry.len();
    let mut v = vec![0usize; k];
    let mut status = 0i32;
    let rc = unsafe { p2v_from_ancestry_host(ancestry.as_ptr().cast(), 8, 1, k as i64, v.as_mut_ptr().cast(), &mut status) };
    check(rc, status, &v);
    v
}

/// Drop-in for `phylo2vec::vector::graph::cophenetic_distances` (graph.rs:107-109): row-major (n,n).
pub fn cophenetic_distances(v: &[usize], unrooted: bool) -> Vec<f64> {
    let k = v.len();
    let n = k + 1;
    let mut out = vec![0f64; n * n];
    let mut status = 0i32;
    let rc = unsafe {
        p2v_cophenetic_host(v.as_ptr().cast(), 8, std::ptr::null(), 1, k as i64, unrooted as c_int, 0, n as i64,
                            out.as_mut_ptr(), &mut status)
    };
    check(rc, status, v);
    out
}

/// Drop-in for `phylo2vec::vector::graph::vcv` (graph.rs:280-282); row-major (n, n).
pub fn vcv(v: &[usize]) -> Vec<f64> {
    let k = v.len();
    assert!(k > 0, "Variance-covariance matrix not supported for trees with < 2 leaves.");
    let n = k + 1;
    let mut out = vec![0f64; n * n];
    let mut status = 0i32;
    let rc = unsafe {
        p2v_vcv_host(v.as_ptr().cast(), 8, std::ptr::null(), 1, k as i64, 0, n as i64, out.as_mut_ptr(), &mut status)
    };
    check(rc, status, v);
    out
}

/// Drop-in for `phylo2vec::vector::ops::get_node_depths` (ops.rs:287-289).
pub fn get_node_depths(v: &[usize]) -> Vec<f64> {
    let k = v.len();
    let mut out = vec![0f64; 2 * k + 1];
    let mut status = 0i32;
    let rc = unsafe {
        p2v_node_depths_host(v.as_ptr().cast(), 8, std::ptr::null(), 1, k as i64, out.as_mut_ptr(), &mut status)
    };
    check(rc, status, v);
    out
}

/// Drop-in for `phylo2vec::vector::convert::to_newick` (convert.rs:394-397).
pub fn to_newick(v: &[usize]) -> String {
    let k = v.len();
    let stride = unsafe { p2v_newick_max_bytes(k as i64) };
    let mut buf = vec![0u8; stride];
    let (mut len, mut status) = (0i64, 0i32);
    let rc = unsafe {
        p2v_to_newick_host(v.as_ptr().cast(), 8, 1, k as i64, buf.as_mut_ptr().cast(), stride as i64, &mut len, &mut status)
    };
    check(rc, status, v);
    buf.truncate(len as usize);
    String::from_utf8(buf).expect("ASCII")
}

/// Drop-in for `phylo2vec::vector::convert::from_newick` (convert.rs:268-278), with or without
/// parent labels.  Panics on a malformed string like the reference's `expect("failed to get cherries")`.
pub fn from_newick(newick: &str) -> Vec<usize> {
    let k = newick.bytes().filter(|&c| c == b',').count();
    let mut v = vec![0usize; k];
    if k == 0 {
        return v;
    }
    let c = std::ffi::CString::new(newick).expect("NUL inside the Newick string");
    let mut status = 0i32;
    let rc = unsafe {
        p2v_from_newick_host(c.as_ptr(), (newick.len() + 1) as i64, 1, k as i64, v.as_mut_ptr().cast(), 8, &mut status)
    };
    check(rc, 0, &v);
    assert!(status == 0, "failed to get cherries");
    v
}

fn balance(v: &[usize]) -> [f64; 3] {
    let mut out = [0f64; 3];
    let mut status = 0i32;
    let rc = unsafe { p2v_balance_host(v.as_ptr().cast(), 8, 1, v.len() as i64, out.as_mut_ptr(), &mut status) };
    check(rc, status, v);
    out
}

/// Drop-ins for `phylo2vec::vector::balance::{sackin, leaf_depth_variance, b2}` (balance.rs:14-58).
pub fn sackin(v: &[usize]) -> usize { balance(v)[0] as usize }
pub fn leaf_depth_variance(v: &[usize]) -> f64 { balance(v)[1] }
pub fn b2(v: &[usize]) -> f64 { balance(v)[2] }
