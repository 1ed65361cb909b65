/*
 * phylo2vec_b200.h -- C ABI of the B200 (sm_100a) implementation of phylo2vec's
 * batched decode / encode / cophenetic-distance path.
 *
 * This is the drop-in boundary: a Rust FFI crate next to the phylo2vec core
 * (see INTEGRATION.md), a ctypes loader (phylo2vec_b200/_capi.py) or any other
 * host binds exactly these symbols.  Plain pointers and sizes only.
 *
 * Conventions
 *  - k = number of vector entries = n_leaves - 1.  Nodes: leaves 0..k,
 *    internal k+1..2k, root 2k  (reference: phylo2vec/src/vector/convert.rs:138-154).
 *  - idx_bytes selects the integer width of every index buffer of the call:
 *    8 = int64 (the reference's usize / numpy int64 layout, py-phylo2vec/src/lib.rs:108-136)
 *    4 = int32 (what the R binding uses, r-phylo2vec/src/rust/src/lib.rs:36-44).
 *  - *_batch functions take DEVICE pointers, never allocate, are asynchronous on
 *    `stream` (a cudaStream_t) and re-entrant.  All arrays are dense, row-major,
 *    batch-major:  v (B,k); pairs (B,k,2); ancestry (B,k,3); edges (B,2k,2);
 *    bls (B,k,2) f64; matrix (B,k,3) f64; distances (B,rows,n) f64.
 *  - status (B int32, may be NULL) receives per tree: 0 = ok; i+1 = v[i] is out of
 *    bounds (the reference panics with "Validation failed: v[i] = .. is out of
 *    bounds", vector/base.rs:60-67); P2V_TREE_MALFORMED for encode input that is
 *    not a tree in the reference's format; P2V_TREE_NEGATIVE_BRANCH for a tree with a negative
 *    or NaN branch length (check_m and every function that takes branch lengths: the reference's
 *    distance functions do not validate them, here such a tree is flagged instead of producing
 *    a meaningless matrix).  Output of a failed tree is unspecified.
 *  - Alignment: every index buffer is aligned to idx_bytes; pairs and edges OUTPUT buffers to
 *    2 * idx_bytes (they are written as two-element vector stores: 16 bytes for int64);
 *    float64 buffers to 8 bytes (16 gets the 128-bit store path of the distance kernels).
 *    A misaligned buffer is rejected with P2V_ERR_INVALID_ARG.
 *  - *_host functions take HOST pointers, copy in, run and copy out in chunks on
 *    internal streams and return after completion (what a single-tree drop-in of
 *    the reference's functions calls).
 *  - Return value: P2V_OK or a negative error code; p2v_last_error() gives text.
 *    Nothing throws, nothing aborts.
 */
#ifndef PHYLO2VEC_B200_H
#define PHYLO2VEC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P2V_OK 0
#define P2V_ERR_INVALID_ARG (-1)
#define P2V_ERR_CUDA (-2)
#define P2V_ERR_UNSUPPORTED (-3)
#define P2V_ERR_WORKSPACE (-4)

#define P2V_TREE_MALFORMED (-2)
#define P2V_TREE_NEGATIVE_BRANCH (-3) /* a branch length below zero or NaN (check_m, matrix/base.rs:88-94) */

typedef void *p2v_stream_t; /* cudaStream_t */

int p2v_version(void);
const char *p2v_last_error(void);
/* number of kernels this library has launched in the calling process */
int64_t p2v_launch_count(void);

/* ---- decode: v -> pairs / ancestry / edges -------------------------------
 * replaces vector::convert::to_pairs (convert.rs:103-109, avl.rs:35-211),
 * to_ancestry (convert.rs:167-188), to_edges (convert.rs:227-248), i.e. the
 * PyO3 entry points get_pairs / get_ancestry / get_edges
 * (py-phylo2vec/src/lib.rs:108-131), batched.
 * Trees with k <= 12287 (batches of trees: k <= 16384) keep all state in shared memory and need
 * no workspace; larger trees need p2v_decode_workspace_bytes(B,k) bytes of device scratch
 * (0 is returned when none is needed). */
size_t p2v_decode_workspace_bytes(int64_t B, int64_t k);
int p2v_to_pairs_batch(const void *v, int idx_bytes, int64_t B, int64_t k, void *pairs,
                       int32_t *status, void *workspace, size_t workspace_bytes,
                       p2v_stream_t stream);
int p2v_to_ancestry_batch(const void *v, int idx_bytes, int64_t B, int64_t k, void *ancestry,
                          int32_t *status, void *workspace, size_t workspace_bytes,
                          p2v_stream_t stream);
int p2v_to_edges_batch(const void *v, int idx_bytes, int64_t B, int64_t k, void *edges,
                       int32_t *status, void *workspace, size_t workspace_bytes,
                       p2v_stream_t stream);

/* ---- encode: pairs / ancestry / edges -> v -------------------------------
 * replaces vector::convert::from_pairs (convert.rs:23-30), from_ancestry
 * (:129-136, build_cherries :399-440), from_edges (:201-213, order_cherries
 * :449-477) and build_vector (:320-337, Fenwick :286-316); PyO3 entry points
 * from_pairs / from_ancestry / from_edges (py-phylo2vec/src/lib.rs:113-136).
 * k < 800 needs no workspace; larger trees need p2v_encode_workspace_bytes(B,k) bytes of
 * device scratch.  A tree flagged P2V_TREE_MALFORMED leaves unspecified values in its row of v. */
size_t p2v_encode_workspace_bytes(int64_t B, int64_t k);
int p2v_from_pairs_batch(const void *pairs, int idx_bytes, int64_t B, int64_t k, void *v,
                         int32_t *status, void *workspace, size_t workspace_bytes,
                         p2v_stream_t stream);
int p2v_from_ancestry_batch(const void *ancestry, int idx_bytes, int64_t B, int64_t k, void *v,
                            int32_t *status, void *workspace, size_t workspace_bytes,
                            p2v_stream_t stream);
int p2v_from_edges_batch(const void *edges, int idx_bytes, int64_t B, int64_t k, void *v,
                         int32_t *status, void *workspace, size_t workspace_bytes,
                         p2v_stream_t stream);

/* ---- validation: vector::base::check_v (base.rs:50-67) ------------------- */
int p2v_check_v_batch(const void *v, int idx_bytes, int64_t B, int64_t k, int32_t *status,
                      p2v_stream_t stream);

/* ---- validation of matrices: matrix::base::check_m (matrix/base.rs:76-95), PyO3 check_m
 * (py-phylo2vec/src/lib.rs:68-88), Python phylo2vec.utils.matrix.check_matrix.
 * matrix (B,k,3) float64 [v, bl1, bl2]; status[b] = 0, i + 1 (v[i] = column 0 cast like Rust's
 * `as usize` is above 2i) or P2V_TREE_NEGATIVE_BRANCH.  The reference's other two assertions
 * (empty matrix, column count) are shape checks the caller makes. */
int p2v_check_m_batch(const double *matrix, int64_t B, int64_t k, int32_t *status, p2v_stream_t stream);
int p2v_check_m_host(const double *matrix, int64_t B, int64_t k, int32_t *status);

/* ---- cophenetic distances ------------------------------------------------
 * replaces vector::graph::_cophenetic_distances (vector/graph.rs:20-90) and
 * matrix::graph::cophenetic_distances (matrix/graph.rs:23-26, parse_matrix
 * matrix/base.rs:108-121); PyO3 entry point cophenetic_distances
 * (py-phylo2vec/src/lib.rs:144-156).
 * Writes rows [row_begin,row_end) of every tree's n x n float64 matrix
 * (n = k+1) to out (B, row_end-row_begin, n).  bls == NULL: topological
 * (unit lengths).  Row blocks are how one very large tree is sharded over GPUs. */
size_t p2v_cophenetic_workspace_bytes(int64_t B, int64_t k);
int p2v_cophenetic_batch(const void *v, int idx_bytes, const double *bls, int64_t B, int64_t k,
                         int unrooted, int64_t row_begin, int64_t row_end, double *out,
                         int32_t *status, void *workspace, size_t workspace_bytes,
                         p2v_stream_t stream);
/* same with the reference's (k,3) float64 matrix input [v, bl1, bl2] */
int p2v_cophenetic_matrix_batch(const double *matrix, int64_t B, int64_t k, int unrooted,
                                int64_t row_begin, int64_t row_end, double *out, int32_t *status,
                                void *workspace, size_t workspace_bytes, p2v_stream_t stream);

/* Two-phase form of the same computation: prepare once (decode + per-tree tables kept
 * in `workspace`), then emit any number of row blocks from the prepared workspace
 * (e.g. to stream a 34 GB matrix in pieces).  weighted = 0 for topological input
 * (v without bls), 1 for branch lengths (bls or matrix input).  The rows call keeps its
 * per-row-block tables in the same workspace: one call at a time per workspace. */
int p2v_cophenetic_prepare_batch(const void *v_or_matrix, int idx_bytes /* 0 = (k,3) f64 matrix */,
                                 const double *bls, int64_t B, int64_t k, int unrooted, int32_t *status,
                                 void *workspace, size_t workspace_bytes, p2v_stream_t stream);
int p2v_cophenetic_rows_batch(void *workspace, size_t workspace_bytes, int64_t B, int64_t k,
                              int weighted, int64_t row_begin, int64_t row_end, double *out,
                              p2v_stream_t stream);

/* ---- variance-covariance matrix and node depths (the path's first "next" row) ----
 * vcv replaces vector::graph::_vcv / vcv (vector/graph.rs:211-282) and matrix::graph::vcv
 * (matrix/graph.rs:61-64); PyO3 entry point vcv (py-phylo2vec/src/lib.rs:166-172), Python
 * phylo2vec.stats.cov (stats/nodewise.py:55-71).  out[l][r] = depth of the node where leaves l
 * and r meet, diagonal = depth of the leaf; rows [row_begin,row_end) of every tree's n x n
 * float64 matrix to out (B, row_end-row_begin, n).  k = 0 is the reference's assert:
 * P2V_ERR_INVALID_ARG.  Same workspace as the distances (p2v_cophenetic_workspace_bytes);
 * p2v_vcv_rows_batch emits row blocks from a workspace prepared by
 * p2v_cophenetic_prepare_batch(unrooted = 0). */
int p2v_vcv_batch(const void *v, int idx_bytes, const double *bls, int64_t B, int64_t k,
                  int64_t row_begin, int64_t row_end, double *out, int32_t *status, void *workspace,
                  size_t workspace_bytes, p2v_stream_t stream);
int p2v_vcv_matrix_batch(const double *matrix, int64_t B, int64_t k, int64_t row_begin, int64_t row_end,
                         double *out, int32_t *status, void *workspace, size_t workspace_bytes,
                         p2v_stream_t stream);
int p2v_vcv_rows_batch(void *workspace, size_t workspace_bytes, int64_t B, int64_t k, int weighted,
                       int64_t row_begin, int64_t row_end, double *out, p2v_stream_t stream);
/* node depths replace vector::ops::_get_node_depths (vector/ops.rs:201-233) and matrix::ops::
 * get_node_depths (matrix/ops.rs:42-45); PyO3 get_node_depths (py-phylo2vec/src/lib.rs:281-287).
 * out (B, 2k+1) float64: distance from the root to node i, root (2k) = 0.
 * idx_bytes 0 = (k,3) float64 matrix input, 4/8 = index vector with optional bls. */
int p2v_node_depths_batch(const void *v_or_matrix, int idx_bytes, const double *bls, int64_t B, int64_t k,
                          double *out, int32_t *status, void *workspace, size_t workspace_bytes,
                          p2v_stream_t stream);

/* ---- balance indices (third "next" row, first half) ------------------------------------------
 * replaces vector::balance::sackin / leaf_depth_variance / b2 (phylo2vec/src/vector/balance.rs:14-58);
 * PyO3 sackin, leaf_depth_variance, b2 (py-phylo2vec/src/lib.rs:311-325); Python
 * phylo2vec.stats.balance.  out (B,3) float64: [Sackin index, variance of the leaf depths, B2] of
 * the topology of every vector. */
size_t p2v_balance_workspace_bytes(int64_t B, int64_t k);
int p2v_balance_batch(const void *v, int idx_bytes, int64_t B, int64_t k, double *out, int32_t *status,
                      void *workspace, size_t workspace_bytes, p2v_stream_t stream);

/* ---- Robinson-Foulds distance (third "next" row, second half) ---------------------------------
 * replaces vector::distance::robinson_foulds (phylo2vec/src/vector/distance.rs:41-162); PyO3
 * robinson_foulds (py-phylo2vec/src/lib.rs:295-309); Python phylo2vec.stats.robinson_foulds
 * (stats/treewise.py:10-58).  v1 (B1,k) and v2 (B2,k) with B1 == B2 (pair b = tree b of each batch)
 * or one of them 1 (that tree against every tree of the other batch); out (max(B1,B2)) float64:
 * the number of non-trivial bipartitions present in exactly one of the two (unrooted) trees,
 * divided by the total number of bipartitions when normalize != 0.  k <= 65535. */
size_t p2v_robinson_foulds_workspace_bytes(int64_t B1, int64_t B2, int64_t k);
int p2v_robinson_foulds_batch(const void *v1, const void *v2, int idx_bytes, int64_t B1, int64_t B2, int64_t k,
                              int normalize, double *out, int32_t *status, void *workspace,
                              size_t workspace_bytes, p2v_stream_t stream);
int p2v_robinson_foulds_host(const void *v1, const void *v2, int idx_bytes, int64_t B1, int64_t B2, int64_t k,
                             int normalize, double *out, int32_t *status);

/* ---- to_newick (second "next" row; vectors only) ------------------------------------------
 * replaces vector::convert::to_newick / build_newick (vector/convert.rs:341-397); PyO3 entry
 * point to_newick, vector arm (py-phylo2vec/src/lib.rs:90-96); Python phylo2vec.base.newick.
 * to_newick (base/newick.py:29-43).  Tree b's string is written NUL-terminated at
 * out + b * out_stride (out_stride >= p2v_newick_max_bytes(k)); lengths[b] (may be NULL) gets
 * its length without the NUL, 0 for an invalid vector.  k <= 2^19 (matrices with branch lengths:
 * p2v_to_newick_matrix_batch below).  The workspace holds the int32 ancestry of the batch, the
 * decoder's scratch and, for trees too large for shared memory, per-CTA tables. */
size_t p2v_newick_max_bytes(int64_t k);
size_t p2v_newick_workspace_bytes(int64_t B, int64_t k);     /* the int32 ancestry of the batch */
int p2v_to_newick_batch(const void *v, int idx_bytes, int64_t B, int64_t k, char *out, int64_t out_stride,
                        int64_t *lengths, int32_t *status, void *workspace, size_t workspace_bytes,
                        p2v_stream_t stream);

/* ---- from_newick (second "next" row; vectors) ------------------------------------------------
 * replaces newick::parse (phylo2vec/src/newick/mod.rs:73-127) and vector::convert::from_newick
 * (vector/convert.rs:268-278: parse, then order_cherries for strings with parent labels or
 * order_cherries_no_parents :534-573 for strings without, then build_vector); PyO3 entry point
 * to_vector (py-phylo2vec/src/lib.rs:98-101); Python phylo2vec.base.newick.from_newick.
 * Every tree of the batch has k + 1 leaves labelled 0..k (and, if labelled, internal nodes
 * k+1..2k).  Tree b's string starts at newick + b * in_stride and is lengths[b] bytes long
 * (lengths == NULL: it is NUL-terminated inside its row).  status[b]: 0 or P2V_TREE_MALFORMED
 * (not a binary Newick of k + 1 leaves made of "(),;" and digits and ending in ";", parent
 * labels on some ")" only, or labels that do not form a tree). */
size_t p2v_from_newick_workspace_bytes(int64_t B, int64_t k, int idx_bytes);
int p2v_from_newick_batch(const char *newick, int64_t in_stride, const int64_t *lengths, int64_t B, int64_t k,
                          void *v, int idx_bytes, int32_t *status, void *workspace, size_t workspace_bytes,
                          p2v_stream_t stream);

/* ---- matrices <-> Newick strings with branch lengths (second "next" row, matrix forms) ----------
 * to_newick replaces matrix::convert::to_newick / build_newick_with_bls (phylo2vec/src/matrix/
 * convert.rs:26-57; PyO3 to_newick, matrix arm, py-phylo2vec/src/lib.rs:90-96): matrix (B,k,3)
 * float64 rows [v, bl1, bl2]; lengths are written as the shortest decimal that reads back to the
 * same double, in the layout of the Rust ryu crate ("0.5", "3.0", "1e-7").  status: as check_m
 * (the reference validates the matrix first).  out_stride >= p2v_newick_matrix_max_bytes(k).
 * from_newick replaces matrix::convert::from_newick (:84-122) with newick::parse_with_bls
 * (newick/mod.rs:161-255) and the branch-length bookkeeping of order_cherries /
 * order_cherries_no_parents (vector/convert.rs:398-477, 534-573); PyO3 to_matrix (py-phylo2vec/src/
 * lib.rs:103-106).  Strings with or without parent labels, every node but the root followed by
 * ":length" (any literal Rust's str::parse::<f64> accepts; correctly rounded); every tree of the
 * batch has k + 1 leaves labelled 0..k.  status[b]: 0 or P2V_TREE_MALFORMED. */
size_t p2v_newick_matrix_max_bytes(int64_t k);
size_t p2v_newick_matrix_workspace_bytes(int64_t B, int64_t k);
int p2v_to_newick_matrix_batch(const double *matrix, int64_t B, int64_t k, char *out, int64_t out_stride,
                               int64_t *lengths, int32_t *status, void *workspace, size_t workspace_bytes,
                               p2v_stream_t stream);
int p2v_to_newick_matrix_host(const double *matrix, int64_t B, int64_t k, char *out, int64_t out_stride,
                              int64_t *lengths, int32_t *status);
size_t p2v_from_newick_matrix_workspace_bytes(int64_t B, int64_t k);
int p2v_from_newick_matrix_batch(const char *newick, int64_t in_stride, const int64_t *lengths, int64_t B, int64_t k,
                                 double *matrix, int32_t *status, void *workspace, size_t workspace_bytes,
                                 p2v_stream_t stream);
int p2v_from_newick_matrix_host(const char *newick, int64_t in_stride, int64_t B, int64_t k, double *matrix,
                                int32_t *status);

/* ---- CSV text -> numbers (fourth "next" row: the file format either side of the batch API) -------
 * the parsing inside phylo2vec.io.reader.load / read (py-phylo2vec/phylo2vec/io/reader.py:16-57,
 * numpy.genfromtxt) for the files io/writer.py:13-50 writes ("%d" vectors, "%.18e" matrices) and
 * for batches (one tree per line).  text: nbytes characters in DEVICE memory; fields are
 * separated by `delimiter`, newlines or blanks.  kind 0: integers into `out` (idx_bytes 4/8;
 * a malformed field becomes -1); kind 1: float64, correctly rounded.  At most `capacity` fields
 * are written, in text order.  counts (device, 3 int64): fields found, rows (non-empty lines)
 * found, error bits (1: a malformed field, 2: more fields than capacity). */
size_t p2v_csv_workspace_bytes(int64_t nbytes);
int p2v_csv_parse_batch(const char *text, int64_t nbytes, int delimiter, int kind, void *out, int idx_bytes,
                        int64_t capacity, int64_t *counts, void *workspace, size_t workspace_bytes,
                        p2v_stream_t stream);

/* ---- host-buffer entry points (H2D + kernels + D2H inside the call) ------ */
int p2v_to_pairs_host(const void *v, int idx_bytes, int64_t B, int64_t k, void *pairs, int32_t *status);
int p2v_to_ancestry_host(const void *v, int idx_bytes, int64_t B, int64_t k, void *ancestry, int32_t *status);
int p2v_to_edges_host(const void *v, int idx_bytes, int64_t B, int64_t k, void *edges, int32_t *status);
int p2v_from_pairs_host(const void *pairs, int idx_bytes, int64_t B, int64_t k, void *v, int32_t *status);
int p2v_from_ancestry_host(const void *ancestry, int idx_bytes, int64_t B, int64_t k, void *v, int32_t *status);
int p2v_from_edges_host(const void *edges, int idx_bytes, int64_t B, int64_t k, void *v, int32_t *status);
int p2v_check_v_host(const void *v, int idx_bytes, int64_t B, int64_t k, int32_t *status);
int p2v_cophenetic_host(const void *v, int idx_bytes, const double *bls, int64_t B, int64_t k,
                        int unrooted, int64_t row_begin, int64_t row_end, double *out, int32_t *status);
int p2v_cophenetic_matrix_host(const double *matrix, int64_t B, int64_t k, int unrooted,
                               int64_t row_begin, int64_t row_end, double *out, int32_t *status);

int p2v_vcv_host(const void *v, int idx_bytes, const double *bls, int64_t B, int64_t k,
                 int64_t row_begin, int64_t row_end, double *out, int32_t *status);
int p2v_vcv_matrix_host(const double *matrix, int64_t B, int64_t k, int64_t row_begin, int64_t row_end,
                        double *out, int32_t *status);
int p2v_node_depths_host(const void *v_or_matrix, int idx_bytes, const double *bls, int64_t B, int64_t k,
                         double *out, int32_t *status);

int p2v_to_newick_host(const void *v, int idx_bytes, int64_t B, int64_t k, char *out, int64_t out_stride,
                       int64_t *lengths, int32_t *status);

int p2v_balance_host(const void *v, int idx_bytes, int64_t B, int64_t k, double *out, int32_t *status);

/* NUL-terminated strings, one per in_stride bytes */
int p2v_from_newick_host(const char *newick, int64_t in_stride, int64_t B, int64_t k, void *v, int idx_bytes,
                         int32_t *status);

/* The *_host entry points keep three streams, their device staging buffers and (for pageable
 * or compact transport) pinned host staging per host thread; they grow on demand, up to about
 * 3 GB of device memory and 1.2 GB of pinned memory for a 1 M-tree to_ancestry call.
 * A call that would leave more than the cache limit behind (default 512 MiB; device + pinned
 * bytes per thread) frees its staging before it returns; a caller that streams many large
 * batches raises the limit (-1 = keep everything) to keep the allocations between calls.
 * A thread that exits frees its cache.  p2v_host_release() frees the calling thread's cache now. */
void p2v_host_release(void);
int64_t p2v_host_cache_bytes(void);          /* bytes the calling thread currently keeps */
/* The worker threads of the *_host entry points run on the CPUs next to the current device (its
 * NUMA node, sysfs local_cpulist) when the machine has more than one node; P2V_HOST_NUMA=0
 * disables that.  p2v_host_bind_thread() does the same for the CALLING thread (call it once per
 * process before allocating pinned buffers, so that they land on that node); returns the number
 * of CPUs it was bound to, 0 if it was left alone. */
int p2v_host_bind_thread(void);
void p2v_host_cache_limit(int64_t bytes);    /* process-wide limit per thread */

/* profiling aid: nanosecond stamps at the phase boundaries of the last whole-grid
 * decode (which = 0) / tables (which = 1) kernel; out[64] */
int p2v_debug_phase_ns(int which, uint64_t *out);

/* ---- measurement helper: a plain 128-bit streaming fill of `bytes` bytes, the
 * write-only ceiling the distance kernel is compared with in bench.py -------- */
int p2v_fill_f64(double *out, int64_t count, double value, p2v_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PHYLO2VEC_B200_H */
