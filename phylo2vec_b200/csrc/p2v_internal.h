// Internal launcher interface between the C ABI (p2v_capi.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace p2v {

// MODE_ANC2: the first two columns of the ancestry only, (k,2) -- the third is always k+1+row; this is
// the compact transport of the host entry point, not a public output format
enum DecodeMode { MODE_PAIRS = 0, MODE_ANCESTRY = 1, MODE_EDGES = 2, MODE_ANC2 = 3 };
enum EncodeMode { FROM_PAIRS = 0, FROM_ANCESTRY = 1, FROM_EDGES = 2 };

constexpr int32_t TREE_MALFORMED = -2;   // P2V_TREE_MALFORMED of the public header
constexpr int SMALL_K_MAX = 2016;      // warp-per-tree decode in shared memory; the half2 rank scan needs K + 30 <= 2048
constexpr int ENC_SMALL_K_MAX = 2047;  // warp-per-tree encode (rows exact in fp16)
constexpr int LARGE_K_MAX = 1 << 19;   // CTA-per-tree kernels (bit mask + Fenwick in shared memory)

int sm_count();
// profiling aid: phase stamps (ns) of the last whole-grid decode / tables kernel
cudaError_t debug_phase_ns_decode(unsigned long long* out);
cudaError_t debug_phase_ns_coph(unsigned long long* out);
// One to four huge trees: every team kernel takes the whole cooperatively launched grid per tree
// (the trees go one after the other); more trees than that are faster side by side on clusters.
inline bool grid_team_trees(int64_t B, int64_t k) { return B >= 1 && B <= 4 && k >= 8192; }
// CTAs of such a grid: enough threads for one element each, at least 16, at most one per SM
inline int grid_team_ctas(int64_t elements, int threads) {
    int64_t c = (elements + threads - 1) / threads;
    if (c < 16) c = 16;
    const int sms = sm_count();
    return (int)(c > sms ? sms : c);
}
void count_launch(int n = 1);

// decode: v -> pairs / ancestry / edges
size_t decode_workspace_bytes(int64_t B, int64_t k);
// out_stride: elements between consecutive trees' outputs (0 = dense)
cudaError_t launch_decode(int mode, const void* v, int idx_bytes, int64_t B, int64_t k, void* out,
                          int32_t* status, void* ws, size_t ws_bytes, cudaStream_t stream,
                          int64_t out_stride = 0);
cudaError_t launch_check_v(const void* v, int idx_bytes, int64_t B, int64_t k, int32_t* status,
                           cudaStream_t stream);

// encode: pairs / ancestry / edges -> v
size_t encode_workspace_bytes(int64_t B, int64_t k);
cudaError_t launch_encode(int mode, const void* in, int idx_bytes, int64_t B, int64_t k, void* v,
                          int32_t* status, void* ws, size_t ws_bytes, cudaStream_t stream);

// cophenetic distances
size_t cophenetic_workspace_bytes(int64_t B, int64_t k);
// v_or_matrix: index vector (idx_bytes 4/8) with optional bls, or (k,3) f64 matrix when idx_bytes == 0
cudaError_t launch_cophenetic(const void* v_or_matrix, int idx_bytes, const double* bls, int64_t B,
                              int64_t k, int unrooted, int64_t row_begin, int64_t row_end,
                              double* out, int32_t* status, void* ws, size_t ws_bytes,
                              cudaStream_t stream, int what = 0 /* 0 distances, 1 variance-covariance */);
cudaError_t launch_cophenetic_prepare(const void* v_or_matrix, int idx_bytes, const double* bls, int64_t B,
                                      int64_t k, int unrooted, int32_t* status, void* ws, size_t ws_bytes,
                                      cudaStream_t stream);
cudaError_t launch_cophenetic_rows(void* ws, size_t ws_bytes, int64_t B, int64_t k, int weighted,
                                   int64_t row_begin, int64_t row_end, double* out, cudaStream_t stream,
                                   int what = 0);
// depth of every node, (B, 2k+1) f64 (vector/ops.rs:201-233)
cudaError_t launch_node_depths(const void* v_or_matrix, int idx_bytes, const double* bls, int64_t B, int64_t k,
                               double* out, int32_t* status, void* ws, size_t ws_bytes, cudaStream_t stream);
// check_m (matrix/base.rs:76-95) on (B,k,3) matrices: status 0, i+1 (v[i] out of bounds) or -3 (negative length)
cudaError_t launch_check_m(const double* m, const double* bl, int bl_stride, int64_t bl_tree_stride, int64_t B,
                           int64_t k, int32_t* status, int keep_status, cudaStream_t stream);
cudaError_t launch_fill_f64(double* out, int64_t count, double value, cudaStream_t stream);

// to_newick (vectors): strings at out + b * out_stride, NUL-terminated; lengths without the NUL
size_t newick_max_bytes(int64_t k);
size_t newick_workspace_bytes(int64_t B, int64_t k);
cudaError_t launch_to_newick(const void* v, int idx_bytes, int64_t B, int64_t k, char* out, int64_t out_stride,
                             int64_t* lengths, int32_t* status, void* ws, size_t ws_bytes, cudaStream_t stream);
// from_newick (vectors, Newick with parent labels): strings at in + b * in_stride, lengths[b] bytes each
// (lengths == nullptr: NUL-terminated inside the row)
size_t from_newick_workspace_bytes(int64_t B, int64_t k, int idx_bytes);
cudaError_t launch_from_newick(const char* in, int64_t in_stride, const int64_t* lengths, int64_t B, int64_t k, void* v,
                               int idx_bytes, int32_t* status, void* ws, size_t ws_bytes, cudaStream_t stream);
// Robinson-Foulds distance of tree pairs (vector/distance.rs:41-162): B1 == B2 pairs, or one tree against many
size_t robinson_foulds_workspace_bytes(int64_t B1, int64_t B2, int64_t k);
cudaError_t launch_robinson_foulds(const void* v1, const void* v2, int idx_bytes, int64_t B1, int64_t B2, int64_t k,
                                   int normalize, double* out, int32_t* status, void* ws, size_t ws_bytes,
                                   cudaStream_t stream);
// matrices <-> Newick strings with branch lengths (matrix/convert.rs:26-122)
size_t newick_matrix_max_bytes(int64_t k);
size_t newick_matrix_workspace_bytes(int64_t B, int64_t k);
cudaError_t launch_to_newick_matrix(const double* m, int64_t B, int64_t k, char* out, int64_t out_stride, int64_t* lengths,
                                    int32_t* status, void* ws, size_t ws_bytes, cudaStream_t stream);
size_t from_newick_matrix_workspace_bytes(int64_t B, int64_t k);
cudaError_t launch_from_newick_matrix(const char* in, int64_t in_stride, const int64_t* lengths, int64_t B, int64_t k,
                                      double* m, int32_t* status, void* ws, size_t ws_bytes, cudaStream_t stream);
// parse_matrix (matrix/base.rs:108-121): column 0 of (count,3) float64 rows as int32, saturating like `as usize`
cudaError_t launch_matrix_v(const double* m, int64_t count, int* out, cudaStream_t stream);
// CSV text -> numbers (io/reader.py:16-57): kind 0 integers (idx_bytes 4/8), 1 float64; counts_out (device,
// 3 int64) = fields found, rows found, error bits (1 malformed field, 2 more than `capacity` fields)
size_t csv_workspace_bytes(int64_t nbytes);
cudaError_t launch_csv_parse(const char* text, int64_t nbytes, char delim, int kind, void* out, int idx_bytes,
                             int64_t capacity, long long* counts_out, void* ws, size_t ws_bytes, cudaStream_t stream);
// int64 -> int32 index vectors (values that do not fit become -1 and fail check_v downstream)
cudaError_t launch_narrow_v(const long long* v, int64_t count, int* out, cudaStream_t stream);

}  // namespace p2v
