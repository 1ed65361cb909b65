// CSV text -> numeric arrays on the GPU: the load side of the reference's file format.
//
// Replaces the parsing inside (reference, paths relative to /root/reference):
//   phylo2vec.io.reader.load / read     py-phylo2vec/phylo2vec/io/reader.py:16-57  (numpy.genfromtxt)
// for single trees (a vector: one integer per line; a matrix: k lines "v,bl1,bl2", written by
// io/writer.py:13-50 with "%d" / "%.18e") and for batches (one tree per line: k integers, or 3k numbers).
//
// The text is cut into chunks of 1024 bytes.  A field starts at a character that is not a separator
// (delimiter, newline, carriage return, blank, tab) and follows one; a row starts at a field that is the first of
// its line.  Pass 1 counts the field and row starts of every chunk, a scan turns the counts into bases, pass 2
// gives every field its index and parses it where it stands: digit runs for integers, Rust / numpy float
// literals through the correctly rounded parser of p2v_float_text.cuh for float64.
#include <cstdlib>

#include "p2v_common.cuh"
#include "p2v_float_text.cuh"
#include "p2v_internal.h"

namespace p2v {

constexpr int CSV_THREADS = 256;
constexpr int CSV_CPT = 4;                         // characters per thread
constexpr int CSV_CHUNK = CSV_THREADS * CSV_CPT;   // 1024 bytes per CTA step

__device__ __forceinline__ bool csv_sep(char c, char delim) {
    return c == delim || c == '\n' || c == '\r' || c == ' ' || c == '\t';
}
// the first field of a line: only blanks between it and the previous newline (or the start of the text)
__device__ __forceinline__ bool csv_row_start(const char* s, int64_t i) {
    for (int64_t j = i - 1; j >= 0; --j) {
        const char c = s[j];
        if (c == '\n') return true;
        if (c != ' ' && c != '\t' && c != '\r') return false;
    }
    return true;
}

// counts[2 * chunk] = fields starting in the chunk, counts[2 * chunk + 1] = rows starting in it
__global__ void __launch_bounds__(CSV_THREADS)
csv_count_kernel(const char* __restrict__ s, int64_t nbytes, char delim, int64_t nchunks, unsigned* __restrict__ counts) {
    __shared__ int s_f[CSV_THREADS / 32], s_r[CSV_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        int nf = 0, nr = 0;
#pragma unroll
        for (int c = 0; c < CSV_CPT; ++c) {
            const int64_t i = chunk * CSV_CHUNK + (int64_t)c * CSV_THREADS + threadIdx.x;
            if (i < nbytes) {
                const char ch = s[i];
                const char pv = i > 0 ? s[i - 1] : '\n';
                if (!csv_sep(ch, delim) && csv_sep(pv, delim)) {
                    ++nf;
                    if (pv != delim && csv_row_start(s, i)) ++nr;
                }
            }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) { nf += __shfl_xor_sync(FULL, nf, o); nr += __shfl_xor_sync(FULL, nr, o); }
        __syncthreads();
        if (lane == 0) { s_f[warp] = nf; s_r[warp] = nr; }
        __syncthreads();
        if (threadIdx.x == 0) {
            int f = 0, r = 0;
            for (int w = 0; w < CSV_THREADS / 32; ++w) { f += s_f[w]; r += s_r[w]; }
            counts[2 * chunk] = (unsigned)f; counts[2 * chunk + 1] = (unsigned)r;
        }
    }
}

// exclusive scan of the per-chunk field counts (one CTA); totals[0] = fields, totals[1] = rows
__global__ void __launch_bounds__(1024)
csv_scan_kernel(unsigned* __restrict__ counts, int64_t nchunks, unsigned long long* __restrict__ bases,
                long long* __restrict__ totals) {
    __shared__ unsigned long long s_w[32], s_r[32];
    __shared__ unsigned long long carry_f, carry_r;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { carry_f = 0; carry_r = 0; }
    __syncthreads();
    for (int64_t base = 0; base < nchunks; base += 1024) {
        const int64_t c = base + threadIdx.x;
        const unsigned long long f = c < nchunks ? counts[2 * c] : 0ull, r = c < nchunks ? counts[2 * c + 1] : 0ull;
        unsigned long long inc = f, rs = r;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long y = __shfl_up_sync(FULL, inc, d), z = __shfl_up_sync(FULL, rs, d);
            if (lane >= d) { inc += y; rs += z; }
        }
        if (lane == 31) { s_w[warp] = inc; s_r[warp] = rs; }
        __syncthreads();
        if (warp == 0) {
            unsigned long long t = s_w[lane], u = s_r[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long y = __shfl_up_sync(FULL, t, d), z = __shfl_up_sync(FULL, u, d);
                if (lane >= d) { t += y; u += z; }
            }
            s_w[lane] = t; s_r[lane] = u;
        }
        __syncthreads();
        const unsigned long long before = carry_f + (warp ? s_w[warp - 1] : 0ull) + inc - f;
        if (c < nchunks) bases[c] = before;
        __syncthreads();
        if (threadIdx.x == 1023) { carry_f += s_w[31]; carry_r += s_r[31]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { totals[0] = (long long)carry_f; totals[1] = (long long)carry_r; }
}

// kind 0: integers into IdxT out; kind 1: float64.  err[0] |= 1 malformed field, |= 2 more fields than `capacity`
template <typename IdxT, int KIND>
__global__ void __launch_bounds__(CSV_THREADS)
csv_parse_kernel(const char* __restrict__ s, int64_t nbytes, char delim, int64_t nchunks,
                 const unsigned long long* __restrict__ bases, void* __restrict__ out_v, int64_t capacity,
                 int* __restrict__ err) {
    __shared__ int s_cnt[CSV_CPT][CSV_THREADS / 32];
    // literals of more than 19 digits that the fast parser leaves undecided (about one in a thousand of those):
    // queued, then decided exactly by one thread with the big-number scratch below (p2v_float_text.cuh)
    constexpr int UNDQ = 32;
    __shared__ int s_undn;
    __shared__ long long s_und[UNDQ][3];
    __shared__ uint32_t s_big[KIND == 1 ? 2 * p2v_ft::FT_BIG_LIMBS : 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
        bool start[CSV_CPT];
        unsigned ball[CSV_CPT];
#pragma unroll
        for (int c = 0; c < CSV_CPT; ++c) {
            const int64_t i = chunk * CSV_CHUNK + (int64_t)c * CSV_THREADS + threadIdx.x;
            start[c] = false;
            if (i < nbytes) {
                const char ch = s[i];
                const char pv = i > 0 ? s[i - 1] : '\n';
                start[c] = !csv_sep(ch, delim) && csv_sep(pv, delim);
            }
            ball[c] = __ballot_sync(FULL, start[c]);
        }
        __syncthreads();
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < CSV_CPT; ++c) s_cnt[c][warp] = __popc(ball[c]);
        }
        if (threadIdx.x == 0) s_undn = 0;
        __syncthreads();
#pragma unroll
        for (int c = 0; c < CSV_CPT; ++c) {
            if (!start[c]) continue;
            // fields of this chunk before mine: earlier character groups, earlier warps of my group, lower lanes
            int before = __popc(ball[c] & lanes_lt(lane));
            for (int cc = 0; cc < CSV_CPT; ++cc)
                for (int w = 0; w < CSV_THREADS / 32; ++w)
                    if (cc < c || (cc == c && w < warp)) before += s_cnt[cc][w];
            const int64_t idx = (int64_t)bases[chunk] + before;
            const int64_t i0 = chunk * CSV_CHUNK + (int64_t)c * CSV_THREADS + threadIdx.x;
            int64_t i1 = i0;
            while (i1 < nbytes && !csv_sep(s[i1], delim) && i1 - i0 < 512) ++i1;
            bool ok = i1 - i0 < 512;
            if (idx >= capacity) { atomicOr(err, 2); continue; }
            if (KIND == 0) {
                long long v = 0;
                int64_t i = i0;
                bool neg = false;
                if (s[i] == '-' || s[i] == '+') { neg = s[i] == '-'; ++i; }
                ok = ok && i < i1 && (i1 - i) <= 18;
                for (; i < i1; ++i) {
                    if (s[i] < '0' || s[i] > '9') { ok = false; break; }
                    v = v * 10 + (s[i] - '0');
                }
                reinterpret_cast<IdxT*>(out_v)[idx] = ok ? (IdxT)(neg ? -v : v) : (IdxT)-1;
            } else {
                double x = 0.0;
                if (ok) {
                    const int rc = p2v_ft::parse_f64(s + i0, (int)(i1 - i0), &x);
                    if (rc == p2v_ft::PARSE_F64_UNDECIDED) {
                        const int slot = atomicAdd(&s_undn, 1);
                        if (slot < UNDQ) { s_und[slot][0] = i0; s_und[slot][1] = i1; s_und[slot][2] = idx; }
                        else ok = false;
                    } else ok = rc == p2v_ft::PARSE_F64_OK;
                }
                reinterpret_cast<double*>(out_v)[idx] = x;
            }
            if (!ok) atomicOr(err, 1);
        }
        __syncthreads();
        if (KIND == 1 && threadIdx.x == 0) {
            const int nu = s_undn < UNDQ ? s_undn : UNDQ;
            for (int u = 0; u < nu; ++u) {
                double x = 0.0;
                const int rc = p2v_ft::parse_f64_exact(s + s_und[u][0], (int)(s_und[u][1] - s_und[u][0]), s_big,
                                                       s_big + p2v_ft::FT_BIG_LIMBS, &x);
                if (rc != p2v_ft::PARSE_F64_OK) atomicOr(err, 1);
                reinterpret_cast<double*>(out_v)[s_und[u][2]] = x;
            }
        }
        __syncthreads();
    }
}

// workspace: [ counts (2 u32 per chunk) | bases (u64 per chunk) | totals (2 i64) | err (int) ]
struct CsvWs { size_t counts, bases, totals, err, total; int64_t nchunks; };
static CsvWs csv_ws(int64_t nbytes) {
    CsvWs w;
    w.nchunks = (nbytes + CSV_CHUNK - 1) / CSV_CHUNK;
    size_t o = 0;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    w.counts = o; o += al(8 * (size_t)(w.nchunks > 0 ? w.nchunks : 1));
    w.bases = o; o += al(8 * (size_t)(w.nchunks > 0 ? w.nchunks : 1));
    w.totals = o; o += 256;
    w.err = o; o += 256;
    w.total = o;
    return w;
}
size_t csv_workspace_bytes(int64_t nbytes) { return nbytes <= 0 ? 512 : csv_ws(nbytes).total; }

// counts_out (device, 3 int64): fields, rows, error bits
cudaError_t launch_csv_parse(const char* text, int64_t nbytes, char delim, int kind, void* out, int idx_bytes,
                             int64_t capacity, long long* counts_out, void* ws, size_t ws_bytes, cudaStream_t stream) {
    const CsvWs W = csv_ws(nbytes > 0 ? nbytes : 1);
    if (!ws || ws_bytes < W.total) return cudaErrorInvalidValue;
    unsigned char* wsb = static_cast<unsigned char*>(ws);
    unsigned* counts = reinterpret_cast<unsigned*>(wsb + W.counts);
    unsigned long long* bases = reinterpret_cast<unsigned long long*>(wsb + W.bases);
    long long* totals = reinterpret_cast<long long*>(wsb + W.totals);
    int* err = reinterpret_cast<int*>(wsb + W.err);
    cudaError_t e = cudaMemsetAsync(err, 0, sizeof(int), stream);
    if (e != cudaSuccess) return e;
    const int64_t nchunks = nbytes > 0 ? W.nchunks : 0;
    const int64_t cap = (int64_t)sm_count() * 16;
    const int grid = (int)(nchunks < cap ? (nchunks > 0 ? nchunks : 1) : cap);
    csv_count_kernel<<<grid, CSV_THREADS, 0, stream>>>(text, nbytes, delim, nchunks, counts);
    csv_scan_kernel<<<1, 1024, 0, stream>>>(counts, nchunks, bases, totals);
    if (kind == 0 && idx_bytes == 8)
        csv_parse_kernel<long long, 0><<<grid, CSV_THREADS, 0, stream>>>(text, nbytes, delim, nchunks, bases, out, capacity, err);
    else if (kind == 0 && idx_bytes == 4)
        csv_parse_kernel<int, 0><<<grid, CSV_THREADS, 0, stream>>>(text, nbytes, delim, nchunks, bases, out, capacity, err);
    else if (kind == 1)
        csv_parse_kernel<long long, 1><<<grid, CSV_THREADS, 0, stream>>>(text, nbytes, delim, nchunks, bases, out, capacity, err);
    else
        return cudaErrorInvalidValue;
    count_launch(3);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(counts_out, totals, 2 * sizeof(long long), cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) return e;
    // error bits as the third int64 (the upper half is cleared first)
    e = cudaMemsetAsync(counts_out + 2, 0, sizeof(long long), stream);
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(counts_out + 2, err, sizeof(int), cudaMemcpyDeviceToDevice, stream);
}

}  // namespace p2v
