// Batched to_newick for Phylo2Vec vectors (topology only), sm_100a.
//
// Replaces (reference, paths relative to /root/reference):
//   vector::convert::to_newick / build_newick / prepare_cache   phylo2vec/src/vector/convert.rs:341-397
//   PyO3 to_newick (vector arm)                                    py-phylo2vec/src/lib.rs:90-96
//
// The reference concatenates per-leaf string caches pair by pair; the result is the Newick string
// of the ancestry tree with children in (c1, c2) order and every node labelled by its id:
//   S(leaf) = dec(leaf),  S(u) = "(" S(c1) "," S(c2) ")" dec(u),  out = S(root) ";".
// The batch is first decoded to an int32 ancestry by the batched decode kernel (workspace); then
// one 128-thread CTA per tree keeps the tree in shared memory and places every token directly:
//   1. the ancestry rows are copied in;
//   2. pointer jumping down the first-child and second-child pointers gives every node its
//      leftmost / rightmost leaf, the number of "(" a first-child chain opens and the bytes a
//      second-child chain closes (")" + label per node);
//   3. every internal node links the last leaf of its first child to the first leaf of its second
//      child: the leaves form a list in string order; weighted list ranking (weight = bytes
//      emitted around the leaf: ",", the "(" run, its label, the ")label" run) gives every leaf
//      its byte offset;
//   4. leaves write [","] "(((" label, internal nodes write ")" label, the string is assembled in
//      shared memory and copied out with coalesced stores.
// Sizes: k <= 1023 (one mask word per lane in the decoder, uint16 tables).  Larger trees and the
// matrix form (branch lengths need the reference's shortest-round-trip float formatting) are
// not built yet: P2V_ERR_UNSUPPORTED.
#include <cstdlib>

#include "p2v_common.cuh"
#include "p2v_internal.h"

namespace p2v {

constexpr int NW_THREADS = 128;
constexpr int NW_K_MAX = 1023;
constexpr unsigned NW_NONE = 0xFFFFu;

__host__ __device__ inline int dec_digits(unsigned x) {
    return 1 + (x >= 10u) + (x >= 100u) + (x >= 1000u) + (x >= 10000u);
}

// exact length of the string of a tree with k + 1 leaves (without the terminating NUL)
__host__ __device__ inline size_t newick_bytes(int64_t k) {
    size_t total = 3 * (size_t)k + 1;                  // "(", ",", ")" per internal node and the ";"
    for (int64_t x = 0; x <= 2 * k; ++x) total += (size_t)dec_digits((unsigned)x);
    return total;
}

struct NewickLayout {
    size_t anc, dl, dr, al, ar, nxt, wsum, wleaf, open, closel, start, str, total;
    size_t strcap;
};
__host__ __device__ inline size_t nw_al16(size_t x) { return (x + 15) & ~(size_t)15; }
__host__ __device__ inline NewickLayout newick_layout(int k) {      // all byte counts fit uint16: <= 11.3 KB
    NewickLayout L;
    const size_t n = (size_t)k + 1, N = 2 * (size_t)k + 1;
    size_t o = 0;
    L.anc = o; o = nw_al16(o + 4 * 3 * (size_t)k);
    L.dl = o; o = nw_al16(o + 2 * 2 * N);           // first-child jump pointers, two buffers
    L.dr = o; o = nw_al16(o + 2 * 2 * N);           // second-child jump pointers
    L.al = o; o = nw_al16(o + 2 * 2 * N);           // "(" opened along the first-child chain
    L.ar = o; o = nw_al16(o + 2 * 2 * N);           // bytes closed along the second-child chain
    L.nxt = o; o = nw_al16(o + 2 * 2 * n);          // leaf list, two buffers
    L.wsum = o; o = nw_al16(o + 2 * 2 * n);         // bytes from this leaf to the end of the string
    L.wleaf = o; o = nw_al16(o + 2 * n);            // bytes emitted around one leaf
    L.open = o; o = nw_al16(o + 2 * n);
    L.closel = o; o = nw_al16(o + 2 * n);
    L.start = o; o = nw_al16(o + 2 * n);
    L.strcap = nw_al16(newick_bytes(k) + 2);
    L.str = o; o = nw_al16(o + L.strcap);
    L.total = o;
    return L;
}

__device__ __forceinline__ void put_dec(char* p, unsigned x, int digits) {
    for (int i = digits - 1; i >= 0; --i) { p[i] = (char)('0' + x % 10u); x /= 10u; }
}

__global__ void __launch_bounds__(NW_THREADS)
newick_small_kernel(const int* __restrict__ anc_all, int64_t B, int k, char* __restrict__ out, int64_t out_stride,
                    int64_t* __restrict__ lengths, const int32_t* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char dyn[];
    const NewickLayout L = newick_layout(k);
    const int n = k + 1, N = 2 * k + 1, root = 2 * k;
    const int tid = threadIdx.x;
    const int NT = blockDim.x;                    // 32 / 64 / 128 threads: small trees get small CTAs
    int* anc = reinterpret_cast<int*>(dyn + L.anc);
    uint16_t* dl = reinterpret_cast<uint16_t*>(dyn + L.dl);
    uint16_t* dr = reinterpret_cast<uint16_t*>(dyn + L.dr);
    uint16_t* al = reinterpret_cast<uint16_t*>(dyn + L.al);
    uint16_t* ar = reinterpret_cast<uint16_t*>(dyn + L.ar);
    uint16_t* nxt = reinterpret_cast<uint16_t*>(dyn + L.nxt);
    uint16_t* wsum = reinterpret_cast<uint16_t*>(dyn + L.wsum);
    uint16_t* wleaf = reinterpret_cast<uint16_t*>(dyn + L.wleaf);
    uint16_t* nopen = reinterpret_cast<uint16_t*>(dyn + L.open);
    uint16_t* closel = reinterpret_cast<uint16_t*>(dyn + L.closel);
    uint16_t* start = reinterpret_cast<uint16_t*>(dyn + L.start);
    char* str = reinterpret_cast<char*>(dyn + L.str);

    for (int64_t tree = blockIdx.x; tree < B; tree += gridDim.x) {
        __syncthreads();                          // the previous tree's readers are done
        if (status[tree] != 0) {                  // uniform: invalid vector, nothing decoded
            if (tid == 0 && lengths) lengths[tree] = 0;
            continue;
        }
        // ---- 1. the ancestry rows of this tree ----------------------------------------------
        {
            const int* src = anc_all + (size_t)tree * 3 * k;
            for (int i = tid; i < 3 * k; i += NT) anc[i] = src[i];
        }
        __syncthreads();
        // ---- 2. child pointers and chain weights ------------------------------------------
        for (int c = tid; c <= k; c += NT) {
            dl[c] = (uint16_t)c; dr[c] = (uint16_t)c; al[c] = 0; ar[c] = 0;
            nopen[c] = 0; closel[c] = 0;
        }
        for (int r = tid; r < k; r += NT) {
            const int u = k + 1 + r;
            dl[u] = (uint16_t)anc[3 * r]; dr[u] = (uint16_t)anc[3 * r + 1];
            al[u] = 1; ar[u] = (uint16_t)(1 + dec_digits((unsigned)u));      // "(" ; ")" + label
        }
        __syncthreads();
        int cur = 0;
        for (;;) {      // jump down both chains; leaves are fixed points with weight 0
            const uint16_t* l_in = dl + cur * N; uint16_t* l_out = dl + (cur ^ 1) * N;
            const uint16_t* r_in = dr + cur * N; uint16_t* r_out = dr + (cur ^ 1) * N;
            const uint16_t* a_in = al + cur * N; uint16_t* a_out = al + (cur ^ 1) * N;
            const uint16_t* b_in = ar + cur * N; uint16_t* b_out = ar + (cur ^ 1) * N;
            int live = 0;
            for (int x = tid; x < N; x += NT) {
                const unsigned l = l_in[x], r = r_in[x];
                const unsigned nl = l_in[l], nr = r_in[r];
                a_out[x] = (uint16_t)(a_in[x] + a_in[l]);
                b_out[x] = (uint16_t)(b_in[x] + b_in[r]);
                l_out[x] = (uint16_t)nl; r_out[x] = (uint16_t)nr;
                live |= ((int)nl > k) | ((int)nr > k);
            }
            cur ^= 1;
            if (!__syncthreads_or(live)) break;
        }
        // NOTE: a node whose chain has already reached a leaf keeps adding the leaf's weight 0.
        const uint16_t* lres = dl + cur * N;      // leftmost leaf of every node
        const uint16_t* rres = dr + cur * N;      // rightmost leaf
        const uint16_t* ares = al + cur * N;      // "(" from this node down its first-child chain
        const uint16_t* bres = ar + cur * N;      // bytes of ")label" from this node down its second-child chain
        // ---- 3. per-leaf runs, the leaf list and its weighted ranking ---------------------------
        // the top of a first-child chain is a second child (or the root): it opens the whole run of
        // "(" before its leftmost leaf; the top of a second-child chain is a first child (or the
        // root): it closes the whole run after its rightmost leaf
        for (int r = tid; r < k; r += NT) {
            const int c1 = anc[3 * r], c2 = anc[3 * r + 1];
            if (c2 > k) nopen[lres[c2]] = ares[c2];
            if (c1 > k) closel[rres[c1]] = bres[c1];
            nxt[rres[c1]] = lres[c2];
        }
        if (tid == 0) {
            nxt[rres[root]] = (uint16_t)NW_NONE;
            nopen[lres[root]] = ares[root]; closel[rres[root]] = bres[root];
        }
        __syncthreads();
        const int first_leaf = lres[root];
        for (int c = tid; c <= k; c += NT) {
            const uint32_t w = (c != first_leaf ? 1u : 0u) + nopen[c] + (uint32_t)dec_digits((unsigned)c) + closel[c];
            wleaf[c] = (uint16_t)w; wsum[c] = (uint16_t)w;
        }
        __syncthreads();
        int lc = 0;
        for (;;) {
            const uint16_t* n_in = nxt + lc * n; uint16_t* n_out = nxt + (lc ^ 1) * n;
            const uint16_t* s_in = wsum + lc * n; uint16_t* s_out = wsum + (lc ^ 1) * n;
            int live = 0;
            for (int x = tid; x < n; x += NT) {
                const unsigned a = n_in[x];
                uint32_t s = s_in[x];
                unsigned na = NW_NONE;
                if (a != NW_NONE) { s += s_in[a]; na = n_in[a]; live |= (na != NW_NONE); }
                n_out[x] = (uint16_t)na; s_out[x] = (uint16_t)s;
            }
            lc ^= 1;
            if (!__syncthreads_or(live)) break;
        }
        const uint16_t* sres = wsum + lc * n;     // bytes from the leaf's run to the end (";" excluded)
        const uint32_t total = sres[first_leaf];
        // ---- 4. tokens ---------------------------------------------------------------------------
        for (int c = tid; c <= k; c += NT) {
            uint32_t o = total - sres[c];
            start[c] = (uint16_t)o;
            if (c != first_leaf) str[o++] = ',';
            const uint32_t no = nopen[c];
            for (uint32_t i = 0; i < no; ++i) str[o + i] = '(';
            put_dec(str + o + no, (unsigned)c, dec_digits((unsigned)c));
        }
        __syncthreads();
        for (int r = tid; r < k; r += NT) {
            const int u = k + 1 + r;
            const int leaf = rres[u];
            const int c2 = anc[3 * r + 1];
            const uint32_t below = (c2 > k) ? bres[c2] : 0u;        // closings of the chain under u
            const uint32_t o = start[leaf] + wleaf[leaf] - closel[leaf] + below;
            str[o] = ')';
            put_dec(str + o + 1, (unsigned)u, dec_digits((unsigned)u));
        }
        if (tid == 0) { str[total] = ';'; str[total + 1] = 0; }
        __syncthreads();
        // ---- copy out (the row is out_stride bytes; bytes past the NUL are left untouched) --------
        char* o = out + tree * out_stride;
        const uint32_t nbytes = total + 2;
        if ((reinterpret_cast<uintptr_t>(o) & 15) == 0) {
            const uint32_t nvec = nbytes >> 4;
            for (uint32_t i = tid; i < nvec; i += NT)
                reinterpret_cast<uint4*>(o)[i] = reinterpret_cast<const uint4*>(str)[i];
            for (uint32_t i = (nvec << 4) + tid; i < nbytes; i += NT) o[i] = str[i];
        } else {
            for (uint32_t i = tid; i < nbytes; i += NT) o[i] = str[i];
        }
        if (tid == 0 && lengths) lengths[tree] = (int64_t)total + 1;
    }
}

size_t newick_max_bytes(int64_t k) { return k < 0 ? 0 : newick_bytes(k) + 1; }

// workspace: [ status scratch (B int32) | int32 v (B*k, only for int64 input) | int32 ancestry (B*k*3) ]
struct NewickWs { size_t status, v32, anc, total; };
static NewickWs newick_ws(int64_t B, int64_t k) {
    NewickWs w;
    size_t o = 0;
    w.status = o; o = nw_al16(o + 4 * (size_t)B);
    w.v32 = o; o = nw_al16(o + 4 * (size_t)B * (size_t)k);
    w.anc = o; o = nw_al16(o + 12 * (size_t)B * (size_t)k);
    w.total = o;
    return w;
}
size_t newick_workspace_bytes(int64_t B, int64_t k) { return (B <= 0 || k <= 0) ? 0 : newick_ws(B, k).total; }

cudaError_t launch_to_newick(const void* v, int idx_bytes, int64_t B, int64_t k, char* out, int64_t out_stride,
                             int64_t* lengths, int32_t* status, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    if (k < 1 || k > NW_K_MAX) return cudaErrorNotSupported;
    if (out_stride < (int64_t)newick_max_bytes(k)) return cudaErrorInvalidValue;
    const NewickWs W = newick_ws(B, k);
    if (!ws || ws_bytes < W.total) return cudaErrorInvalidValue;
    unsigned char* wsb = static_cast<unsigned char*>(ws);
    int32_t* st = reinterpret_cast<int32_t*>(wsb + W.status);
    int* anc = reinterpret_cast<int*>(wsb + W.anc);
    const void* vin = v;
    cudaError_t e;
    if (idx_bytes == 8) {
        e = launch_narrow_v(static_cast<const long long*>(v), B * k, reinterpret_cast<int*>(wsb + W.v32), stream);
        if (e != cudaSuccess) return e;
        vin = wsb + W.v32;
    } else if (idx_bytes != 4) {
        return cudaErrorInvalidValue;
    }
    e = launch_decode(MODE_ANCESTRY, vin, 4, B, k, anc, st, nullptr, 0, stream, 0);     // k <= 1023: no workspace
    if (e != cudaSuccess) return e;
    const size_t smem = newick_layout((int)k).total;
    e = cudaFuncSetAttribute(newick_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int threads = NW_THREADS;      // measured at n = 100: 32-thread CTAs are no faster (32 CTAs/SM cap)
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, newick_small_kernel, threads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const int64_t cap = (int64_t)sm_count() * per_sm;
    newick_small_kernel<<<(int)(B < cap ? B : cap), threads, smem, stream>>>(anc, B, (int)k, out, out_stride, lengths, st);
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (status) e = cudaMemcpyAsync(status, st, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, stream);
    return e;
}

}  // namespace p2v
