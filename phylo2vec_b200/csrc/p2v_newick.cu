// Batched to_newick / from_newick for Phylo2Vec vectors (topology only), sm_100a.
//
// Replaces (reference, paths relative to /root/reference):
//   vector::convert::to_newick / build_newick / prepare_cache   phylo2vec/src/vector/convert.rs:341-397
//   PyO3 to_newick (vector arm)                                    py-phylo2vec/src/lib.rs:90-96
//   newick::parse, vector::convert::from_newick                 newick/mod.rs:73-127, vector/convert.rs:268-278
//   (see the second half of this file)
//
// to_newick.  The reference concatenates per-leaf string caches pair by pair; the result is the
// Newick string of the ancestry tree with children in (c1, c2) order and every node labelled:
//   S(leaf) = dec(leaf),  S(u) = "(" S(c1) "," S(c2) ")" dec(u),  out = S(root) ";".
// The batch is first decoded by the batched decode kernel (int32, workspace).
//   * k <= 4095: newick_warp_kernel, one warp per tree (described where it is defined);
//   * larger trees: newick_kernel<uint32_t>, one 1024-thread CTA per tree with its tables in a
//     per-CTA slice of the workspace (L2 resident) and the tokens written straight into the
//     output row:
//       1. pointer jumping down the first-child and second-child pointers gives every node its
//          leftmost / rightmost leaf, the number of "(" a first-child chain opens and the bytes a
//          second-child chain closes (")" + label per node);
//       2. every internal node links the last leaf of its first child to the first leaf of its
//          second child: the leaves form a list in string order; weighted list ranking (weight =
//          bytes emitted around the leaf: ",", the "(" run, its label, the ")label" run) gives
//          every leaf its byte offset;
//       3. leaves write [","] "(((" label, internal nodes write ")" label.
//     (The template's uint16_t form -- tables, ancestry copy and string staging in shared memory
//     -- was the first version for small trees; it is no longer instantiated.)
// The matrix form (branch lengths need the reference's shortest-round-trip float formatting) is
// not built.
#include <cstdlib>

#include "p2v_common.cuh"
#include "p2v_internal.h"

namespace p2v {

constexpr int NW_THREADS_MAX = 1024;
constexpr size_t NW_SMEM_MAX = 227 * 1024;

__host__ __device__ inline int dec_digits(unsigned x) {
    return 1 + (x >= 10u) + (x >= 100u) + (x >= 1000u) + (x >= 10000u) + (x >= 100000u) + (x >= 1000000u) +
           (x >= 10000000u);
}

// exact length of the string of a tree with k + 1 leaves (without the terminating NUL)
__host__ __device__ inline size_t newick_bytes(int64_t k) {
    size_t total = 3 * (size_t)k + 1;                  // "(", ",", ")" per internal node and the ";"
    const int64_t top = 2 * k;                         // labels 0 .. 2k: one digit each, plus one per power of ten passed
    total += (size_t)(top + 1);
    for (int64_t p = 10; p <= top; p *= 10) total += (size_t)(top - p + 1);
    return total;
}

struct NewickLayout {
    size_t anc, dl, dr, al, ar, nxt, wsum, wleaf, open, closel, start, str, total;
    size_t strcap;
};
__host__ __device__ inline size_t nw_al16(size_t x) { return (x + 15) & ~(size_t)15; }
// es = bytes per table entry: 2 in shared memory (ancestry copy and string staging included),
// 4 in the workspace (the ancestry is read in place and the string is written in place)
__host__ __device__ inline NewickLayout newick_layout(int64_t k, size_t es) {
    NewickLayout L;
    const size_t n = (size_t)k + 1, N = 2 * (size_t)k + 1;
    const bool staged = es == 2;
    size_t o = 0;
    L.anc = o; o = nw_al16(o + (staged ? 4 * 3 * (size_t)k : 0));
    L.dl = o; o = nw_al16(o + es * 2 * N);           // first-child jump pointers, two buffers
    L.dr = o; o = nw_al16(o + es * 2 * N);           // second-child jump pointers
    L.al = o; o = nw_al16(o + es * 2 * N);           // "(" opened along the first-child chain
    L.ar = o; o = nw_al16(o + es * 2 * N);           // bytes closed along the second-child chain
    L.nxt = o; o = nw_al16(o + es * 2 * n);          // leaf list, two buffers
    L.wsum = o; o = nw_al16(o + es * 2 * n);         // bytes from this leaf to the end of the string
    L.wleaf = o; o = nw_al16(o + es * n);            // bytes emitted around one leaf
    L.open = o; o = nw_al16(o + es * n);
    L.closel = o; o = nw_al16(o + es * n);
    L.start = o; o = nw_al16(o + es * n);
    L.strcap = staged ? nw_al16(newick_bytes(k) + 2) : 0;
    L.str = o; o = nw_al16(o + L.strcap);
    L.total = o;
    return L;
}
__device__ __forceinline__ void put_dec(char* p, unsigned x, int digits) {
    for (int i = digits - 1; i >= 0; --i) { p[i] = (char)('0' + x % 10u); x /= 10u; }
}

// T = T: tables, ancestry copy and string staging in dynamic shared memory;
// T = uint32_t: tables in this CTA's slice of the workspace, ancestry and string in place.
template <typename T>
__global__ void __launch_bounds__(NW_THREADS_MAX)
newick_kernel(const int* __restrict__ anc_all, int64_t B, int k, char* __restrict__ out, int64_t out_stride,
              int64_t* __restrict__ lengths, const int32_t* __restrict__ status, unsigned char* __restrict__ tables,
              size_t tables_stride) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    constexpr bool STAGED = sizeof(T) == 2;
    constexpr unsigned NW_NONE = (unsigned)(T)~(T)0;
    unsigned char* dyn = STAGED ? dyn_smem : tables + (size_t)blockIdx.x * tables_stride;
    const NewickLayout L = newick_layout(k, sizeof(T));
    const int n = k + 1, N = 2 * k + 1, root = 2 * k;
    const int tid = threadIdx.x;
    const int NT = blockDim.x;
    T* dl = reinterpret_cast<T*>(dyn + L.dl);
    T* dr = reinterpret_cast<T*>(dyn + L.dr);
    T* al = reinterpret_cast<T*>(dyn + L.al);
    T* ar = reinterpret_cast<T*>(dyn + L.ar);
    T* nxt = reinterpret_cast<T*>(dyn + L.nxt);
    T* wsum = reinterpret_cast<T*>(dyn + L.wsum);
    T* wleaf = reinterpret_cast<T*>(dyn + L.wleaf);
    T* nopen = reinterpret_cast<T*>(dyn + L.open);
    T* closel = reinterpret_cast<T*>(dyn + L.closel);
    T* start = reinterpret_cast<T*>(dyn + L.start);

    for (int64_t tree = blockIdx.x; tree < B; tree += gridDim.x) {
        __syncthreads();                          // the previous tree's readers are done
        if (status[tree] != 0) {                  // uniform: invalid vector, nothing decoded
            if (tid == 0 && lengths) lengths[tree] = 0;
            continue;
        }
        // ---- 1. the ancestry rows of this tree ----------------------------------------------
        const int* anc = anc_all + (size_t)tree * 3 * k;
        char* str = out + tree * out_stride;
        if (STAGED) {
            int* sanc = reinterpret_cast<int*>(dyn + L.anc);
            for (int i = tid; i < 3 * k; i += NT) sanc[i] = anc[i];
            anc = sanc;
            str = reinterpret_cast<char*>(dyn + L.str);
            __syncthreads();
        }
        // ---- 2. child pointers and chain weights ------------------------------------------
        for (int c = tid; c <= k; c += NT) {
            dl[c] = (T)c; dr[c] = (T)c; al[c] = 0; ar[c] = 0;
            nopen[c] = 0; closel[c] = 0;
        }
        for (int r = tid; r < k; r += NT) {
            const int u = k + 1 + r;
            dl[u] = (T)anc[3 * r]; dr[u] = (T)anc[3 * r + 1];
            al[u] = 1; ar[u] = (T)(1 + dec_digits((unsigned)u));      // "(" ; ")" + label
        }
        __syncthreads();
        int cur = 0;
        for (;;) {      // jump down both chains; leaves are fixed points with weight 0
            const T* l_in = dl + cur * N; T* l_out = dl + (cur ^ 1) * N;
            const T* r_in = dr + cur * N; T* r_out = dr + (cur ^ 1) * N;
            const T* a_in = al + cur * N; T* a_out = al + (cur ^ 1) * N;
            const T* b_in = ar + cur * N; T* b_out = ar + (cur ^ 1) * N;
            int live = 0;
            for (int x = tid; x < N; x += NT) {
                const unsigned l = l_in[x], r = r_in[x];
                const unsigned nl = l_in[l], nr = r_in[r];
                a_out[x] = (T)(a_in[x] + a_in[l]);
                b_out[x] = (T)(b_in[x] + b_in[r]);
                l_out[x] = (T)nl; r_out[x] = (T)nr;
                live |= ((int)nl > k) | ((int)nr > k);
            }
            cur ^= 1;
            if (!__syncthreads_or(live)) break;
        }
        // NOTE: a node whose chain has already reached a leaf keeps adding the leaf's weight 0.
        const T* lres = dl + cur * N;      // leftmost leaf of every node
        const T* rres = dr + cur * N;      // rightmost leaf
        const T* ares = al + cur * N;      // "(" from this node down its first-child chain
        const T* bres = ar + cur * N;      // bytes of ")label" from this node down its second-child chain
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
            nxt[rres[root]] = (T)NW_NONE;
            nopen[lres[root]] = ares[root]; closel[rres[root]] = bres[root];
        }
        __syncthreads();
        const int first_leaf = lres[root];
        for (int c = tid; c <= k; c += NT) {
            const uint32_t w = (c != first_leaf ? 1u : 0u) + nopen[c] + (uint32_t)dec_digits((unsigned)c) + closel[c];
            wleaf[c] = (T)w; wsum[c] = (T)w;
        }
        __syncthreads();
        int lc = 0;
        for (;;) {
            const T* n_in = nxt + lc * n; T* n_out = nxt + (lc ^ 1) * n;
            const T* s_in = wsum + lc * n; T* s_out = wsum + (lc ^ 1) * n;
            int live = 0;
            for (int x = tid; x < n; x += NT) {
                const unsigned a = n_in[x];
                uint32_t s = s_in[x];
                unsigned na = NW_NONE;
                if (a != NW_NONE) { s += s_in[a]; na = n_in[a]; live |= (na != NW_NONE); }
                n_out[x] = (T)na; s_out[x] = (T)s;
            }
            lc ^= 1;
            if (!__syncthreads_or(live)) break;
        }
        const T* sres = wsum + lc * n;     // bytes from the leaf's run to the end (";" excluded)
        const uint32_t total = sres[first_leaf];
        // ---- 4. tokens ---------------------------------------------------------------------------
        for (int c = tid; c <= k; c += NT) {
            uint32_t o = total - sres[c];
            start[c] = (T)o;
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
        if (STAGED) {
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
        }
        if (tid == 0 && lengths) lengths[tree] = (int64_t)total + 1;
    }
}

// --------------------------------------------------------------------------------------------
// k <= 4095: one warp per tree, no block-wide barrier.  S(u) = "(" S(c1) "," S(c2) ")" u, so with
// len[u] = bytes of S(u) and off[u] = where S(u) starts, every internal node places its own tokens
// and the labels of its leaf children:
//   1. bottom-up over the ancestry rows, 32 at a time: len[u] = 3 + len[c1] + len[c2] + digits(u);
//      a child that is a row of the same batch is picked up with a shuffle once it is known
//      (rounds = dependency depth inside the batch); every internal child records its parent row;
//   2. top-down, 32 rows at a time from the root: off[u] = off[parent] + 1 (first child) or
//      + 2 + len[c1(parent)]; parents inside the batch are followed by pointer doubling
//      (5 shuffle rounds, whatever the shape);
//   3. tokens; the string is assembled in shared memory and copied out with 128-bit stores.
// Input: the first two ancestry columns as int32 (MODE_ANC2).  Per tree: children (2 x uint16 per
// row), len, off and parent (uint16 per row), the string: 21 KB at n = 1000.
// --------------------------------------------------------------------------------------------
constexpr int NWW_K_MAX = 4095;
__host__ __device__ inline size_t newick_warp_bytes(int k) {
    return nw_al16(4 * (size_t)k) + 3 * nw_al16(2 * (size_t)k) + nw_al16(newick_bytes(k) + 2);
}

__device__ __forceinline__ void put_dec4(char* p, unsigned x, int digits) {      // x < 10000
    const unsigned a = x / 10u, b = a / 10u, c = b / 10u;
    const char d0 = (char)('0' + (x - a * 10u)), d1 = (char)('0' + (a - b * 10u)), d2 = (char)('0' + (b - c * 10u)),
               d3 = (char)('0' + c);
    if (digits == 1) { p[0] = d0; }
    else if (digits == 2) { p[0] = d1; p[1] = d0; }
    else if (digits == 3) { p[0] = d2; p[1] = d1; p[2] = d0; }
    else { p[0] = d3; p[1] = d2; p[2] = d1; p[3] = d0; }
}

__global__ void __launch_bounds__(128)
newick_warp_kernel(const int* __restrict__ anc_all, int64_t B, int k, char* __restrict__ out, int64_t out_stride,
                   int64_t* __restrict__ lengths, const int32_t* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    unsigned char* mine = dyn + (size_t)warp * newick_warp_bytes(k);
    const size_t tb = nw_al16(2 * (size_t)k);
    uint32_t* ch = reinterpret_cast<uint32_t*>(mine);                                   // c1 | c2 << 16 of row r
    uint16_t* len = reinterpret_cast<uint16_t*>(mine + nw_al16(4 * (size_t)k));         // by row
    uint16_t* off = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(len) + tb);
    uint16_t* par = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(off) + tb);   // parent row << 1 | second child
    char* str = reinterpret_cast<char*>(reinterpret_cast<unsigned char*>(par) + tb);
    const int nb = (k + 31) >> 5;

    for (int64_t tree = (int64_t)blockIdx.x * nwarps + warp; tree < B; tree += (int64_t)gridDim.x * nwarps) {
        if (status[tree] != 0) {                  // invalid vector, nothing decoded
            if (lane == 0 && lengths) lengths[tree] = 0;
            continue;
        }
        const int2* anc = reinterpret_cast<const int2*>(anc_all + (size_t)tree * 2 * k);
        for (int r = lane; r < k; r += 32) {
            const int2 x = anc[r];
            ch[r] = (uint32_t)x.x | ((uint32_t)x.y << 16);
        }
        __syncwarp();
        // ---- 1. len, bottom-up ---------------------------------------------------------------------
        for (int b = 0; b < nb; ++b) {
            const int base = b * 32, r = base + lane;
            const bool act = r < k;
            const uint32_t cc = act ? ch[r] : 0u;
            const int c1 = (int)(cc & 0xFFFFu), c2 = (int)(cc >> 16);
            const int r1 = c1 - k - 1, r2 = c2 - k - 1;             // child rows (negative: a leaf)
            int l1 = 0, l2 = 0, d1 = -1, d2 = -1;                  // d: lane of a child row inside this batch
            if (r1 < 0) l1 = dec_digits((unsigned)c1); else if (r1 < base) l1 = len[r1]; else d1 = r1 - base;
            if (r2 < 0) l2 = dec_digits((unsigned)c2); else if (r2 < base) l2 = len[r2]; else d2 = r2 - base;
            if (act && r1 >= 0) par[r1] = (uint16_t)(r << 1);
            if (act && r2 >= 0) par[r2] = (uint16_t)((r << 1) | 1);
            const int own = 3 + dec_digits((unsigned)(k + 1 + r));
            // Rows hanging off one another form chains inside the batch (the usual case: a row
            // continues the subtree of the row before it).  A row with one open child follows that
            // link; a chain is summed by pointer doubling in one round, so the number of rounds is the
            // nesting depth through rows with two open children, not the length of the chains.
            uint32_t mylen = act ? 0u : 1u;                         // 0 = not known yet
            for (;;) {
                const bool one = mylen == 0u && ((d1 ^ d2) < 0);    // exactly one child still open
                uint32_t acc = mylen ? mylen : (uint32_t)(own + l1 + l2);      // open children count 0 so far
                bool fin = mylen != 0u || (d1 & d2) < 0;
                int ptr = one ? (d1 >= 0 ? d1 : d2) : lane;         // two open children: the row blocks its chain
#pragma unroll
                for (int it = 0; it < 5; ++it) {
                    const uint32_t a = __shfl_sync(FULL, acc | (fin ? 0x80000000u : 0u), ptr);
                    const int pp = __shfl_sync(FULL, ptr, ptr);
                    if (!fin && ptr != lane) { acc += a & 0x7FFFFFFFu; fin = (a >> 31) != 0u; ptr = pp; }
                }
                if (fin) mylen = acc;
                if (__all_sync(FULL, mylen != 0u)) break;
                const uint32_t v1 = __shfl_sync(FULL, mylen, d1), v2 = __shfl_sync(FULL, mylen, d2);
                if (d1 >= 0 && v1) { l1 = (int)v1; d1 = -1; }
                if (d2 >= 0 && v2) { l2 = (int)v2; d2 = -1; }
            }
            if (act) len[r] = (uint16_t)mylen;
            __syncwarp();
        }
        const uint32_t total = len[k - 1];
        // ---- 2. off, top-down ------------------------------------------------------------------------
        for (int b = nb - 1; b >= 0; --b) {
            const int base = b * 32, r = base + lane;
            const bool act = r < k;
            uint32_t acc = 0;
            int ptr = lane;                       // lane of the parent row while it is inside the batch and not final
            bool fin = true;
            if (act && r != k - 1) {
                const uint32_t p = par[r];
                const int pr = (int)(p >> 1);
                uint32_t delta = 1;
                if (p & 1u) {
                    const int pc1 = (int)(ch[pr] & 0xFFFFu);
                    delta = 2u + (pc1 <= k ? (uint32_t)dec_digits((unsigned)pc1) : (uint32_t)len[pc1 - k - 1]);
                }
                acc = delta;
                if (pr >= base + 32) acc += off[pr];               // a later batch: final
                else { ptr = pr - base; fin = false; }
            }
#pragma unroll
            for (int it = 0; it < 5; ++it) {
                const uint32_t a = __shfl_sync(FULL, acc | (fin ? 0x80000000u : 0u), ptr);     // offsets are below 2^16
                const int pp = __shfl_sync(FULL, ptr, ptr);
                if (!fin) { acc += a & 0x7FFFFFFFu; fin = (a >> 31) != 0u; ptr = pp; }
            }
            if (act) off[r] = (uint16_t)acc;
            __syncwarp();
        }
        // ---- 3. tokens -----------------------------------------------------------------------------------
        for (int r = lane; r < k; r += 32) {
            const uint32_t cc = ch[r];
            const int c1 = (int)(cc & 0xFFFFu), c2 = (int)(cc >> 16);
            const uint32_t o = off[r];
            uint32_t l1, l2;
            str[o] = '(';
            if (c1 <= k) { l1 = (uint32_t)dec_digits((unsigned)c1); put_dec4(str + o + 1, (unsigned)c1, (int)l1); }
            else l1 = len[c1 - k - 1];
            str[o + 1 + l1] = ',';
            if (c2 <= k) { l2 = (uint32_t)dec_digits((unsigned)c2); put_dec4(str + o + 2 + l1, (unsigned)c2, (int)l2); }
            else l2 = len[c2 - k - 1];
            const uint32_t p = o + 2 + l1 + l2;
            str[p] = ')';
            const unsigned u = (unsigned)(k + 1 + r);
            put_dec4(str + p + 1, u, dec_digits(u));
        }
        if (lane == 0) { str[total] = ';'; str[total + 1] = 0; }
        __syncwarp();
        // ---- copy out (the row is out_stride bytes; bytes past the NUL are left untouched) --------
        char* o = out + tree * out_stride;
        const uint32_t nbytes = total + 2;
        if ((reinterpret_cast<uintptr_t>(o) & 15) == 0) {
            const uint32_t nvec = nbytes >> 4;
            for (uint32_t i = lane; i < nvec; i += 32)
                reinterpret_cast<uint4*>(o)[i] = reinterpret_cast<const uint4*>(str)[i];
            for (uint32_t i = (nvec << 4) + lane; i < nbytes; i += 32) o[i] = str[i];
        } else {
            for (uint32_t i = lane; i < nbytes; i += 32) o[i] = str[i];
        }
        if (lane == 0 && lengths) lengths[tree] = (int64_t)total + 1;
        __syncwarp();
    }
}

size_t newick_max_bytes(int64_t k) { return k < 0 ? 0 : newick_bytes(k) + 1; }

// CTAs of the workspace variant (each owns a slice of tables)
static int64_t newick_teams(int64_t B, int64_t k) {
    const int64_t cap = (int64_t)sm_count() * (k <= 16384 ? 2 : 1);
    return B < cap ? B : cap;
}

// workspace: [ status scratch (B int32) | int32 v (B*k, only for int64 input) | int32 ancestry (B*k*3)
//              | decode workspace | per-CTA tables (large trees only) ]
struct NewickWs { size_t status, v32, anc, dec, dec_bytes, tables, tables_stride, total; };
static NewickWs newick_ws(int64_t B, int64_t k) {
    NewickWs w;
    size_t o = 0;
    w.status = o; o = nw_al16(o + 4 * (size_t)B);
    w.v32 = o; o = nw_al16(o + 4 * (size_t)B * (size_t)k);
    w.anc = o; o = nw_al16(o + 12 * (size_t)B * (size_t)k);
    w.dec_bytes = decode_workspace_bytes(B, k);
    w.dec = o; o = nw_al16(o + w.dec_bytes);
    w.tables_stride = k <= NWW_K_MAX ? 0 : nw_al16(newick_layout(k, 4).total);
    w.tables = o; o = nw_al16(o + w.tables_stride * (size_t)newick_teams(B, k));
    w.total = o;
    return w;
}
size_t newick_workspace_bytes(int64_t B, int64_t k) { return (B <= 0 || k <= 0) ? 0 : newick_ws(B, k).total; }

cudaError_t launch_to_newick(const void* v, int idx_bytes, int64_t B, int64_t k, char* out, int64_t out_stride,
                             int64_t* lengths, int32_t* status, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    if (k < 1 || k > LARGE_K_MAX) return cudaErrorNotSupported;
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
    e = launch_decode(k <= NWW_K_MAX ? MODE_ANC2 : MODE_ANCESTRY, vin, 4, B, k, anc, st, W.dec_bytes ? wsb + W.dec : nullptr,
                      W.dec_bytes, stream, 0);
    if (e != cudaSuccess) return e;
    if (k <= NWW_K_MAX) {
        // warps (= trees in flight) per CTA: enough small CTAs to fill the SM's warp slots
        const size_t per = newick_warp_bytes((int)k);
        int wpc = 1, best = 0;
        for (int w = 4; w >= 1; w >>= 1) {
            if (per * w > NW_SMEM_MAX) continue;
            size_t ctas = ((size_t)228 * 1024) / (per * w + 1024);
            if (ctas > 32) ctas = 32;
            if (ctas > (size_t)(64 / w)) ctas = 64 / w;
            if ((int)ctas * w > best) { best = (int)ctas * w; wpc = w; }
        }
        const size_t smem = per * wpc;
        e = cudaFuncSetAttribute(newick_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, newick_warp_kernel, wpc * 32, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        const int64_t cap = (int64_t)sm_count() * per_sm;
        const int64_t want = (B + wpc - 1) / wpc;
        newick_warp_kernel<<<(int)(want < cap ? want : cap), wpc * 32, smem, stream>>>(anc, B, (int)k, out, out_stride, lengths, st);
    } else {
        newick_kernel<uint32_t><<<(int)newick_teams(B, k), NW_THREADS_MAX, 0, stream>>>(
            anc, B, (int)k, out, out_stride, lengths, st, wsb + W.tables, W.tables_stride);
    }
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (status) e = cudaMemcpyAsync(status, st, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, stream);
    return e;
}


// ======================================================================================
// from_newick (vectors)
//
// Replaces (reference):
//   newick::parse                                   phylo2vec/src/newick/mod.rs:73-127
//   vector::convert::from_newick                    phylo2vec/src/vector/convert.rs:268-278
//   order_cherries_no_parents                       phylo2vec/src/vector/convert.rs:534-573
//   PyO3 to_vector                                  py-phylo2vec/src/lib.rs:98-101
//
// The reference walks the string with a stack: a number pushes a node, ")" pops two children and
// pushes the label that follows (or, without parent labels, the smaller child).  The nodes
// therefore appear in post-order, and with S(i) = stack height after node i (leaf +1, internal
// -1): the second child of internal node i is node i-1, the first child is the latest node
// j < i with S(j) = S(i).  One CTA per tree:
//   1. the string is staged in shared memory; every thread classifies a contiguous piece
//      (leaf token, ")" with its label, punctuation) and a block scan numbers the nodes and
//      gives S; the labels are parsed where the tokens start;
//   2. one warp sweeps the nodes 32 at a time with a table "latest node of height h"
//      (same-batch matches through MATCH.ANY) and writes row r = (c1, p), (c2, p) of the
//      cherry list in string order -- newick::parse's output as an edge list;
//   3. the batched encode kernel (from_edges: order_cherries + build_cherries + build_vector)
//      turns the rows into v.
// Strings without parent labels take a second kernel (newick_parse_np_kernel) that gives every
// internal node the label the reference's ordering implies.  order_cherries_no_parents sorts the
// cherries by their "sister" key, descending and stable; along the chain of cherries that share a
// smaller child m the key is the running minimum of the larger children, i.e. the second-smallest
// leaf label of the cherry's subtree.  A subtree is a contiguous range of leaves in string order,
// so the key is three range-minimum queries on a sparse table over the leaf labels; equal keys
// only occur along one chain, in row order.  The cherry with sorted position q is labelled
// k + 1 + q and the tree goes through the same from_edges encoder.
// Anything that is not a well-formed binary Newick of k + 1 leaves gets P2V_TREE_MALFORMED.
// ======================================================================================

constexpr int NP_NOPARENTS = -3;        // first kernel -> second kernel: ")" without a label seen
constexpr int NP_MALFORMED = -2;        // P2V_TREE_MALFORMED
constexpr uint32_t NP_INTERNAL = 0x80000000u;

struct ParseLayout { size_t str, lab, sh, last, c1a, first, leaf, sister, gt, sparse, total; int levels; };
// noparents: the extra tables of the second kernel
__host__ __device__ inline ParseLayout parse_layout(int64_t k, bool staged, bool noparents) {
    ParseLayout L;
    const size_t N = 2 * (size_t)k + 1, n = (size_t)k + 1;
    size_t o = 0;
    L.str = o; o = nw_al16(o + (staged ? newick_bytes(k) + 16 : 0));
    L.lab = o; o = nw_al16(o + 4 * N);
    L.sh = o; o = nw_al16(o + 4 * N);
    L.last = o; o = nw_al16(o + 4 * ((size_t)k + 3));
    int lv = 1;
    while (((size_t)1 << lv) <= n) ++lv;          // levels 0 .. lv-1, 2^(lv-1) <= n
    L.levels = lv;
    L.c1a = o; o = nw_al16(o + (noparents ? 4 * N : 0));
    L.first = o; o = nw_al16(o + (noparents ? 4 * N : 0));
    L.leaf = o; o = nw_al16(o + (noparents ? 4 * n : 0));
    L.sister = o; o = nw_al16(o + (noparents ? 4 * n : 0));
    L.gt = o; o = nw_al16(o + (noparents ? 4 * (n + 1) : 0));
    L.sparse = o; o = nw_al16(o + (noparents ? 4 * n * (size_t)lv : 0));
    L.total = o;
    return L;
}
static bool parse_fits_smem(int64_t k, bool noparents) {
    return k <= 8192 && parse_layout(k, true, noparents).total <= 200 * 1024;
}

__device__ __forceinline__ bool np_digit(char c) { return c >= '0' && c <= '9'; }

struct ParseShared { int cnt[32], del[32]; int len, bad; };

// Length of tree's string (0 = unusable), staged copy, node labels (lab: label | NP_INTERNAL, in
// post-order) and stack heights (sh).  LABELLED: every ")" carries a label (a ")" without one
// makes the tree NP_NOPARENTS); otherwise no ")" may carry one.  Returns the tree's verdict
// (uniform over the CTA): 0, NP_NOPARENTS or NP_MALFORMED.
template <bool LABELLED, bool STAGED>
__device__ __forceinline__ int newick_tokenize(const char* __restrict__ src, int64_t in_stride, const int64_t* __restrict__ lengths,
                                               int64_t tree, int k, int maxlen, char* sstr, uint32_t* lab, int* sh,
                                               ParseShared& P) {
    const int N = 2 * k + 1;
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = NT >> 5;
    __syncthreads();                          // the previous tree is done with the tables
    if (tid == 0) { P.bad = 0; P.len = maxlen + 1; }
    __syncthreads();
    int len;
    if (lengths) {
        const int64_t l64 = lengths[tree];
        len = (l64 < 1 || l64 > maxlen || l64 > in_stride) ? 0 : (int)l64;    // never read past the row
    } else {                                  // NUL-terminated inside the row
        const int lim = (int)(in_stride < (int64_t)maxlen + 1 ? in_stride : (int64_t)maxlen + 1);
        for (int i = tid; i < lim; i += NT) if (src[i] == 0) { atomicMin(&P.len, i); break; }
        __syncthreads();
        len = (P.len < 1 || P.len > maxlen) ? 0 : P.len;
    }
    const char* s = src;
    if (STAGED) {
        for (int i = tid; i < len; i += NT) sstr[i] = src[i];
        s = sstr;
        __syncthreads();
    }
    // classify, count, scan
    const int CH = (len + NT - 1) / NT;
    const int i0 = min(len, tid * CH), i1 = min(len, i0 + CH);
    int cnt = 0, del = 0, bad = (len == 0 || s[len - 1] != ';') ? NP_MALFORMED : 0;    // node_substr needs the ";"
    for (int i = i0; i < i1; ++i) {
        const char c = s[i];
        const char prev = i > 0 ? s[i - 1] : '\0';
        if (c == ')') { ++cnt; --del; }
        else if (np_digit(c)) { if (prev == '(' || prev == ',') { ++cnt; ++del; } else if (i == 0) bad = NP_MALFORMED; }
        else if (c == '(' || c == ',') {}
        else if (c == ';') { if (i != len - 1) bad = NP_MALFORMED; }
        else bad = NP_MALFORMED;
    }
    int icnt = cnt, idel = del;               // inclusive scans over the threads
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int a = __shfl_up_sync(FULL, icnt, d), b = __shfl_up_sync(FULL, idel, d);
        if (lane >= d) { icnt += a; idel += b; }
    }
    if (lane == 31) { P.cnt[warp] = icnt; P.del[warp] = idel; }
    __syncthreads();
    if (warp == 0) {
        int a = lane < nwarp ? P.cnt[lane] : 0, b = lane < nwarp ? P.del[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int x = __shfl_up_sync(FULL, a, d), y = __shfl_up_sync(FULL, b, d);
            if (lane >= d) { a += x; b += y; }
        }
        P.cnt[lane] = a; P.del[lane] = b;
    }
    __syncthreads();
    const int tot_nodes = P.cnt[nwarp - 1], tot_del = P.del[nwarp - 1];
    if (tot_nodes != N || tot_del != 1) bad = bad ? bad : NP_MALFORMED;
    int idx = icnt - cnt + (warp ? P.cnt[warp - 1] : 0);
    int S = idel - del + (warp ? P.del[warp - 1] : 0);
    if (!bad) {
        for (int i = i0; i < i1; ++i) {
            const char c = s[i];
            int from;
            uint32_t kind;
            if (c == ')') { from = i + 1; kind = NP_INTERNAL; --S; }
            else if (np_digit(c) && (s[i - 1] == '(' || s[i - 1] == ',')) { from = i; kind = 0; ++S; }
            else continue;
            uint32_t val = 0;
            int nd = 0;
            while (from + nd < len && nd < 9 && np_digit(s[from + nd])) { val = val * 10u + (uint32_t)(s[from + nd] - '0'); ++nd; }
            if (kind && !LABELLED) { if (nd != 0) bad = NP_MALFORMED; }           // labelled and unlabelled ")" mixed
            else if (nd == 0) bad = NP_NOPARENTS;                                 // ")" without a parent label
            else if (nd > 8 || val > (kind ? 2u * (uint32_t)k : (uint32_t)k)) bad = bad ? bad : NP_MALFORMED;
            if (S < 1) bad = NP_MALFORMED;
            lab[idx] = val | kind; sh[idx] = S;
            ++idx;
        }
    }
    if (bad) atomicMax(&P.bad, bad == NP_MALFORMED ? 2 : 1);      // malformed wins
    __syncthreads();
    return P.bad == 0 ? 0 : (P.bad == 2 ? NP_MALFORMED : NP_NOPARENTS);
}

template <typename OutT, bool STAGED>
__global__ void __launch_bounds__(NW_THREADS_MAX)
newick_parse_kernel(const char* __restrict__ in, int64_t in_stride, const int64_t* __restrict__ lengths, int64_t B, int k,
                    OutT* __restrict__ edges, int32_t* __restrict__ pstatus, unsigned char* __restrict__ tables,
                    size_t tables_stride, int maxlen) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ ParseShared P;
    unsigned char* dyn = STAGED ? dyn_smem : tables + (size_t)blockIdx.x * tables_stride;
    const ParseLayout L = parse_layout(k, STAGED, false);
    const int N = 2 * k + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* lab = reinterpret_cast<uint32_t*>(dyn + L.lab);      // label | NP_INTERNAL, by node (post-order)
    int* sh = reinterpret_cast<int*>(dyn + L.sh);                  // stack height after the node
    int* last = reinterpret_cast<int*>(dyn + L.last);              // latest node of height h

    for (int64_t tree = blockIdx.x; tree < B; tree += gridDim.x) {
        OutT* erow = edges + (size_t)tree * 4 * (size_t)k;
        const int verdict = newick_tokenize<true, STAGED>(in + tree * in_stride, in_stride, lengths, tree, k, maxlen,
                                                          reinterpret_cast<char*>(dyn + L.str), lab, sh, P);
        if (verdict) {                            // uniform
            if (threadIdx.x == 0) {
                pstatus[tree] = verdict;
                erow[0] = (OutT)-1; erow[1] = (OutT)-1;           // the encoder rejects the tree as well
            }
            continue;
        }
        if (threadIdx.x == 0) pstatus[tree] = 0;
        // children of every internal node: one warp, 32 nodes at a time
        if (warp == 0) {
            for (int base = 0; base < N; base += 32) {
                const int node = base + lane;
                const bool act = node < N;
                const uint32_t lb = act ? lab[node] : 0u;
                const int Sv = act ? sh[node] : -1 - lane;
                const uint32_t peers = __match_any_sync(FULL, Sv);
                if (act && (lb & NP_INTERNAL)) {
                    const uint32_t lower = peers & lanes_lt(lane);
                    const int j = lower ? base + 31 - __clz(lower) : last[Sv];
                    const uint32_t c1 = lab[j] & ~NP_INTERNAL, c2 = lab[node - 1] & ~NP_INTERNAL, pl = lb & ~NP_INTERNAL;
                    const int r = ((node + 1 - Sv) >> 1) - 1;          // internal nodes before this one
                    OutT* o = erow + 4 * (size_t)r;
                    o[0] = (OutT)c1; o[1] = (OutT)pl; o[2] = (OutT)c2; o[3] = (OutT)pl;
                }
                __syncwarp();
                if (act && (peers & lanes_gt(lane)) == 0) last[Sv] = node;
                __syncwarp();
            }
        }
    }
}

// --------------------------------------------------------------------------------------------
// k <= 4095 and 16-byte aligned rows: one warp per tree, no block-wide barrier, the string is
// read straight from global memory, 512 bytes per step (one 128-bit load per lane).
//   * every lane classifies its 16 characters into bit masks (")", start of a leaf token, digit);
//     a warp scan of the packed (nodes, leaves) counts gives its first node number and stack
//     height; it then walks its characters once more: stack height after every node, and the
//     value of every digit run that ends inside its piece;
//   * a digit run that touches the end of a lane's piece is finished by the next lane (labels have
//     at most 8 digits, so a run spans at most two pieces);
//   * a node is internal iff the stack height drops at it, so only values are stored (uint16).
// The sweep over the nodes is the one of newick_parse_kernel with uint16 tables: 10 KB per tree.
// --------------------------------------------------------------------------------------------
constexpr uint32_t NPW_NONE = 0xFFFFu;            // ")" without a label
// SWAR helpers on 4 characters in one word (byte 0 = first character)
__device__ __forceinline__ uint32_t swar_zero(uint32_t v) {          // 0x80 in every byte of v that is zero
    return ~(((v & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | v | 0x7F7F7F7Fu);
}
__device__ __forceinline__ uint32_t swar_eq(uint32_t v, char c) { return swar_zero(v ^ (0x01010101u * (uint32_t)(unsigned char)c)); }
__device__ __forceinline__ uint32_t swar_digit(uint32_t v) {         // 0x80 in every byte that is '0'..'9'
    const uint32_t t = v ^ 0x30303030u;
    return ~(((t & 0x7F7F7F7Fu) + 0x76767676u) | t | 0x7F7F7F7Fu);
}
__device__ __forceinline__ uint32_t swar_nibble(uint32_t f) {        // the four 0x80 flags as bits 0..3
    return (((f >> 7) * 0x00204081u) >> 21) & 0xFu;
}
__device__ __forceinline__ uint32_t swar_parse4(uint32_t x) {        // four ASCII digits, first one in byte 0
    x = ((x & 0x0F0F0F0Fu) * 2561u) >> 8;
    return ((x & 0x00FF00FFu) * 6553601u) >> 16;
}
// value of the L (1..8) digits that start at byte a (0..15) of the 16-byte piece ws
__device__ __forceinline__ uint32_t swar_value(const uint32_t (&ws)[4], int a, int L) {
    const int q = a >> 2, r8 = (a & 3) * 8;
    const uint32_t x0 = q == 0 ? ws[0] : (q == 1 ? ws[1] : (q == 2 ? ws[2] : ws[3]));
    const uint32_t x1 = q == 0 ? ws[1] : (q == 1 ? ws[2] : (q == 2 ? ws[3] : 0u));
    const uint32_t x2 = q == 0 ? ws[2] : (q == 1 ? ws[3] : 0u);
    const uint32_t lo = __funnelshift_r(x0, x1, r8), hi = __funnelshift_r(x1, x2, r8);    // 8 bytes from a on
    // left-align: the L digits become the last L of 8 characters, zero bytes (digit 0) in front
    const int sl = 8 * (8 - L);
    const uint32_t hi2 = sl >= 32 ? lo << (sl - 32) : __funnelshift_l(lo, hi, sl);
    const uint32_t lo2 = sl >= 32 ? 0u : lo << sl;
    return swar_parse4(lo2) * 10000u + swar_parse4(hi2);
}

__host__ __device__ inline size_t parse_warp_bytes(int k) {
    return 2 * nw_al16(2 * (2 * (size_t)k + 1)) + nw_al16(2 * ((size_t)k + 3));
}

template <typename OutT>
__global__ void __launch_bounds__(128)
newick_parse_warp_kernel(const char* __restrict__ in, int64_t in_stride, const int64_t* __restrict__ lengths, int64_t B, int k,
                         OutT* __restrict__ edges, int32_t* __restrict__ pstatus, int maxlen) {
    extern __shared__ __align__(16) unsigned char dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int N = 2 * k + 1;
    unsigned char* mine = dyn + (size_t)warp * parse_warp_bytes(k);
    uint16_t* lab = reinterpret_cast<uint16_t*>(mine);                                   // label value by node
    uint16_t* sh = reinterpret_cast<uint16_t*>(mine + nw_al16(2 * (size_t)N));           // stack height after the node
    uint16_t* last = reinterpret_cast<uint16_t*>(mine + 2 * nw_al16(2 * (size_t)N));     // latest node of height h

    for (int64_t tree = (int64_t)blockIdx.x * nwarps + warp; tree < B; tree += (int64_t)gridDim.x * nwarps) {
        const char* src = in + tree * in_stride;
        OutT* erow = edges + (size_t)tree * 4 * (size_t)k;
        const int chunks = (int)((in_stride < (int64_t)maxlen + 16 ? in_stride : (int64_t)maxlen + 16) >> 4);   // readable 16-byte pieces
        int len = 0;
        if (lengths) {
            const int64_t l64 = lengths[tree];
            len = (l64 < 1 || l64 > maxlen || l64 > in_stride) ? 0 : (int)l64;
        } else {                                  // NUL-terminated inside the row
            for (int c0 = 0; c0 < chunks; c0 += 32) {
                const int c = c0 + lane;
                uint4 w = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
                if (c < chunks) w = reinterpret_cast<const uint4*>(src)[c];
                int z = 16;                       // first zero byte of the piece
                const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int q = 3; q >= 0; --q)
#pragma unroll
                    for (int j = 3; j >= 0; --j) if (((ws[q] >> (8 * j)) & 0xFFu) == 0u) z = 4 * q + j;
                const uint32_t m = __ballot_sync(FULL, z < 16);
                if (m) {
                    const int l0 = __ffs(m) - 1;
                    const int p = (c0 + l0) * 16 + __shfl_sync(FULL, z, l0);
                    len = (p < 1 || p > maxlen) ? 0 : p;
                    break;
                }
            }
        }
        int bad = (len == 0) ? 1 : 0;
        // ---- tokens --------------------------------------------------------------------------------
        int carry_idx = 0, carry_S = 0;
        uint32_t carry_prev = 0, carry_tail = 0;                    // tail: value << 4 | digits
        for (int pos = 0; pos < len; pos += 512) {
            const int p0 = pos + lane * 16;
            uint4 w = make_uint4(0, 0, 0, 0);
            if (p0 < len) w = reinterpret_cast<const uint4*>(src)[p0 >> 4];
            const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
            uint32_t prev = __shfl_up_sync(FULL, ws[3] >> 24, 1);
            if (lane == 0) prev = carry_prev;
            // character classes of the 16 bytes as bit masks (SWAR compares, 4 bytes at a time)
            uint32_t mclose = 0, msep = 0, msemi = 0, mdig = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t x = ws[q];
                mclose |= swar_nibble(swar_eq(x, ')')) << (4 * q);
                msep |= swar_nibble(swar_eq(x, '(') | swar_eq(x, ',')) << (4 * q);
                msemi |= swar_nibble(swar_eq(x, ';')) << (4 * q);
                mdig |= swar_nibble(swar_digit(x)) << (4 * q);
            }
            const int left = len - p0;                               // valid characters from p0 on
            const uint32_t vm = left >= 16 ? 0xFFFFu : (left > 0 ? (1u << left) - 1u : 0u);
            if (((mclose | msep | msemi | mdig) & vm) != vm) bad = 1;           // a character outside "(),;" and the digits
            mclose &= vm; msep &= vm; msemi &= vm; mdig &= vm;
            const uint32_t lastbit = (left >= 1 && left <= 16) ? 1u << (left - 1) : 0u;     // the ";" closes the string
            if (msemi != lastbit) bad = 1;
            const uint32_t pdig = ((mdig << 1) | ((prev - '0') <= 9u ? 1u : 0u)) & 0xFFFFu;
            const uint32_t psep = ((msep << 1) | ((prev == '(' || prev == ',') ? 1u : 0u)) & 0xFFFFu;
            const uint32_t pclose = ((mclose << 1) | (prev == ')' ? 1u : 0u)) & 0xFFFFu;
            if (mdig & ~(pdig | psep | pclose)) bad = 1;             // a digit after ";" or at the very start
            const uint32_t mleaf = mdig & psep;
            const uint32_t mnode = mclose | mleaf;
            const int ncl = __popc(mclose), nlf = __popc(mleaf);
            uint32_t sc = (uint32_t)(ncl + nlf) | ((uint32_t)nlf << 16);                // inclusive scan of (nodes, leaves)
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(FULL, sc, d); if (lane >= d) sc += y; }
            const uint32_t tot = __shfl_sync(FULL, sc, 31);
            const uint32_t ex = sc - ((uint32_t)(ncl + nlf) | ((uint32_t)nlf << 16));
            const int idx0 = carry_idx + (int)(ex & 0xFFFFu);
            const int S0 = carry_S + 2 * (int)(ex >> 16) - (int)(ex & 0xFFFFu);
            // the nodes of this piece, one per round: height after the node, and its label if the
            // digits end inside the piece; digits touching the end are handed to the next lane
            uint32_t tail = 0;
            uint32_t mrem = mnode;
            const int rounds = __reduce_max_sync(FULL, ncl + nlf);
            for (int it = 0; it < rounds; ++it) {
                if (mrem) {
                    const int j = __ffs(mrem) - 1;
                    mrem &= mrem - 1;
                    const uint32_t upto = (2u << j) - 1u;
                    const int nidx = idx0 + __popc(mnode & (upto >> 1));
                    const int Sj = S0 + __popc(mleaf & upto) - __popc(mclose & upto);
                    const bool isclose = (mclose >> j) & 1u;
                    if (Sj < 1) bad = 1;
                    const int a0 = j + (isclose ? 1 : 0);            // where the node's digits start
                    const int L = __ffs(~(mdig >> a0)) - 1;          // mdig has 16 bits: L <= 16 - a0
                    uint32_t v = isclose ? NPW_NONE : 0u;
                    if (L > 0) {
                        if (L > 8) bad = 1;
                        const uint32_t val = swar_value(ws, a0, L);
                        if (a0 + L == 16) tail = (val << 4) | (uint32_t)L;
                        else v = val > 0xFFFEu ? 0xFFFEu : val;
                    }
                    if (nidx < N) { sh[nidx] = (uint16_t)Sj; lab[nidx] = (uint16_t)v; }
                    else bad = 1;
                }
            }
            // a piece that starts inside a run (a label continued from, or starting right after, the piece before)
            uint32_t headval = 0, headlen = 0;
            if ((mdig & 1u) && !(mleaf & 1u)) {
                const int L0 = __ffs(~mdig) - 1;
                if (L0 >= 16 || L0 > 8) bad = 1;
                else { headlen = (uint32_t)L0; headval = swar_value(ws, 0, L0); }
            }
            __syncwarp();
            uint32_t ptail = __shfl_up_sync(FULL, tail, 1);
            if (lane == 0) ptail = carry_tail;
            if (ptail | headlen) {
                const uint32_t pv = ptail >> 4, pn = ptail & 15u;
                if (pn + headlen > 8) bad = 1;
                uint32_t total = pv;
                for (uint32_t q = 0; q < headlen; ++q) total *= 10u;
                total += headval;
                if (idx0 >= 1 && idx0 - 1 < N) lab[idx0 - 1] = (uint16_t)(total > 0xFFFEu ? 0xFFFEu : total);
                else bad = 1;
            }
            carry_idx += (int)(tot & 0xFFFFu);
            carry_S += 2 * (int)(tot >> 16) - (int)(tot & 0xFFFFu);
            carry_prev = __shfl_sync(FULL, ws[3] >> 24, 31);
            carry_tail = __shfl_sync(FULL, tail, 31);
        }
        if (carry_idx != N || carry_S != 1) bad = 1;
        __syncwarp();
        if (__any_sync(FULL, bad)) {
            if (lane == 0) { pstatus[tree] = NP_MALFORMED; erow[0] = (OutT)-1; erow[1] = (OutT)-1; }
            continue;
        }
        // ---- children of every internal node, 32 nodes at a time -----------------------------------------
        int verdict = 0;                          // bit 0: malformed label, bit 1: ")" without a label
        int carryS = 0;
        for (int base = 0; base < N; base += 32) {
            const int node = base + lane;
            const bool act = node < N;
            const int Sv = act ? (int)sh[node] : -1 - lane;
            int Sp = __shfl_up_sync(FULL, Sv, 1);
            if (lane == 0) Sp = carryS;
            carryS = __shfl_sync(FULL, Sv, 31);
            const bool internal = act && Sv < Sp;
            const uint32_t lb = act ? (uint32_t)lab[node] : 0u;
            const uint32_t peers = __match_any_sync(FULL, Sv);
            if (act && !internal && lb > (uint32_t)k) verdict |= 1;
            if (internal) {
                if (lb == NPW_NONE) verdict |= 2; else if (lb > 2u * (uint32_t)k) verdict |= 1;
                const uint32_t lower = peers & lanes_lt(lane);
                const int j = lower ? base + 31 - __clz(lower) : (int)last[Sv];
                const uint32_t c1 = lab[j], c2 = lab[node - 1];
                const int r = ((node + 1 - Sv) >> 1) - 1;            // internal nodes before this one
                OutT* o = erow + 4 * (size_t)r;
                o[0] = (OutT)c1; o[1] = (OutT)lb; o[2] = (OutT)c2; o[3] = (OutT)lb;
            }
            __syncwarp();
            if (act && (peers & lanes_gt(lane)) == 0) last[Sv] = (uint16_t)node;
            __syncwarp();
        }
        const uint32_t anym = __ballot_sync(FULL, verdict & 1), anyn = __ballot_sync(FULL, verdict & 2);
        if (lane == 0) {
            const int v = anym ? NP_MALFORMED : (anyn ? NP_NOPARENTS : 0);
            pstatus[tree] = v;
            if (v) { erow[0] = (OutT)-1; erow[1] = (OutT)-1; }
        }
        __syncwarp();
    }
}

// Trees the first kernel marked NP_NOPARENTS.
template <typename OutT, bool STAGED>
__global__ void __launch_bounds__(NW_THREADS_MAX)
newick_parse_np_kernel(const char* __restrict__ in, int64_t in_stride, const int64_t* __restrict__ lengths, int64_t B, int k,
                       OutT* __restrict__ edges, int32_t* __restrict__ pstatus, unsigned char* __restrict__ tables,
                       size_t tables_stride, int maxlen) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ ParseShared P;
    __shared__ int s_scan[32];
    unsigned char* dyn = STAGED ? dyn_smem : tables + (size_t)blockIdx.x * tables_stride;
    const ParseLayout L = parse_layout(k, STAGED, true);
    const int N = 2 * k + 1, n = k + 1;
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = NT >> 5;
    uint32_t* lab = reinterpret_cast<uint32_t*>(dyn + L.lab);
    int* sh = reinterpret_cast<int*>(dyn + L.sh);
    int* last = reinterpret_cast<int*>(dyn + L.last);
    int* c1a = reinterpret_cast<int*>(dyn + L.c1a);                // first child (node index) of an internal node
    int* first = reinterpret_cast<int*>(dyn + L.first);            // first node of the node's subtree
    uint32_t* leaf = reinterpret_cast<uint32_t*>(dyn + L.leaf);    // leaf labels in string order
    uint32_t* sister = reinterpret_cast<uint32_t*>(dyn + L.sister);  // sort key of cherry r
    int* gt = reinterpret_cast<int*>(dyn + L.gt);                  // histogram, then cherries with a larger key
    uint32_t* sparse = reinterpret_cast<uint32_t*>(dyn + L.sparse);  // argmin of leaf[] over [a, a + 2^l)
    const int levels = L.levels;

    for (int64_t tree = blockIdx.x; tree < B; tree += gridDim.x) {
        if (pstatus[tree] != NP_NOPARENTS) continue;               // uniform
        OutT* erow = edges + (size_t)tree * 4 * (size_t)k;
        const int verdict = newick_tokenize<false, STAGED>(in + tree * in_stride, in_stride, lengths, tree, k, maxlen,
                                                           reinterpret_cast<char*>(dyn + L.str), lab, sh, P);
        if (verdict) {                            // uniform; row 0 already holds the reject marker
            if (tid == 0) pstatus[tree] = NP_MALFORMED;
            continue;
        }
        // ---- first child and first node of every subtree: one warp, 32 nodes at a time ----------
        if (warp == 0) {
            for (int base = 0; base < N; base += 32) {
                const int node = base + lane;
                const bool act = node < N;
                const bool internal = act && (lab[act ? node : 0] & NP_INTERNAL);
                const int Sv = act ? sh[node] : -1 - lane;
                const uint32_t peers = __match_any_sync(FULL, Sv);
                int val = node, ptr = lane;       // leaves (and padding) are resolved: their subtree starts at themselves
                if (internal) {
                    const uint32_t lower = peers & lanes_lt(lane);
                    const int j = lower ? base + 31 - __clz(lower) : last[Sv];
                    c1a[node] = j;
                    if (j >= base) ptr = j - base;                 // first(node) = first(first child), still open
                    else val = first[j];
                }
#pragma unroll
                for (int it = 0; it < 5; ++it) {                   // chains inside the batch
                    const int nv = __shfl_sync(FULL, val, ptr), np = __shfl_sync(FULL, ptr, ptr);
                    if (ptr != lane) { val = nv; ptr = np; }
                }
                if (act) first[node] = val;
                __syncwarp();
                if (act && (peers & lanes_gt(lane)) == 0) last[Sv] = node;
                __syncwarp();
            }
        }
        // ---- leaf labels in string order, sparse table of argmins ---------------------------------
        for (int i = tid; i < N; i += NT) {
            const uint32_t lb = lab[i];
            if (!(lb & NP_INTERNAL)) {
                const int ord = ((i + 1 + sh[i]) >> 1) - 1;          // leaves before this one
                leaf[ord] = lb; sparse[ord] = (uint32_t)ord;
            }
        }
        for (int i = tid; i <= n; i += NT) gt[i] = 0;
        __syncthreads();
        for (int l = 1; l < levels; ++l) {
            const uint32_t* lo = sparse + (size_t)(l - 1) * n;
            uint32_t* hi = sparse + (size_t)l * n;
            const int half = 1 << (l - 1);
            for (int a = tid; a + 2 * half <= n; a += NT) {
                const uint32_t x = lo[a], y = lo[a + half];
                hi[a] = leaf[y] < leaf[x] ? y : x;
            }
            __syncthreads();
        }
        auto argmin = [&](int a, int b) -> int {                   // [a, b], a <= b
            const int l = 31 - __clz(b - a + 1);
            const uint32_t x = sparse[(size_t)l * n + a], y = sparse[(size_t)l * n + b - (1 << l) + 1];
            return (int)(leaf[y] < leaf[x] ? y : x);
        };
        // ---- sort key of every cherry: the second-smallest leaf of its subtree ------------------------
        for (int i = tid; i < N; i += NT) {
            if (!(lab[i] & NP_INTERNAL)) continue;
            const int f = first[i];
            const int la = f == 0 ? 0 : (f + sh[f - 1]) >> 1;       // leaves before the subtree
            const int lb = ((i + 1 + sh[i]) >> 1) - 1;              // its last leaf
            const int a = argmin(la, lb);
            uint32_t m2 = 0xFFFFFFFFu;
            if (a > la) m2 = leaf[argmin(la, a - 1)];
            if (a < lb) m2 = min(m2, leaf[argmin(a + 1, lb)]);
            const int r = ((i + 1 - sh[i]) >> 1) - 1;
            sister[r] = m2;
            atomicAdd(&gt[m2 <= (uint32_t)k ? m2 : 0], 1);
        }
        __syncthreads();
        // ---- gt[s] = cherries whose key is larger than s (suffix sums over the bins 0..k) ------------
        {
            const int CH = (n + NT - 1) / NT;
            const int hi_end = n - tid * CH, lo_end = max(0, hi_end - CH);      // this thread's bins [lo_end, hi_end), from the top
            int sum = 0;
            for (int b = hi_end - 1; b >= lo_end; --b) sum += gt[b];
            int inc = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += y; }
            if (lane == 31) s_scan[warp] = inc;
            __syncthreads();
            if (warp == 0) {
                int a = lane < nwarp ? s_scan[lane] : 0;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(FULL, a, d); if (lane >= d) a += y; }
                s_scan[lane] = a;
            }
            __syncthreads();
            int run = inc - sum + (warp ? s_scan[warp - 1] : 0);                 // cherries in the bins above this thread's
            for (int b = hi_end - 1; b >= lo_end; --b) { const int c = gt[b]; gt[b] = run; run += c; }
        }
        __syncthreads();
        // ---- position of every cherry: larger keys first, equal keys in row order; label k + 1 + q ---
        if (warp == 0) {
            for (int base = 0; base < N; base += 32) {
                const int node = base + lane;
                const bool internal = node < N && (lab[node < N ? node : 0] & NP_INTERNAL);
                const int r = internal ? ((node + 1 - sh[node]) >> 1) - 1 : 0;
                const uint32_t key = internal ? sister[r] : 0x80000000u + (uint32_t)lane;
                const uint32_t peers = __match_any_sync(FULL, key);
                if (internal) {
                    const uint32_t bin = key <= (uint32_t)k ? key : 0u;
                    const int q = gt[bin] + __popc(peers & lanes_lt(lane));
                    lab[node] = NP_INTERNAL | (uint32_t)(k + 1 + q);
                }
                __syncwarp();
                if (internal && (peers & lanes_gt(lane)) == 0) gt[key <= (uint32_t)k ? key : 0u] += __popc(peers);
                __syncwarp();
            }
        }
        __syncthreads();
        for (int i = tid; i < N; i += NT) {
            const uint32_t lb = lab[i];
            if (!(lb & NP_INTERNAL)) continue;
            const uint32_t c1 = lab[c1a[i]] & ~NP_INTERNAL, c2 = lab[i - 1] & ~NP_INTERNAL, pl = lb & ~NP_INTERNAL;
            const int r = ((i + 1 - sh[i]) >> 1) - 1;
            OutT* o = erow + 4 * (size_t)r;
            o[0] = (OutT)c1; o[1] = (OutT)pl; o[2] = (OutT)c2; o[3] = (OutT)pl;
        }
        if (tid == 0) pstatus[tree] = 0;
    }
}

__global__ void merge_status_kernel(const int32_t* __restrict__ pst, int32_t* __restrict__ status, int64_t B) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < B; t += (int64_t)gridDim.x * blockDim.x)
        if (pst[t] != 0) status[t] = pst[t];
}

// workspace: [ parse status (B int32) | edges (B,2k,2) of the index type | encode workspace | per-CTA tables ]
struct FromNewickWs { size_t pst, edges, enc, enc_bytes, tables, tables_stride, total; int64_t teams, teams_np; };
static FromNewickWs from_newick_ws(int64_t B, int64_t k, int idx_bytes) {
    FromNewickWs w;
    size_t o = 0;
    w.pst = o; o = nw_al16(o + 4 * (size_t)B);
    w.edges = o; o = nw_al16(o + (size_t)idx_bytes * 4 * (size_t)B * (size_t)k);
    w.enc_bytes = encode_workspace_bytes(B, k);
    w.enc = o; o = nw_al16(o + w.enc_bytes);
    // the two parse kernels run one after the other and share the table area
    const size_t s1 = parse_fits_smem(k, false) ? 0 : nw_al16(parse_layout(k, false, false).total);
    const size_t s2 = parse_fits_smem(k, true) ? 0 : nw_al16(parse_layout(k, false, true).total);
    const int64_t cap = (int64_t)sm_count() * 2;
    w.teams = B < cap ? B : cap;
    w.tables_stride = s2 > s1 ? s2 : s1;
    w.teams_np = w.teams;
    if (s2) {                                     // the sparse table is 4 (k+1) log2(k+1) bytes per CTA: at most 1 GB in all
        const int64_t fit = (int64_t)(((size_t)1 << 30) / s2);
        w.teams_np = fit < 1 ? 1 : (fit < w.teams ? fit : w.teams);
    }
    const size_t area1 = s1 * (size_t)w.teams, area2 = s2 * (size_t)w.teams_np;
    w.tables = o; o = nw_al16(o + (area1 > area2 ? area1 : area2));
    w.total = o;
    return w;
}
size_t from_newick_workspace_bytes(int64_t B, int64_t k, int idx_bytes) {
    return (B <= 0 || k <= 0) ? 0 : from_newick_ws(B, k, idx_bytes).total;
}

template <typename OutT>
static cudaError_t launch_parse_t(const char* in, int64_t in_stride, const int64_t* lengths, int64_t B, int64_t k,
                                  OutT* edges, int32_t* pst, unsigned char* tables, const FromNewickWs& W, cudaStream_t stream) {
    const int maxlen = (int)newick_bytes(k);
    for (int pass = 0; pass < 2; ++pass) {        // 0: strings with parent labels, 1: the trees pass 0 left as NP_NOPARENTS
        const bool np = pass == 1;
        if (!np && k <= NWW_K_MAX && (reinterpret_cast<uintptr_t>(in) & 15) == 0 && (in_stride & 15) == 0) {
            const size_t per = parse_warp_bytes((int)k);
            int wpc = 1, best = 0;
            for (int w = 4; w >= 1; w >>= 1) {
                size_t ctas = ((size_t)228 * 1024) / (per * w + 1024);
                if (ctas > 32) ctas = 32;
                if (ctas > (size_t)(64 / w)) ctas = 64 / w;
                if ((int)ctas * w > best) { best = (int)ctas * w; wpc = w; }
            }
            const size_t smem = per * wpc;
            auto kern = newick_parse_warp_kernel<OutT>;
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            int per_sm = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpc * 32, smem);
            if (e != cudaSuccess) return e;
            if (per_sm < 1) per_sm = 1;
            const int64_t cap = (int64_t)sm_count() * per_sm, want = (B + wpc - 1) / wpc;
            kern<<<(int)(want < cap ? want : cap), wpc * 32, smem, stream>>>(in, in_stride, lengths, B, (int)k, edges, pst, maxlen);
        } else if (parse_fits_smem(k, np)) {
            const size_t smem = parse_layout(k, true, np).total;
            auto kern = np ? newick_parse_np_kernel<OutT, true> : newick_parse_kernel<OutT, true>;
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            // the kernel for unlabelled strings has more all-thread phases (sparse table, keys): wider CTAs
            const int threads = np ? (k <= 127 ? 64 : (k <= 511 ? 128 : 512)) : (k <= 255 ? 64 : (k <= 2047 ? 128 : 256));
            int per_sm = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
            if (e != cudaSuccess) return e;
            if (per_sm < 1) per_sm = 1;
            const int64_t cap = (int64_t)sm_count() * per_sm;
            kern<<<(int)(B < cap ? B : cap), threads, smem, stream>>>(in, in_stride, lengths, B, (int)k, edges, pst, nullptr, 0, maxlen);
        } else {
            const size_t stride = nw_al16(parse_layout(k, false, np).total);
            auto kern = np ? newick_parse_np_kernel<OutT, false> : newick_parse_kernel<OutT, false>;
            kern<<<(int)(np ? W.teams_np : W.teams), NW_THREADS_MAX, 0, stream>>>(in, in_stride, lengths, B, (int)k, edges, pst,
                                                                                 tables, stride, maxlen);
        }
        count_launch();
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_from_newick(const char* in, int64_t in_stride, const int64_t* lengths, int64_t B, int64_t k, void* v,
                               int idx_bytes, int32_t* status, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    if (k < 1 || k > LARGE_K_MAX) return cudaErrorNotSupported;
    if (idx_bytes != 4 && idx_bytes != 8) return cudaErrorInvalidValue;
    const FromNewickWs W = from_newick_ws(B, k, idx_bytes);
    if (!ws || ws_bytes < W.total) return cudaErrorInvalidValue;
    unsigned char* wsb = static_cast<unsigned char*>(ws);
    int32_t* pst = reinterpret_cast<int32_t*>(wsb + W.pst);
    cudaError_t e;
    if (idx_bytes == 8)
        e = launch_parse_t<long long>(in, in_stride, lengths, B, k, reinterpret_cast<long long*>(wsb + W.edges), pst,
                                      wsb + W.tables, W, stream);
    else
        e = launch_parse_t<int>(in, in_stride, lengths, B, k, reinterpret_cast<int*>(wsb + W.edges), pst, wsb + W.tables, W,
                                stream);
    if (e != cudaSuccess) return e;
    e = launch_encode(FROM_EDGES, wsb + W.edges, idx_bytes, B, k, v, status, W.enc_bytes ? wsb + W.enc : nullptr,
                      W.enc_bytes, stream);
    if (e != cudaSuccess) return e;
    if (status) {
        merge_status_kernel<<<(int)((B + 255) / 256 < 2048 ? (B + 255) / 256 : 2048), 256, 0, stream>>>(pst, status, B);
        count_launch();
        e = cudaGetLastError();
    }
    return e;
}

}  // namespace p2v
