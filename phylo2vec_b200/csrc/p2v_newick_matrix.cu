// Newick strings with branch lengths <-> Phylo2Vec matrices, batched, for sm_100a.
//
// Replaces (reference, paths relative to /root/reference):
//   matrix::convert::to_newick / build_newick_with_bls     phylo2vec/src/matrix/convert.rs:26-57
//   matrix::convert::from_newick                           phylo2vec/src/matrix/convert.rs:84-122
//   newick::parse_with_bls, node_substr                    phylo2vec/src/newick/mod.rs:23-38, 161-255
//   vector::convert::order_cherries / build_cherries / order_cherries_no_parents (their branch-length outputs)
//                                                          phylo2vec/src/vector/convert.rs:398-477, 534-573
//   PyO3 to_newick (matrix arm), to_matrix                 py-phylo2vec/src/lib.rs:90-106
//
// to_newick(matrix).  S(leaf) = label, S(u) = "(" S(c1) ":" bl1 "," S(c2) ":" bl2 ")" u for an internal node u
// whose ancestry row holds (c1, c2) and the matrix row (bl1, bl2); the tree's string is S(root) ";".  Lengths
// are the shortest round-trip decimals in the Rust ryu layout (p2v_float_text.cuh).  One warp per tree: the
// batch is decoded first (p2v_decode.cu, int32 children columns), then len[u] bottom-up, off[u] top-down
// (p2v_tree_pass.cuh) and every row writes its own characters.
//
// from_newick(matrix).  One warp per tree.  1. classify every character 32 at a time: a token starts at ")" or at
// a digit that follows "(" or ","; the prefix counts of leaf / internal tokens give every token its post-order
// index and the stack height S after it (parse_with_bls is a stack walk: a leaf pushes, ")" pops two and
// pushes one).  2. every token is parsed by one lane: label, ":" and the length (correctly rounded,
// p2v_float_text.cuh).  3. the first child of internal node i is the latest j < i with S(j) = S(i), the second
// is node i - 1: a sweep over the nodes with a "latest node of height h" table.  4. the row of a cherry in the
// matrix: its parent label - k - 1 with parent labels (order_cherries); without them the stable descending
// order of the second-smallest leaf label below the cherry (order_cherries_no_parents: (min1, min2) of every
// subtree bottom-up, then a counting sort), and the cherry is given the label k + 1 + row.  5. min_desc of
// every row bottom-up: the two lengths of a row swap when the first child's smallest leaf is the larger one
// (build_cherries).  The rows (children labels, parent) then go through the batched from_ancestry encoder
// (build_cherries + build_vector, p2v_encode.cu), and column 0 of the matrix is that vector.
#include <cstdlib>

#include "p2v_common.cuh"
#include "p2v_float_text.cuh"
#include "p2v_internal.h"
#include "p2v_tree_pass.cuh"

namespace p2v {

constexpr int NM_WARPS = 4;
constexpr int NM_FLOAT_MAX = 24;               // longest shortest-round-trip string of a double

__device__ __forceinline__ int nm_digits(uint32_t v) {
    int n = 1;
    while (v >= 10u) { v /= 10u; ++n; }
    return n;
}
__device__ __forceinline__ void nm_put_uint(char* p, uint32_t v, int nd) {
    for (int i = nd - 1; i >= 0; --i) { p[i] = (char)('0' + (int)(v % 10u)); v /= 10u; }
}

// sum_{x=0..n-1} digits(x) * ... : characters of the labels of all 2k+1 nodes plus the structure of a binary
// tree; here only an upper bound on the string is needed
size_t newick_matrix_max_bytes(int64_t k) {
    if (k < 0) return 0;
    return newick_max_bytes(k) + (size_t)(2 * k) * (1 + NM_FLOAT_MAX);
}

// ---------------------------------------------------------------------------------------- to_newick
// scratch words per warp: ch[2k] par[k] len[k] off[k] fl[2k]
__host__ __device__ inline size_t nm_to_words(int64_t k) { return (size_t)(7 * k + 32); }

__global__ void __launch_bounds__(NM_WARPS * 32)
newick_matrix_write_kernel(const int* __restrict__ anc2, const double* __restrict__ m, int64_t B, int k,
                           char* __restrict__ out, int64_t out_stride, int64_t* __restrict__ lengths,
                           const int32_t* __restrict__ st, uint32_t* __restrict__ scratch, size_t scratch_words) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    uint32_t* scr = scratch + ((size_t)blockIdx.x * nwarps + warp) * scratch_words;
    uint32_t* ch = scr;
    uint32_t* par = ch + 2 * k;
    uint32_t* len = par + k;
    uint32_t* off = len + k;
    uint32_t* fl = off + k;
    for (int64_t t = (int64_t)blockIdx.x * nwarps + warp; t < B; t += (int64_t)gridDim.x * nwarps) {
        char* o = out + t * out_stride;
        if (st[t] != 0) {
            if (lane == 0) { o[0] = 0; if (lengths) lengths[t] = 0; }
            continue;
        }
        const int2* anc = reinterpret_cast<const int2*>(anc2 + (size_t)t * 2 * k);
        const double* mt = m + (size_t)t * 3 * k;
        for (int r = lane; r < k; r += 32) {
            const int2 x = anc[r];
            ch[2 * r] = (uint32_t)x.x; ch[2 * r + 1] = (uint32_t)x.y;
            if (x.x > k) par[x.x - k - 1] = (uint32_t)r;
            if (x.y > k) par[x.y - k - 1] = (uint32_t)r;
            char tmp[NM_FLOAT_MAX + 1];
            fl[2 * r] = (uint32_t)p2v_ft::f64_to_shortest(mt[3 * r + 1], tmp);
            fl[2 * r + 1] = (uint32_t)p2v_ft::f64_to_shortest(mt[3 * r + 2], tmp);
        }
        __syncwarp();
        // len[r] = characters of S(node of row r)
        rf_bottom_up<uint32_t>(ch, k, lane, [&](int r, int c1, int c2) {
            const uint32_t l1 = c1 > k ? len[c1 - k - 1] : (uint32_t)nm_digits((uint32_t)c1);
            const uint32_t l2 = c2 > k ? len[c2 - k - 1] : (uint32_t)nm_digits((uint32_t)c2);
            len[r] = 3u + (uint32_t)nm_digits((uint32_t)(k + 1 + r)) + l1 + 1u + fl[2 * r] + l2 + 1u + fl[2 * r + 1];
        });
        __syncwarp();
        // off[r] = where S(node of row r) starts
        rf_top_down<uint32_t>(par, k, lane, [&](int r, int p) {
            uint32_t start = 0;
            if (p >= 0) {
                const int pc1 = (int)ch[2 * p];
                start = off[p] + 1u;
                if (pc1 != k + 1 + r) {
                    const uint32_t l1 = pc1 > k ? len[pc1 - k - 1] : (uint32_t)nm_digits((uint32_t)pc1);
                    start += l1 + 1u + fl[2 * p] + 1u;
                }
            }
            off[r] = start;
        });
        __syncwarp();
        for (int r = lane; r < k; r += 32) {
            const int c1 = (int)ch[2 * r], c2 = (int)ch[2 * r + 1];
            char* p = o + off[r];
            *p++ = '(';
            if (c1 <= k) { const int nd = nm_digits((uint32_t)c1); nm_put_uint(p, (uint32_t)c1, nd); p += nd; }
            else p += len[c1 - k - 1];
            *p++ = ':';
            p += p2v_ft::f64_to_shortest(mt[3 * r + 1], p);
            *p++ = ',';
            if (c2 <= k) { const int nd = nm_digits((uint32_t)c2); nm_put_uint(p, (uint32_t)c2, nd); p += nd; }
            else p += len[c2 - k - 1];
            *p++ = ':';
            p += p2v_ft::f64_to_shortest(mt[3 * r + 2], p);
            *p++ = ')';
            const int nd = nm_digits((uint32_t)(k + 1 + r));
            nm_put_uint(p, (uint32_t)(k + 1 + r), nd);
            p += nd;
            if (r == k - 1) { *p++ = ';'; *p = 0; if (lengths) lengths[t] = (int64_t)(p - o); }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------- from_newick
constexpr uint32_t NM_INT = 0x80000000u;        // token of an internal node
constexpr uint32_t NM_NOLAB = 0xFFFFFFFFu;
constexpr uint32_t NM_INF = 0x7FFFFFFFu;

// scratch words per warp: pos[N] ks[N] lab[N] | bl[N] doubles | hl[k+2] | cn1 cn2 q mn1 mn2 mark [k] | cnt[k+2]
__host__ __device__ inline size_t nm_from_words(int64_t k) {
    const size_t N = 2 * (size_t)k + 1;
    return 3 * N + 1 + 2 * N + ((size_t)k + 2) + 6 * (size_t)k + ((size_t)k + 2) + 32;
}

__device__ __forceinline__ bool nm_delim(char c) { return c == '(' || c == ')' || c == ',' || c == ';'; }

// anc out: (k,3) int32 rows [c1, c2, k+1+row]; bl out: (k,2) float64 in row order, swapped where build_cherries swaps
__global__ void __launch_bounds__(NM_WARPS * 32)
newick_matrix_parse_kernel(const char* __restrict__ in, int64_t in_stride, const int64_t* __restrict__ lengths, int64_t B,
                           int k, int* __restrict__ anc_out, double* __restrict__ bl_out, int32_t* __restrict__ status,
                           int maxlen, uint32_t* __restrict__ scratch, size_t scratch_words) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int N = 2 * k + 1;
    // big-number scratch of the exact decision for long literals (p2v_float_text.cuh: parse_f64_exact), one per warp
    __shared__ uint32_t s_big[NM_WARPS][2 * p2v_ft::FT_BIG_LIMBS];
    uint32_t* scr = scratch + ((size_t)blockIdx.x * nwarps + warp) * scratch_words;
    uint32_t* pos = scr;
    uint32_t* ks = pos + N;
    uint32_t* lab = ks + N;
    double* bl = reinterpret_cast<double*>(lab + N + ((3 * N) & 1));      // 8-byte aligned (scratch_words is even)
    uint32_t* hl = reinterpret_cast<uint32_t*>(bl + N);
    uint32_t* cn1 = hl + k + 2;
    uint32_t* cn2 = cn1 + k;
    uint32_t* q = cn2 + k;
    uint32_t* mn1 = q + k;
    uint32_t* mn2 = mn1 + k;
    uint32_t* mark = mn2 + k;
    uint32_t* cnt = mark + k;
    for (int64_t t = (int64_t)blockIdx.x * nwarps + warp; t < B; t += (int64_t)gridDim.x * nwarps) {
        const char* s = in + t * in_stride;
        int* anc = anc_out + (size_t)t * 3 * k;
        double* blo = bl_out + (size_t)t * 2 * k;
        int len = 0;
        if (lengths) {
            const int64_t l64 = lengths[t];
            len = (l64 < 1 || l64 > maxlen || l64 > in_stride) ? 0 : (int)l64;
        } else {
            const int lim = (int)(in_stride < (int64_t)maxlen + 1 ? in_stride : (int64_t)maxlen + 1);
            int found = -1;
            for (int base = 0; base < lim && found < 0; base += 32) {
                const int i = base + lane;
                const unsigned z = __ballot_sync(FULL, i < lim && s[i] == 0);
                if (z) found = base + __ffs(z) - 1;
            }
            len = found < 1 ? 0 : found;
        }
        int bad = (len == 0) ? 1 : 0;
        // ---- 1. tokens: post-order index and stack height from prefix counts -------------------------
        int nnodes = 0, nleaves = 0;
        for (int base = 0; base < len && !bad; base += 32) {
            const int i = base + lane;
            const char c = i < len ? s[i] : '(';
            const char p = (i > 0 && i <= len) ? s[i - 1] : '(';
            const bool in = i < len;
            const bool dig = c >= '0' && c <= '9';
            const bool after_open = p == '(' || p == ',';
            const bool leafstart = in && dig && after_open;
            const bool intstart = in && c == ')';
            bool wrong = false;
            if (in) {
                if (c == ';') wrong = i != len - 1;
                else if (c == '(') wrong = !(after_open);                 // "(" only after "(" or "," (or at the start)
                else if (c == ',' || c == ')') wrong = after_open;         // an empty node
                else wrong = after_open && !dig;                           // a token body starts with a digit
                if (i == len - 1 && c != ';') wrong = true;
            }
            const unsigned ml = __ballot_sync(FULL, leafstart), mi = __ballot_sync(FULL, intstart);
            const unsigned below = lanes_lt(lane);
            const int idx = nnodes + __popc((ml | mi) & below);
            const int leaves_incl = nleaves + __popc(ml & lanes_le(lane));
            if ((leafstart || intstart) && idx < N) {
                const int ints_incl = idx + 1 - leaves_incl;
                const int S = leaves_incl - ints_incl;
                pos[idx] = (uint32_t)i;
                ks[idx] = (uint32_t)(S < 0 ? 0 : S) | (intstart ? NM_INT : 0u);
                if (S < 1) wrong = true;
            }
            nnodes += __popc(ml | mi);
            nleaves += __popc(ml);
            if (__any_sync(FULL, wrong) || nnodes > N) bad = 1;       // warp-uniform: every lane leaves the loop together
        }
        if (nnodes != N || nleaves != k + 1) bad = 1;
        __syncwarp();
        // ---- 2. every token: label, ":" and length ---------------------------------------------------
        int any_lab = 0, any_nolab = 0, und = 0;
        if (!bad) {
            for (int idx = lane; idx < N; idx += 32) {
                const int start = (int)pos[idx];
                const bool isint = (ks[idx] & NM_INT) != 0;
                int i = start + (isint ? 1 : 0);
                uint32_t v = 0;
                int nd = 0;
                while (i < len && s[i] >= '0' && s[i] <= '9' && nd < 10) { v = v * 10u + (uint32_t)(s[i] - '0'); ++nd; ++i; }
                if (nd > 9) bad = 1;
                double x = 0.0;
                bool has_bl = false;
                if (i < len && s[i] == ':') {
                    const int f0 = i + 1;
                    int f1 = f0;
                    while (f1 < len && !nm_delim(s[f1])) ++f1;
                    if (f1 - f0 > 1200) bad = 1;
                    else {
                        const int rc = p2v_ft::parse_f64(s + f0, f1 - f0, &x);
                        if (rc == p2v_ft::PARSE_F64_UNDECIDED) und = 1;                 // decided below, one lane at a time
                        else if (rc != p2v_ft::PARSE_F64_OK) bad = 1;
                    }
                    has_bl = true;
                    i = f1;
                }
                if (i >= len || !nm_delim(s[i]) || s[i] == '(') bad = 1;          // the token ends at "," ")" or ";"
                if (!isint) {
                    if (nd == 0 || v > (uint32_t)k || !has_bl) bad = 1;           // "label:length" (split_once(':').unwrap())
                    lab[idx] = v;
                } else {
                    if (idx == N - 1) { if (i != len - 1) bad = 1; }               // the root closes the string
                    else if (!has_bl) bad = 1;                                     // "Missing annotation in the Newick string"
                    if (nd == 0) { lab[idx] = NM_NOLAB; any_nolab = 1; }
                    else { lab[idx] = v; any_lab = 1; if (v <= (uint32_t)k || v > 2u * (uint32_t)k) bad = 1; }
                }
                bl[idx] = x;
            }
            if (lane == 0 && !(ks[N - 1] & NM_INT)) bad = 1;
            if (lane == 0 && (ks[N - 1] & ~NM_INT) != 1u) bad = 1;
        }
        bad = __any_sync(FULL, bad) ? 1 : 0;
        if (!bad && __any_sync(FULL, und)) {
            // lengths of more than 19 digits that the fast parser left undecided (about one in a thousand of those):
            // the lanes that saw one find it again and decide it exactly, one at a time on the warp's big-number scratch
            for (int base = 0; base < N; base += 32) {
                const int idx = base + lane;
                int f0 = 0, f1 = 0;
                bool need = false;
                if (und && idx < N) {
                    int i = (int)pos[idx] + ((ks[idx] & NM_INT) ? 1 : 0);
                    while (i < len && s[i] >= '0' && s[i] <= '9') ++i;
                    if (i < len && s[i] == ':') {
                        f0 = i + 1; f1 = f0;
                        while (f1 < len && !nm_delim(s[f1])) ++f1;
                        double tmp;
                        need = p2v_ft::parse_f64(s + f0, f1 - f0, &tmp) == p2v_ft::PARSE_F64_UNDECIDED;
                    }
                }
                unsigned m = __ballot_sync(FULL, need);
                while (m) {
                    const int turn = __ffs(m) - 1;
                    m &= m - 1;
                    if (lane == turn) {
                        double x = 0.0;
                        if (p2v_ft::parse_f64_exact(s + f0, f1 - f0, s_big[warp], s_big[warp] + p2v_ft::FT_BIG_LIMBS, &x) != p2v_ft::PARSE_F64_OK) bad = 1;
                        bl[idx] = x;
                    }
                    __syncwarp();
                }
            }
            bad = __any_sync(FULL, bad) ? 1 : 0;
        }
        any_lab = __any_sync(FULL, any_lab); any_nolab = __any_sync(FULL, any_nolab);
        if (any_lab && any_nolab) bad = 1;                                       // parent labels on some ")" only
        const bool labelled = any_lab != 0;
        if (bad) { if (lane == 0) status[t] = TREE_MALFORMED; continue; }
        __syncwarp();
        // ---- 3. children of every internal node: the sweep over the stack heights ---------------------
        for (int h = lane; h < k + 2; h += 32) hl[h] = NM_NOLAB;
        __syncwarp();
        for (int base = 0; base < N; base += 32) {
            const int i = base + lane;
            const bool act = i < N;
            const uint32_t w = act ? ks[i] : 0u;
            const uint32_t S = w & ~NM_INT;
            const bool isint = act && (w & NM_INT);
            const unsigned peers = __match_any_sync(FULL, act ? S : (0x40000000u + (uint32_t)lane));
            if (isint) {
                const unsigned lower = peers & lanes_lt(lane);
                const uint32_t c1n = lower ? (uint32_t)(base + 31 - __clz(lower)) : hl[S];
                const int tcherry = (i + 1 - (int)S) / 2 - 1;          // internals up to and including i, minus one
                if (c1n == NM_NOLAB || tcherry < 0 || tcherry >= k) bad = 1;
                else { cn1[tcherry] = c1n; cn2[tcherry] = (uint32_t)(i - 1); }
            }
            __syncwarp();
            if (act && (peers & lanes_gt(lane)) == 0) hl[S] = (uint32_t)i;
            __syncwarp();
        }
        bad = __any_sync(FULL, bad) ? 1 : 0;
        if (bad) { if (lane == 0) status[t] = TREE_MALFORMED; continue; }
        // cherry index of an internal node
        auto cherry_of = [&](uint32_t node) -> int { return ((int)node + 1 - (int)(ks[node] & ~NM_INT)) / 2 - 1; };
        // ---- 4. the row of every cherry ----------------------------------------------------------------
        for (int r = lane; r < k; r += 32) mark[r] = 0u;
        __syncwarp();
        if (labelled) {
            for (int c = lane; c < k; c += 32) {
                // the internal node of cherry c is its second child's node + 1
                const uint32_t node = cn2[c] + 1u;
                q[c] = lab[node] - (uint32_t)k - 1u;                 // k < label <= 2k was checked
            }
        } else {
            // (min1, min2) of the leaf labels below every cherry, bottom-up in appearance order
            // (children cherries come first); rows as an ancestry-like table for the generic pass
            uint32_t* chx = pos;                                     // pos[] is no longer needed: 2k words
            for (int c = lane; c < k; c += 32) {
                const uint32_t a = cn1[c], b = cn2[c];
                chx[2 * c] = (ks[a] & NM_INT) ? (uint32_t)(k + 1 + cherry_of(a)) : lab[a];
                chx[2 * c + 1] = (ks[b] & NM_INT) ? (uint32_t)(k + 1 + cherry_of(b)) : lab[b];
            }
            __syncwarp();
            rf_bottom_up<uint32_t>(chx, k, lane, [&](int r, int c1, int c2) {
                uint32_t a1, a2, b1, b2;
                if (c1 > k) { a1 = mn1[c1 - k - 1]; a2 = mn2[c1 - k - 1]; } else { a1 = (uint32_t)c1; a2 = NM_INF; }
                if (c2 > k) { b1 = mn1[c2 - k - 1]; b2 = mn2[c2 - k - 1]; } else { b1 = (uint32_t)c2; b2 = NM_INF; }
                const uint32_t lo = min(a1, b1), hi = max(a1, b1);
                mn1[r] = lo;
                mn2[r] = min(hi, min(a2, b2));
            });
            __syncwarp();
            // stable descending counting sort by key = mn2 (a leaf label in 1..k)
            for (int h = lane; h < k + 2; h += 32) cnt[h] = 0u;
            __syncwarp();
            for (int c = lane; c < k; c += 32) {
                if (mn2[c] == 0u || mn2[c] > (uint32_t)k) bad = 1;
                else atomicAdd(&cnt[mn2[c]], 1u);
            }
            bad = __any_sync(FULL, bad) ? 1 : 0;
            if (bad) { if (lane == 0) status[t] = TREE_MALFORMED; continue; }
            __syncwarp();
            {   // cnt[key] <- number of cherries with a LARGER key (where the key's group starts)
                uint32_t carry = 0;
                for (int base = ((k + 1) / 32) * 32; base >= 0; base -= 32) {
                    const int key = base + lane;
                    const uint32_t v = (key <= k + 1) ? cnt[key] : 0u;
                    uint32_t sfx = v;                                // inclusive suffix sum inside the warp
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t y = __shfl_down_sync(FULL, sfx, d);
                        if (lane + d < 32) sfx += y;
                    }
                    if (key <= k + 1) cnt[key] = carry + sfx - v;
                    carry += __shfl_sync(FULL, sfx, 0);
                }
            }
            __syncwarp();
            for (int base = 0; base < k; base += 32) {               // positions in appearance order: stable
                const int c = base + lane;
                const bool act = c < k;
                const uint32_t key = act ? mn2[c] : (0x40000000u + (uint32_t)lane);
                const unsigned peers = __match_any_sync(FULL, key);
                if (act) q[c] = cnt[key] + (uint32_t)__popc(peers & lanes_lt(lane));
                __syncwarp();
                if (act && (peers & lanes_gt(lane)) == 0) cnt[key] += (uint32_t)__popc(peers);
                __syncwarp();
            }
        }
        __syncwarp();
        // ---- 5. rows in matrix order: children relabelled, lengths, duplicates ------------------------
        for (int c = lane; c < k; c += 32) {
            const uint32_t row = q[c];
            if (row >= (uint32_t)k) { bad = 1; continue; }
            if (atomicExch(&mark[row], 1u) != 0u) { bad = 1; continue; }           // two cherries with one parent label
            const uint32_t a = cn1[c], b = cn2[c];
            uint32_t l1 = lab[a], l2 = lab[b];
            if (ks[a] & NM_INT) l1 = (uint32_t)k + 1u + q[cherry_of(a)];
            if (ks[b] & NM_INT) l2 = (uint32_t)k + 1u + q[cherry_of(b)];
            // children must be leaves or earlier rows (parent labels grow towards the root)
            if ((l1 > (uint32_t)k && l1 - (uint32_t)k - 1u >= row) || (l2 > (uint32_t)k && l2 - (uint32_t)k - 1u >= row)) bad = 1;
            anc[3 * row] = (int)l1; anc[3 * row + 1] = (int)l2; anc[3 * row + 2] = k + 1 + (int)row;
            blo[2 * row] = bl[a]; blo[2 * row + 1] = bl[b];
        }
        bad = __any_sync(FULL, bad) ? 1 : 0;
        if (bad) { if (lane == 0) status[t] = TREE_MALFORMED; continue; }
        __syncwarp();
        // ---- 6. min_desc bottom-up over the rows; the two lengths of a row swap when m1 > m2 -----------
        uint32_t* chr = pos;                                         // 2k words: children labels by row
        for (int r = lane; r < k; r += 32) { chr[2 * r] = (uint32_t)anc[3 * r]; chr[2 * r + 1] = (uint32_t)anc[3 * r + 1]; }
        __syncwarp();
        rf_bottom_up<uint32_t>(chr, k, lane, [&](int r, int c1, int c2) {
            const uint32_t m1 = c1 > k ? mn1[c1 - k - 1] : (uint32_t)c1;
            const uint32_t m2 = c2 > k ? mn1[c2 - k - 1] : (uint32_t)c2;
            mn1[r] = min(m1, m2);
            if (!(m1 < m2)) { const double x = blo[2 * r]; blo[2 * r] = blo[2 * r + 1]; blo[2 * r + 1] = x; }
        });
        if (lane == 0) status[t] = 0;
        __syncwarp();
    }
}

// matrix rows [v, bl1, bl2] from the encoder's vector and the ordered lengths; the final status
__global__ void newick_matrix_assemble_kernel(const int* __restrict__ v32, const double* __restrict__ bl, int64_t B, int k,
                                              double* __restrict__ m, const int32_t* __restrict__ st_parse,
                                              const int32_t* __restrict__ st_enc, int32_t* __restrict__ status) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B * k; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = i / k;
        const bool ok = st_parse[t] == 0 && st_enc[t] == 0;
        m[3 * i] = ok ? (double)v32[i] : 0.0;
        m[3 * i + 1] = ok ? bl[2 * i] : 0.0;
        m[3 * i + 2] = ok ? bl[2 * i + 1] : 0.0;
        if (status && i == t * k) status[t] = ok ? 0 : TREE_MALFORMED;
    }
}
// trees the parse rejected get a harmless ancestry so that the encoder has something well-formed to chew on
__global__ void newick_matrix_fix_kernel(int* __restrict__ anc, int64_t B, int k, const int32_t* __restrict__ st) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B * k; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = i / k;
        if (st[t] == 0) continue;
        const int r = (int)(i - t * k);
        anc[3 * i] = r == 0 ? 0 : k + r;                // a ladder: ((0,1),2),3 ...
        anc[3 * i + 1] = r + 1;
        anc[3 * i + 2] = k + 1 + r;
    }
}

// ---------------------------------------------------------------------------------------- launchers
static size_t nm_al(size_t x) { return (x + 255) & ~(size_t)255; }
static int nm_warps_in_flight(size_t words_per_warp) {
    int64_t w = (int64_t)sm_count() * 8 * NM_WARPS;
    const int64_t cap = (int64_t)(((size_t)1 << 30) / (words_per_warp * 4 + 1));     // at most 1 GiB of scratch
    if (w > cap) w = cap;
    w = (w / NM_WARPS) * NM_WARPS;
    return (int)(w < NM_WARPS ? NM_WARPS : w);
}

// to_newick workspace: [ status | int32 v | int32 (k,2) children | decode scratch | per-warp tables ]
struct NmToWs { size_t st, v32, anc, dec, scr, total; };
static NmToWs nm_to_ws(int64_t B, int64_t k) {
    NmToWs w;
    size_t o = 0;
    w.st = o; o += nm_al(4 * (size_t)B);
    w.v32 = o; o += nm_al(4 * (size_t)B * (size_t)k);
    w.anc = o; o += nm_al(8 * (size_t)B * (size_t)k);
    w.dec = o; o += nm_al(decode_workspace_bytes(B, k));
    w.scr = o; o += nm_al((size_t)nm_warps_in_flight(nm_to_words(k)) * nm_to_words(k) * 4);
    w.total = o;
    return w;
}
size_t newick_matrix_workspace_bytes(int64_t B, int64_t k) {
    if (B <= 0 || k <= 0) return 0;
    return nm_to_ws(B, k).total;
}

cudaError_t launch_to_newick_matrix(const double* m, int64_t B, int64_t k, char* out, int64_t out_stride, int64_t* lengths,
                                    int32_t* status, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    if (k <= 0) return cudaErrorInvalidValue;                       // "Matrix should not be empty"
    if (out_stride < (int64_t)newick_matrix_max_bytes(k)) return cudaErrorInvalidValue;
    const NmToWs W = nm_to_ws(B, k);
    if (!ws || ws_bytes < W.total) return cudaErrorInvalidValue;
    unsigned char* wsb = static_cast<unsigned char*>(ws);
    int32_t* st = reinterpret_cast<int32_t*>(wsb + W.st);
    int* v32 = reinterpret_cast<int*>(wsb + W.v32);
    int* anc = reinterpret_cast<int*>(wsb + W.anc);
    cudaError_t e = launch_matrix_v(m, B * k, v32, stream);
    if (e != cudaSuccess) return e;
    const size_t dwb = decode_workspace_bytes(B, k);
    e = launch_decode(MODE_ANC2, v32, 4, B, k, anc, st, dwb ? wsb + W.dec : nullptr, dwb, stream, 0);
    if (e != cudaSuccess) return e;
    // check_m (matrix/convert.rs:28): negative or NaN lengths, on top of the decoder's check_v
    e = launch_check_m(nullptr, m + 1, 3, 3 * k, B, k, st, 1, stream);
    if (e != cudaSuccess) return e;
    const int wif = nm_warps_in_flight(nm_to_words(k));
    const int64_t want = (B + NM_WARPS - 1) / NM_WARPS, cap = wif / NM_WARPS;
    newick_matrix_write_kernel<<<(int)(want < cap ? want : cap), NM_WARPS * 32, 0, stream>>>(
        anc, m, B, (int)k, out, out_stride, lengths, st, reinterpret_cast<uint32_t*>(wsb + W.scr), nm_to_words(k));
    count_launch();
    e = cudaGetLastError();
    if (e == cudaSuccess && status) e = cudaMemcpyAsync(status, st, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, stream);
    return e;
}

// from_newick workspace: [ st_parse | st_enc | int32 (k,3) rows | (k,2) lengths | int32 v | encode scratch | per-warp tables ]
struct NmFromWs { size_t stp, ste, anc, bl, v32, enc, scr, total; size_t words; };
static NmFromWs nm_from_ws(int64_t B, int64_t k) {
    NmFromWs w;
    size_t o = 0;
    w.words = (nm_from_words(k) + 1) & ~(size_t)1;
    w.stp = o; o += nm_al(4 * (size_t)B);
    w.ste = o; o += nm_al(4 * (size_t)B);
    w.anc = o; o += nm_al(12 * (size_t)B * (size_t)k);
    w.bl = o; o += nm_al(16 * (size_t)B * (size_t)k);
    w.v32 = o; o += nm_al(4 * (size_t)B * (size_t)k);
    w.enc = o; o += nm_al(encode_workspace_bytes(B, k));
    w.scr = o; o += nm_al((size_t)nm_warps_in_flight(w.words) * w.words * 4);
    w.total = o;
    return w;
}
size_t from_newick_matrix_workspace_bytes(int64_t B, int64_t k) {
    if (B <= 0 || k <= 0) return 0;
    return nm_from_ws(B, k).total;
}

cudaError_t launch_from_newick_matrix(const char* in, int64_t in_stride, const int64_t* lengths, int64_t B, int64_t k,
                                      double* m, int32_t* status, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (B <= 0 || k <= 0) return cudaSuccess;
    const NmFromWs W = nm_from_ws(B, k);
    if (!ws || ws_bytes < W.total) return cudaErrorInvalidValue;
    unsigned char* wsb = static_cast<unsigned char*>(ws);
    int32_t* stp = reinterpret_cast<int32_t*>(wsb + W.stp);
    int32_t* ste = reinterpret_cast<int32_t*>(wsb + W.ste);
    int* anc = reinterpret_cast<int*>(wsb + W.anc);
    double* bl = reinterpret_cast<double*>(wsb + W.bl);
    int* v32 = reinterpret_cast<int*>(wsb + W.v32);
    const int wif = nm_warps_in_flight(W.words);
    const int64_t want = (B + NM_WARPS - 1) / NM_WARPS, cap = wif / NM_WARPS;
    // strings longer than any well-formed one of this size are rejected without being read to the end
    const int maxlen = (int)(newick_matrix_max_bytes(k) + (size_t)(2 * k) * 400);
    newick_matrix_parse_kernel<<<(int)(want < cap ? want : cap), NM_WARPS * 32, 0, stream>>>(
        in, in_stride, lengths, B, (int)k, anc, bl, stp, maxlen, reinterpret_cast<uint32_t*>(wsb + W.scr), W.words);
    count_launch();
    const int64_t w1 = (B * k + 255) / 256, c1 = (int64_t)sm_count() * 16;
    newick_matrix_fix_kernel<<<(int)(w1 < c1 ? w1 : c1), 256, 0, stream>>>(anc, B, (int)k, stp);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    const size_t ewb = encode_workspace_bytes(B, k);
    e = launch_encode(FROM_ANCESTRY, anc, 4, B, k, v32, ste, ewb ? wsb + W.enc : nullptr, ewb, stream);
    if (e != cudaSuccess) return e;
    newick_matrix_assemble_kernel<<<(int)(w1 < c1 ? w1 : c1), 256, 0, stream>>>(v32, bl, B, (int)k, m, stp, ste, status);
    count_launch();
    return cudaGetLastError();
}

}  // namespace p2v
