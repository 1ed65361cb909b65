// Batched Robinson-Foulds distance between trees for sm_100a.
//
// Replaces (reference, paths relative to /root/reference):
//   vector::distance::robinson_foulds / get_bipartitions / canonicalize_split
//                                                   phylo2vec/src/vector/distance.rs:41-162
//   PyO3 robinson_foulds                            py-phylo2vec/src/lib.rs:295-309
//
// The reference builds one bit set of descendant leaves per internal node (O(n^2) bits), canonicalises
// each split to its smaller side and intersects two hash sets.  Here a pair of trees takes O(n) work
// (Day's algorithm):
//   * a split is named by its side that does NOT contain leaf 0;
//   * T1's leaves are ranked in T1's own DFS order, rotated so that leaf 0 has rank 0: every such side of
//     a T1 split is then a contiguous interval [a, b] of ranks, 1 <= a <= b <= n-1.  The interval of
//     every non-trivial split goes into an open-addressing hash set (key a * n + b);
//   * for T2, every node's descendant set gets (min rank, max rank, count) bottom-up, and the same for
//     everything OUTSIDE the node top-down (the side without leaf 0 of a node that contains leaf 0);
//     a side is an interval iff max - min + 1 == count, and then one hash probe tells whether T1 has it.
//   A binary tree of n >= 4 leaves has exactly n - 3 distinct non-trivial splits (the two children of
//   the root name the same one: the one containing leaf 0 is skipped), hence
//       RF = 2 (n - 3) - 2 shared,      normalised: RF / (2 (n - 3)).
//
// One warp per pair.  The tree passes walk the ancestry rows 32 at a time in row order (children come
// before parents); a row whose child / parent row lies in the same batch waits for it, the batch is
// relaxed until every lane has fired (a few rounds on random trees, 32 on a ladder).  k <= RF_SMEM_K_MAX:
// all tables in shared memory as uint16; larger trees: the same code on uint32 tables in the workspace.
#include <cstdlib>

#include "p2v_common.cuh"
#include "p2v_internal.h"
#include "p2v_tree_pass.cuh"

namespace p2v {

constexpr int RF_SMEM_K_MAX = 4095;
constexpr int RF_WARPS = 4;

template <typename ST> struct RfNone;
template <> struct RfNone<uint16_t> { static constexpr uint32_t v = 0xFFFFu; };
template <> struct RfNone<uint32_t> { static constexpr uint32_t v = 0xFFFFFFFFu; };

// elements of ST one pair needs, then the hash table (uint32 entries, a power of two >= 2n)
__host__ __device__ inline size_t rf_elems(int64_t k) { return (size_t)(3 * (k + 1) + 9 * k + 32); }
__host__ __device__ inline uint32_t rf_hash_size(int64_t k) {
    uint32_t h = 64;
    while (h < 2u * (uint32_t)(k + 1)) h <<= 1;
    return h;
}

__device__ __forceinline__ uint32_t rf_slot(uint32_t key, uint32_t mask) { return (key * 2654435761u) >> 7 & mask; }

// anc: (k,2) int32 rows [child1, child2] of the two trees (the third column is k+1+row).
template <typename ST>
__device__ __forceinline__ double rf_pair_warp(const int2* __restrict__ anc1, const int2* __restrict__ anc2, int k,
                                               ST* scr, uint32_t* hash, uint32_t hsize, int lane, int normalize) {
    constexpr uint32_t NONE = RfNone<ST>::v;
    const int n = k + 1;
    // layout: rank[n] | ch[2k] par[k] | a[k] b[k] c[k] | oa[k] ob[k] oc[k]   (T1 uses a = size, b = left end)
    ST* rank = scr;
    ST* ch = rank + n + 1;
    ST* par = ch + 2 * k;
    ST* va = par + k;
    ST* vb = va + k;
    ST* vc = vb + k;
    ST* oa = vc + k;
    ST* ob = oa + k;
    ST* oc = ob + k;
    const uint32_t hmask = hsize - 1;
    for (uint32_t i = lane; i < hsize; i += 32) hash[i] = 0xFFFFFFFFu;

    // ---------------- T1: sizes bottom-up, left ends top-down, ranks of the leaves ----------------
    for (int r = lane; r < k; r += 32) {
        const int2 x = anc1[r];
        ch[2 * r] = (ST)x.x; ch[2 * r + 1] = (ST)x.y;
        if (x.x > k) par[x.x - k - 1] = (ST)r;
        if (x.y > k) par[x.y - k - 1] = (ST)r;
    }
    __syncwarp();
    rf_bottom_up<ST>(ch, k, lane, [&](int r, int c1, int c2) {
        va[r] = (ST)((c1 > k ? (int)va[c1 - k - 1] : 1) + (c2 > k ? (int)va[c2 - k - 1] : 1));
    });
    __syncwarp();
    rf_top_down<ST>(par, k, lane, [&](int r, int p) {
        int left = 0;
        if (p >= 0) {
            const int pc1 = (int)ch[2 * p];
            left = (int)vb[p];
            if (pc1 != k + 1 + r) left += (pc1 > k ? (int)va[pc1 - k - 1] : 1);        // second child: after the first
        }
        vb[r] = (ST)left;
        const int c1 = (int)ch[2 * r], c2 = (int)ch[2 * r + 1];
        const int s1 = c1 > k ? (int)va[c1 - k - 1] : 1;
        if (c1 <= k) rank[c1] = (ST)left;
        if (c2 <= k) rank[c2] = (ST)(left + s1);
    });
    __syncwarp();
    const int p0 = (int)rank[0];                 // unrotated position of leaf 0
    // the side of every non-root node that does not contain leaf 0, as an interval of rotated ranks
    for (int r = lane; r < k - 1; r += 32) {
        const int s = (int)va[r], L = (int)vb[r];
        const bool has0 = (L <= p0) && (p0 < L + s);
        int a, len;
        if (!has0) { a = L - p0; if (a < 0) a += n; len = s; }
        else { a = L + s - p0; len = n - s; }      // the complement starts right after the node's interval
        if (len >= 2 && len <= n - 2) {
            const uint32_t key = (uint32_t)a * (uint32_t)n + (uint32_t)(a + len - 1);
            uint32_t h = rf_slot(key, hmask);
            for (;;) {
                const uint32_t old = atomicCAS(&hash[h], 0xFFFFFFFFu, key);
                if (old == 0xFFFFFFFFu || old == key) break;
                h = (h + 1) & hmask;
            }
        }
    }
    __syncwarp();
    // rotated rank of every leaf
    for (int c = lane; c <= k; c += 32) { int x = (int)rank[c] - p0; if (x < 0) x += n; rank[c] = (ST)x; }
    __syncwarp();

    // ---------------- T2: (min, max, count) inside bottom-up, outside top-down ---------------------
    for (int r = lane; r < k; r += 32) {
        const int2 x = anc2[r];
        ch[2 * r] = (ST)x.x; ch[2 * r + 1] = (ST)x.y;
        if (x.x > k) par[x.x - k - 1] = (ST)r;
        if (x.y > k) par[x.y - k - 1] = (ST)r;
    }
    __syncwarp();
    rf_bottom_up<ST>(ch, k, lane, [&](int r, int c1, int c2) {
        int mn1, mx1, n1, mn2, mx2, n2;
        if (c1 > k) { mn1 = (int)va[c1 - k - 1]; mx1 = (int)vb[c1 - k - 1]; n1 = (int)vc[c1 - k - 1]; }
        else { mn1 = mx1 = (int)rank[c1]; n1 = 1; }
        if (c2 > k) { mn2 = (int)va[c2 - k - 1]; mx2 = (int)vb[c2 - k - 1]; n2 = (int)vc[c2 - k - 1]; }
        else { mn2 = mx2 = (int)rank[c2]; n2 = 1; }
        va[r] = (ST)min(mn1, mn2); vb[r] = (ST)max(mx1, mx2); vc[r] = (ST)(n1 + n2);
    });
    __syncwarp();
    int shared_cnt = 0;
    rf_top_down<ST>(par, k, lane, [&](int r, int p) {
        // outside(r) = outside(parent) + inside(sibling); the root's outside is empty (count 0)
        int omn = (int)NONE, omx = 0, ocn = 0;
        if (p >= 0) {
            const int pc1 = (int)ch[2 * p], pc2 = (int)ch[2 * p + 1];
            const int sib = (pc1 == k + 1 + r) ? pc2 : pc1;
            int smn, smx, sn;
            if (sib > k) { smn = (int)va[sib - k - 1]; smx = (int)vb[sib - k - 1]; sn = (int)vc[sib - k - 1]; }
            else { smn = smx = (int)rank[sib]; sn = 1; }
            ocn = (int)oc[p] + sn;
            omn = ((int)oc[p] > 0) ? min((int)oa[p], smn) : smn;
            omx = ((int)oc[p] > 0) ? max((int)ob[p], smx) : smx;
        }
        oa[r] = (ST)omn; ob[r] = (ST)omx; oc[r] = (ST)ocn;
        if (p < 0) return;                                    // the root names no split
        const bool has0 = (int)va[r] == 0;                    // rotated rank 0 is leaf 0
        if (has0 && p == k - 1) return;                       // root child with leaf 0: its split is its sibling's
        const int a = has0 ? omn : (int)va[r], b = has0 ? omx : (int)vb[r], cnt = has0 ? ocn : (int)vc[r];
        if (cnt < 2 || cnt > n - 2 || b - a + 1 != cnt) return;
        const uint32_t key = (uint32_t)a * (uint32_t)n + (uint32_t)b;
        uint32_t h = rf_slot(key, hmask);
        for (;;) {
            const uint32_t cur = hash[h];
            if (cur == key) { ++shared_cnt; break; }
            if (cur == 0xFFFFFFFFu) break;
            h = (h + 1) & hmask;
        }
    });
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) shared_cnt += __shfl_xor_sync(FULL, shared_cnt, o);
    const double max_rf = 2.0 * (double)(n - 3);
    const double rf = max_rf - 2.0 * (double)shared_cnt;
    return normalize ? __ddiv_rn(rf, max_rf) : rf;
}

// pair p: tree i1 = (B1 == 1 ? 0 : p) of the first batch against tree i2 = (B2 == 1 ? 0 : p) of the second
template <typename ST, bool SMEM>
__global__ void __launch_bounds__(RF_WARPS * 32)
rf_kernel(const int* __restrict__ anc1, const int* __restrict__ anc2, int64_t B1, int64_t B2, int k, int normalize,
          double* __restrict__ out, const int32_t* __restrict__ st1, const int32_t* __restrict__ st2,
          int32_t* __restrict__ status, unsigned char* __restrict__ scratch, size_t scratch_stride) {
    extern __shared__ __align__(16) unsigned char rf_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int64_t P = B1 > B2 ? B1 : B2;
    const uint32_t hsize = rf_hash_size(k);
    const size_t per = (rf_elems(k) * sizeof(ST) + 15) / 16 * 16 + (size_t)hsize * 4;
    unsigned char* mine = SMEM ? rf_dyn + (size_t)warp * per
                               : scratch + ((size_t)blockIdx.x * nwarps + warp) * scratch_stride;
    ST* scr = reinterpret_cast<ST*>(mine);
    uint32_t* hash = reinterpret_cast<uint32_t*>(mine + (rf_elems(k) * sizeof(ST) + 15) / 16 * 16);
    for (int64_t p = (int64_t)blockIdx.x * nwarps + warp; p < P; p += (int64_t)gridDim.x * nwarps) {
        const int64_t i1 = B1 == 1 ? 0 : p, i2 = B2 == 1 ? 0 : p;
        const int s1 = st1[i1], s2 = st2[i2];
        if (s1 != 0 || s2 != 0) {                 // an invalid vector on either side
            if (lane == 0) { out[p] = 0.0; if (status) status[p] = s1 != 0 ? s1 : s2; }
            continue;
        }
        const double v = rf_pair_warp<ST>(reinterpret_cast<const int2*>(anc1 + (size_t)i1 * 2 * k),
                                          reinterpret_cast<const int2*>(anc2 + (size_t)i2 * 2 * k), k, scr, hash, hsize,
                                          lane, normalize);
        if (lane == 0) { out[p] = v; if (status) status[p] = 0; }
        __syncwarp();
    }
}

__global__ void rf_trivial_kernel(int64_t P, double* __restrict__ out, int32_t* __restrict__ status) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (int64_t)gridDim.x * blockDim.x) {
        out[p] = 0.0;                              // n <= 3: no non-trivial split (distance.rs:50-53)
        if (status) status[p] = 0;
    }
}

static size_t rf_al(size_t x) { return (x + 255) & ~(size_t)255; }
static size_t rf_pair_bytes(int64_t k, bool smem) {
    const size_t es = smem ? 2 : 4;
    return (rf_elems(k) * es + 15) / 16 * 16 + (size_t)rf_hash_size(k) * 4;
}
static int rf_global_warps() { return sm_count() * 8 * RF_WARPS; }

// workspace: [ st1 | st2 | v32 x2 | anc2 of batch 1 | anc2 of batch 2 | decode scratch | per-warp tables (large k) ]
struct RfWs { size_t st1, st2, v1, v2, a1, a2, dec, scr, total; };
static RfWs rf_ws(int64_t B1, int64_t B2, int64_t k) {
    RfWs w;
    size_t o = 0;
    const int64_t Bm = B1 > B2 ? B1 : B2;
    w.st1 = o; o += rf_al(4 * (size_t)B1);
    w.st2 = o; o += rf_al(4 * (size_t)B2);
    w.v1 = o; o += rf_al(4 * (size_t)B1 * (size_t)k);
    w.v2 = o; o += rf_al(4 * (size_t)B2 * (size_t)k);
    w.a1 = o; o += rf_al(8 * (size_t)B1 * (size_t)k);
    w.a2 = o; o += rf_al(8 * (size_t)B2 * (size_t)k);
    w.dec = o; o += rf_al(decode_workspace_bytes(Bm, k));
    w.scr = o; o += (k > RF_SMEM_K_MAX) ? rf_al(rf_pair_bytes(k, false)) * (size_t)rf_global_warps() : 0;
    w.total = o;
    return w;
}
size_t robinson_foulds_workspace_bytes(int64_t B1, int64_t B2, int64_t k) {
    if (B1 <= 0 || B2 <= 0 || k < 3) return 0;
    return rf_ws(B1, B2, k).total;
}

cudaError_t launch_robinson_foulds(const void* v1, const void* v2, int idx_bytes, int64_t B1, int64_t B2, int64_t k,
                                   int normalize, double* out, int32_t* status, void* ws, size_t ws_bytes,
                                   cudaStream_t stream) {
    if (B1 <= 0 || B2 <= 0) return cudaSuccess;
    if (B1 != B2 && B1 != 1 && B2 != 1) return cudaErrorInvalidValue;
    if (k > 65535) return cudaErrorNotSupported;        // interval keys a * n + b are 32-bit
    const int64_t P = B1 > B2 ? B1 : B2;
    if (k < 3) {                                   // at most 3 leaves: RF = 0 whatever the vectors are
        rf_trivial_kernel<<<(int)((P + 255) / 256 < 1024 ? (P + 255) / 256 : 1024), 256, 0, stream>>>(P, out, status);
        count_launch();
        return cudaGetLastError();
    }
    const RfWs W = rf_ws(B1, B2, k);
    if (!ws || ws_bytes < W.total) return cudaErrorInvalidValue;
    unsigned char* wsb = static_cast<unsigned char*>(ws);
    cudaError_t e;
    const void* vin[2] = {v1, v2};
    const int64_t Bs[2] = {B1, B2};
    int32_t* sts[2] = {reinterpret_cast<int32_t*>(wsb + W.st1), reinterpret_cast<int32_t*>(wsb + W.st2)};
    int* v32[2] = {reinterpret_cast<int*>(wsb + W.v1), reinterpret_cast<int*>(wsb + W.v2)};
    int* anc[2] = {reinterpret_cast<int*>(wsb + W.a1), reinterpret_cast<int*>(wsb + W.a2)};
    for (int t = 0; t < 2; ++t) {
        const void* vi = vin[t];
        if (idx_bytes == 8) {
            e = launch_narrow_v(static_cast<const long long*>(vin[t]), Bs[t] * k, v32[t], stream);
            if (e != cudaSuccess) return e;
            vi = v32[t];
        } else if (idx_bytes != 4) {
            return cudaErrorInvalidValue;
        }
        const size_t dwb = decode_workspace_bytes(Bs[t], k);
        e = launch_decode(MODE_ANC2, vi, 4, Bs[t], k, anc[t], sts[t], dwb ? wsb + W.dec : nullptr, dwb, stream, 0);
        if (e != cudaSuccess) return e;
    }
    if (k <= RF_SMEM_K_MAX) {
        const size_t per = rf_pair_bytes(k, true);
        const int wpc = (per <= (size_t)8 * 1024) ? RF_WARPS : 1;      // big tables: one warp per CTA, more CTAs per SM
        const size_t smem = per * wpc;
        auto kern = rf_kernel<uint16_t, true>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, wpc * 32, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        const int64_t want = (P + wpc - 1) / wpc, cap = (int64_t)sm_count() * per_sm;
        kern<<<(int)(want < cap ? want : cap), wpc * 32, smem, stream>>>(anc[0], anc[1], B1, B2, (int)k, normalize, out,
                                                                    sts[0], sts[1], status, nullptr, 0);
    } else {
        const int64_t want = (P + RF_WARPS - 1) / RF_WARPS, cap = rf_global_warps() / RF_WARPS;
        rf_kernel<uint32_t, false><<<(int)(want < cap ? want : cap), RF_WARPS * 32, 0, stream>>>(
            anc[0], anc[1], B1, B2, (int)k, normalize, out, sts[0], sts[1], status, wsb + W.scr,
            rf_al(rf_pair_bytes(k, false)));
    }
    count_launch();
    return cudaGetLastError();
}

}  // namespace p2v
