// Cophenetic (leaf-to-leaf path length) matrices for sm_100a.
//
// Replaces (reference, paths relative to /root/reference):
//   vector::graph::_cophenetic_distances   phylo2vec/src/vector/graph.rs:20-90
//   matrix::graph::cophenetic_distances    phylo2vec/src/matrix/graph.rs:23-26
//   matrix::base::parse_matrix             phylo2vec/src/matrix/base.rs:108-121
//
// The reference fills a (2k+1)^2 scratch matrix with ~1.5 k^2 x 4 scattered
// stores and slices the leaf block out of it.  Here the n x n leaf block is
// written once, row by row, with 128-bit stores, and nothing else of size n^2
// is touched:
//   D[i][j] = depth(i) + depth(j) - 2 depth(LCA(i,j)),   D[i][i] = 0,
// "unrooted" zeroes the root edge that leads to node 2k-1 (graph.rs:46-50 with
// the write order at :72-79), k = 0 and k = 1 are the special cases at :28-42.
//
// Pipeline per tree (k >= 2):
//   prepare (independent of the requested rows)
//     1. decode v -> ancestry (p2v_decode.cu)
//     2. tables_kernel: parent pointers; root-path sums (topological depth and
//        double-double branch-length depth) and leftmost / rightmost leaf of every
//        node by pointer jumping; every internal node links the last leaf of its
//        first child to the first leaf of its second, and ranking that leaf list
//        by pointer jumping gives the DFS position of every leaf and, for every
//        DFS position r, the internal node sep[r] that separates leaves r-1 and r.
//        LCA(a,b) for DFS positions a<b is the shallowest sep in (a,b].
//   rows [row_begin,row_end)
//     3. blocks_kernel: the DFS positions are cut into blocks that hold at most
//        ROW_R requested rows and at most ROW_PMAX positions; per block: prefix /
//        suffix minima of sep and the block minimum.
//     4. rows_kernel: one CTA = (block with its <= ROW_R requested rows) x (chunk of
//        1024 columns in label order).  With rl = the row's candidate inside its own
//        block and cl = the column's candidate (inside its block, combined once per
//        CTA with the minimum over the whole blocks in between), the LCA is the
//        shallower of the two and both lie on one root path, so
//            D = (d_i - d_rl) + (d_j - d_cl) + |d_rl - d_cl|
//        -- three non-negative terms, no compare/select on the hot path, and for
//        fp64 no cancellation: the first two are rounded once per row / column, the
//        third is a double-double difference (DESIGN.md "fp64 parity").
//        Columns whose leaf lies inside the row block are patched afterwards by the
//        same CTA from a sparse table over the block's separators.
#include <cooperative_groups.h>

#include <cstdlib>
#include <type_traits>

#include "p2v_decode_warp.cuh"

namespace cg = cooperative_groups;

namespace p2v {

cudaError_t debug_phase_ns_coph(unsigned long long* out) {
    return cudaMemcpyFromSymbol(out, g_phase_ns, sizeof(unsigned long long) * PHASE_SLOTS);
}

constexpr int TAB_THREADS = 1024;
constexpr int TAB_CLUSTER = 16;            // CTAs per tree when there are few, large trees
constexpr int TAB_FLAGS = 128;             // per-tree convergence flags (one per jumping round)
constexpr int BLK_THREADS = 1024;
#ifndef P2V_ROW_R
#define P2V_ROW_R 96
#endif
constexpr int ROW_R = P2V_ROW_R;           // capacity: requested rows per CTA (tile)
constexpr int ROW_RT = 64;                 // rows a tile is meant to hold
constexpr int ROW_PMAX_LARGE = 1024;       // capacity: DFS positions per block at most (n > 512)
constexpr int ROW_PMAX_SMALL = 64;         //                                            (n <= 512)
constexpr int ilog2c(int x) { return x <= 1 ? 0 : 1 + ilog2c(x >> 1); }
constexpr int ROW_CPT = 4;                 // columns per thread (two 16-byte stores per row)
constexpr int ROW_NINCAP = 64;             // in-block columns per chunk whose values are staged in shared memory
// Shape of one rows call.  Block ids follow two lattices: a cut every `pcap` DFS positions and a cut
// every `rcap` requested rows.  All rows requested: blocks of exactly ROW_RT positions = rows.  A row
// subset (one rank's block of a huge tree): the requested rows are spread over the DFS order, so the
// POSITION lattice is the one that shapes the tiles -- pcap = ROW_RT * n / m positions hold ROW_RT
// requested rows on average -- and the row lattice (rcap = ROW_R, 1.5x the mean) only catches dense
// stretches.  (Cutting at 64 rows AND at 1.25x their mean span, as round 1 did, left tiles of 36 rows
// on average: both lattices cut all the time.)  `bsz` bounds the positions in a block (capacity of
// the in-block tables), `nidb` the number of block ids.
static bool tables_smem_on() {            // P2V_TABLES_SMEM=0: the global-memory tables kernel for every size (comparison runs)
    static const int v = [] { const char* e = getenv("P2V_TABLES_SMEM"); return e ? atoi(e) : 1; }();
    return v != 0;
}
struct CophCall { int pcap, rcap, bsz, nidb, nincap; };
inline CophCall coph_call(int64_t n, int64_t m) {
    CophCall c;
    if (m >= n) {
        c.rcap = (n <= 512) ? 32 : ROW_RT;
        c.pcap = 1 << 30; c.bsz = c.rcap;
        c.nincap = c.bsz < ROW_NINCAP ? c.bsz : ROW_NINCAP;
    } else {
        const int rt = (n <= 512) ? 32 : ROW_RT;
        c.rcap = (n <= 512) ? 32 : ROW_R;
        int64_t want = ((int64_t)rt * n) / (m > 0 ? m : 1);
        want = (want + 31) & ~(int64_t)31;
        if (want < rt) want = rt;
        const int64_t cap = (n <= 512) ? ROW_PMAX_SMALL : ROW_PMAX_LARGE;
        if (want > cap) want = cap;
        c.pcap = (int)want; c.bsz = c.pcap;
        // a chunk of 1024 label-ordered columns holds about 1024 * bsz / n of the block's own leaves
        c.nincap = (n > 512 && c.bsz * 1024 <= 16 * n) ? 32 : (c.bsz < ROW_NINCAP ? c.bsz : ROW_NINCAP);
    }
    c.nidb = (int)(n / c.pcap + (m + c.rcap - 1) / c.rcap + 2);
    return c;
}
// threads per rows-CTA: 4 columns per thread, so small trees get small CTAs
inline int coph_row_threads(int n) { return n <= 128 ? 32 : (n <= 256 ? 64 : (n <= 512 ? 128 : 256)); }
constexpr int KEY_INF = 0x3FFFFFFF;

struct CophLayout {
    // byte offsets inside one tree's table area
    size_t anc, jAD, jDD, jDI, jDH, jDL, tn, tc, pos, leaf_at, sepk, sephi, seplo;     // prepare
    size_t blk_of, rowlist, prek, prea, sufk, sufa, bstart, bend, rowoff, cend, blkk, blka, tilelist, hdr, flags;  // rows call (+ flags: prepare)
    size_t c_blk, c_prek, c_sufk, c_prehi, c_prelo, c_sufhi, c_suflo;      // rows call, by leaf label
    size_t total;
    int n, N, A, nbmax;
};

__host__ __device__ inline size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }

__host__ __device__ inline CophLayout coph_layout(int64_t k) {
    CophLayout L;
    L.n = (int)(k + 1); L.N = (int)(2 * k + 1); L.A = (int)(4 * k);
    L.nbmax = L.n / 32 + L.n / 32 + 4;     // bound on block ids for any call (pcap >= 32, rcap >= 32)
    size_t o = 0;
    L.anc = o; o = al16(o + 4 * 3 * (size_t)k);
    L.jAD = o; o = al16(o + 8 * 2 * (size_t)L.N);      // jumping buffers: (ancestor, int depth) x 2
    L.jDD = o; o = al16(o + 16 * 2 * (size_t)L.N);     //                  double-double depth x 2
    L.jDI = o; o = al16(o + 4 * (size_t)L.N);          // results: depth of every node (int, dd hi, dd lo)
    L.jDH = o; o = al16(o + 8 * (size_t)L.N);
    L.jDL = o; o = al16(o + 8 * (size_t)L.N);
    L.tn = o; o = al16(o + 4 * (2 * (size_t)L.A + 16));   // dl x 2 buffers, then the (next leaf, count) list x 2
    L.tc = o; o = al16(o + 4 * 2 * (size_t)L.A);
    L.pos = o; o = al16(o + 4 * (size_t)L.n);
    L.leaf_at = o; o = al16(o + 4 * (size_t)L.n);
    L.sepk = o; o = al16(o + 4 * (size_t)(L.n + 1));
    L.sephi = o; o = al16(o + 8 * (size_t)(L.n + 1));
    L.seplo = o; o = al16(o + 8 * (size_t)(L.n + 1));
    L.blk_of = o; o = al16(o + 4 * (size_t)L.n);
    L.rowlist = o; o = al16(o + 4 * (size_t)L.n);
    L.prek = o; o = al16(o + 4 * (size_t)L.n);
    L.prea = o; o = al16(o + 4 * (size_t)L.n);
    L.sufk = o; o = al16(o + 4 * (size_t)L.n);
    L.sufa = o; o = al16(o + 4 * (size_t)L.n);
    L.bstart = o; o = al16(o + 4 * (size_t)L.nbmax);
    L.bend = o; o = al16(o + 4 * (size_t)L.nbmax);
    L.rowoff = o; o = al16(o + 4 * (size_t)L.nbmax);
    L.cend = o; o = al16(o + 4 * (size_t)L.nbmax);
    L.blkk = o; o = al16(o + 4 * (size_t)L.nbmax);
    L.blka = o; o = al16(o + 4 * (size_t)L.nbmax);
    L.tilelist = o; o = al16(o + 4 * (size_t)L.nbmax);
    L.c_blk = o; o = al16(o + 4 * (size_t)L.n);
    L.c_prek = o; o = al16(o + 4 * (size_t)L.n);
    L.c_sufk = o; o = al16(o + 4 * (size_t)L.n);
    L.c_prehi = o; o = al16(o + 8 * (size_t)L.n);
    L.c_prelo = o; o = al16(o + 8 * (size_t)L.n);
    L.c_sufhi = o; o = al16(o + 8 * (size_t)L.n);
    L.c_suflo = o; o = al16(o + 8 * (size_t)L.n);
    L.hdr = o; o = al16(o + 64);
    L.flags = o; o = al16(o + 4 * (size_t)TAB_FLAGS);
    L.total = al16(o + 64);
    return L;
}

// ---- input adapters ----------------------------------------------------------------
__global__ void narrow_v_kernel(const long long* __restrict__ v, int64_t count, int* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        const long long x = v[i];
        out[i] = (x < 0 || x > 0x7FFFFFFFLL) ? -1 : (int)x;  // -1 fails check_v downstream
    }
}

cudaError_t launch_narrow_v(const long long* v, int64_t count, int* out, cudaStream_t stream) {
    if (count <= 0) return cudaSuccess;
    narrow_v_kernel<<<sm_count() * 8, 256, 0, stream>>>(v, count, out);
    count_launch();
    return cudaGetLastError();
}

// parse_matrix (matrix/base.rs:108-121): v = m[:,0] as usize (truncating, saturating cast)
__global__ void matrix_v_kernel(const double* __restrict__ m, int64_t count, int* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double x = m[3 * i];
        int r;
        if (!(x > 0.0)) r = 0;                    // negatives and NaN saturate to 0
        else if (x >= 2147483647.0) r = 0x7FFFFFFF;
        else r = (int)x;
        out[i] = r;
    }
}

// check_m (matrix/base.rs:76-95), one warp per tree.  bl == m + 1, bl_stride == 3 for a (k,3) matrix.
//   v_from_matrix: also check_v on column 0 cast like `as usize` (negatives and NaN -> 0, too big
//   saturates): status = first i with v[i] > 2i, + 1.
//   Lengths: any bl < 0 or NaN ("Branch lengths must be positive") -> P2V_TREE_NEGATIVE_BRANCH,
//   only for trees whose status is still 0 (keep_status: the v part was checked by the decoder).
__global__ void check_m_kernel(const double* __restrict__ m, const double* __restrict__ bl, int bl_stride,
                               int64_t bl_tree_stride, int64_t B, int k, int32_t* __restrict__ status,
                               int keep_status) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t tree = wid; tree < B; tree += nwarps) {
        int st = keep_status ? status[tree] : 0;
        if (m && st == 0) {
            const double* mt = m + tree * 3 * (int64_t)k;
            for (int b = 0; b < k; b += 32) {
                const int i = b + lane;
                const bool bad = (i < k) && (mt[3 * (size_t)i] >= 2.0 * i + 1.0);     // (x as usize) > 2i
                const unsigned bm = __ballot_sync(FULL, bad);
                if (bm) { st = b + __ffs(bm); break; }
            }
        }
        if (bl && st == 0) {
            const double* bt = bl + tree * bl_tree_stride;
            bool neg = false;
            for (int i = lane; i < k; i += 32) {
                const double b1 = bt[(size_t)i * bl_stride], b2 = bt[(size_t)i * bl_stride + 1];
                neg |= !(b1 >= 0.0 && b2 >= 0.0);
            }
            if (__any_sync(FULL, neg)) st = -3;   // P2V_TREE_NEGATIVE_BRANCH
        }
        if (lane == 0) status[tree] = st;
    }
}

// The distance path's own guard: a tree with a negative or NaN branch length gets
// P2V_TREE_NEGATIVE_BRANCH (if its vector was valid) instead of a silently wrong matrix -- the
// three-term distance assumes depth grows downwards.  One thread per row of the batch.
__global__ void flag_negative_bl_kernel(const double* __restrict__ bl, int bl_stride, int64_t bl_tree_stride,
                                        int64_t B, int k, int32_t* __restrict__ status) {
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < B * k; w += (int64_t)gridDim.x * blockDim.x) {
        const int64_t tree = w / k;
        const int i = (int)(w - tree * k);
        const double* bt = bl + tree * bl_tree_stride + (size_t)i * bl_stride;
        if (!(bt[0] >= 0.0 && bt[1] >= 0.0)) atomicCAS(status + tree, 0, -3);
    }
}

cudaError_t launch_check_m(const double* m, const double* bl, int bl_stride, int64_t bl_tree_stride, int64_t B,
                           int64_t k, int32_t* status, int keep_status, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    const int64_t want = (B * 32 + 255) / 256, cap = (int64_t)sm_count() * 8;
    check_m_kernel<<<(int)(want < cap ? want : cap), 256, 0, stream>>>(m, bl, bl_stride, bl_tree_stride, B, (int)k,
                                                                      status, keep_status);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_matrix_v(const double* m, int64_t count, int* out, cudaStream_t stream) {
    if (count <= 0) return cudaSuccess;
    matrix_v_kernel<<<sm_count() * 8, 256, 0, stream>>>(m, count, out);
    count_launch();
    return cudaGetLastError();
}

// ---- k < 2 special cases (graph.rs:28-42) ---------------------------------------------
__global__ void tiny_sum_kernel(const double* __restrict__ bl, int64_t bl_tree_stride, int64_t B,
                                double* __restrict__ sums) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < B; t += (int64_t)gridDim.x * blockDim.x)
    {
        sums[2 * t] = bl ? bl[t * bl_tree_stride] : 1.0;
        sums[2 * t + 1] = bl ? bl[t * bl_tree_stride + 1] : 1.0;
    }
}

// what = 1: variance-covariance (graph.rs:211-258 with k = 1: diag = the two lengths, 0 elsewhere)
__global__ void tiny_kernel(const double* __restrict__ sums, int64_t B, int k, int64_t row_begin,
                            int64_t row_end, double* __restrict__ out, int what) {
    const int n = k + 1;
    const int64_t rows = row_end - row_begin;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < B; t += (int64_t)gridDim.x * blockDim.x) {
        const double s = (k == 1 && !what) ? sums[2 * t] + sums[2 * t + 1] : 0.0;
        for (int64_t r = row_begin; r < row_end; ++r)
            for (int c = 0; c < n; ++c) {
                double x = (r == c) ? 0.0 : s;
                if (what && r == c && k == 1) x = sums[2 * t + r];
                out[(t * rows + (r - row_begin)) * n + c] = x;
            }
    }
}

// node depths (vector/ops.rs:201-233): depth of every node 0..2k from the prepared tables
__global__ void node_depths_kernel(const unsigned char* __restrict__ tab, size_t tab_stride, int64_t B, int k,
                                   int weighted, double* __restrict__ out, const int32_t* __restrict__ status) {
    const int N = 2 * k + 1;
    if (k < 2) {    // the table area holds the two root edge lengths (k == 1) or nothing (k == 0)
        const double* sums = reinterpret_cast<const double*>(tab);
        for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < B; t += (int64_t)gridDim.x * blockDim.x) {
            if (status && status[t] != 0) continue;
            if (k == 1) { out[3 * t] = sums[2 * t]; out[3 * t + 1] = sums[2 * t + 1]; out[3 * t + 2] = 0.0; }
            else out[t] = 0.0;
        }
        return;
    }
    const CophLayout L = coph_layout(k);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B * N; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = i / N;
        const int u = (int)(i - t * N);
        if (status && status[t] != 0) continue;
        const unsigned char* T = tab + (size_t)t * tab_stride;
        out[i] = weighted ? reinterpret_cast<const double*>(T + L.jDH)[u]
                          : (double)reinterpret_cast<const int*>(T + L.jDI)[u];
    }
}

// ---- 2. per-tree tables ------------------------------------------------------------------
// A team of CTAs works on one tree: one CTA per tree for batches (CSIZE = 1); for a few large trees
// a thread-block cluster of TAB_CLUSTER CTAs (the jumping rounds spread over 16 SMs, cluster
// barriers), and for one to four huge trees the whole cooperatively launched grid (CSIZE = 0: every
// round is bound by scattered L2 accesses per SM, so all 148 SMs take a tree together and the trees
// go one after the other).
// The pointer jumping is radix 4: a round composes the current jump four times (three dependent
// gathers), so root-path sums and the leaf ranking take ceil(log4) rounds -- half the barriers of
// the radix-2 form for 1.5x the gathers.  (ancestor, int depth) travel as one int2 and the
// double-double depth as one double2; the latter only exist for weighted trees.
template <int CSIZE>
__global__ void __launch_bounds__(TAB_THREADS)
tables_kernel(unsigned char* __restrict__ tab, size_t tab_stride, const double* __restrict__ bl,
              int bl_stride, int64_t bl_tree_stride, int64_t B, int k, int unrooted,
              const int32_t* __restrict__ status) {
    const CophLayout L = coph_layout(k);
    const int crank = team_rank<CSIZE>();
    const int tid = crank * TAB_THREADS + threadIdx.x;       // thread index inside the team
    const int TT = team_ctas<CSIZE>() * TAB_THREADS;          // threads per team
    const int n = L.n, N = L.N, root = 2 * k;
    const bool weighted = bl != nullptr;
    const int64_t team = team_index<CSIZE>(), nteams = team_count<CSIZE>();
    for (int64_t tree = team; tree < B; tree += nteams) {
        if (status && status[tree] != 0) continue;   // uniform per team
        unsigned char* T = tab + (size_t)tree * tab_stride;
        const int* anc = reinterpret_cast<const int*>(T + L.anc);
        int2* jAD = reinterpret_cast<int2*>(T + L.jAD);
        double2* jDD = reinterpret_cast<double2*>(T + L.jDD);
        int* jDI = reinterpret_cast<int*>(T + L.jDI);
        double* jDH = reinterpret_cast<double*>(T + L.jDH);
        double* jDL = reinterpret_cast<double*>(T + L.jDL);
        int* tn = reinterpret_cast<int*>(T + L.tn);
        int* tc = reinterpret_cast<int*>(T + L.tc);
        int* pos = reinterpret_cast<int*>(T + L.pos);
        int* leaf_at = reinterpret_cast<int*>(T + L.leaf_at);
        int* sepk = reinterpret_cast<int*>(T + L.sepk);
        double* sephi = reinterpret_cast<double*>(T + L.sephi);
        double* seplo = reinterpret_cast<double*>(T + L.seplo);
        int* flags = reinterpret_cast<int*>(T + L.flags);
        const double* blt = bl ? bl + tree * bl_tree_stride : nullptr;

        phase_stamp(15, CSIZE == 0 && tid == 0);
        if (tid < TAB_FLAGS) flags[tid] = 0;
        // parents and edge weights (row r of bls: lengths above ancestry[r][0], [r][1];
        // graph.rs:57-58).  Buffer 0 of jAD / jDD = initial state of the jumping.
        // first / second child pointers for the leftmost / rightmost-leaf jumping (leaves: themselves).
        // dl = tn[0 .. 2N), dr = tc[0 .. 2N) (two buffers each); the leaf list lives behind them.
        int* dl = tn; int* dr = tc;
        int2* lnc = reinterpret_cast<int2*>(tn + 2 * N);     // 2 buffers x n (next leaf, count): 2N + 4n <= 8k
        for (int r = tid; r < k; r += TT) {
            const int c1 = anc[3 * r], c2 = anc[3 * r + 1], p = k + 1 + r;
            int w1 = 1, w2 = 1;
            if (unrooted) {
                if (c1 == 2 * k - 1) w1 = 0;
                if (c2 == 2 * k - 1) w2 = 0;
            }
            jAD[c1] = make_int2(p, w1); jAD[c2] = make_int2(p, w2);
            if (weighted) {
                const double b1 = w1 ? blt[(size_t)r * bl_stride] : 0.0;
                const double b2 = w2 ? blt[(size_t)r * bl_stride + 1] : 0.0;
                jDD[c1] = make_double2(b1, 0.0); jDD[c2] = make_double2(b2, 0.0);
            }
            dl[p] = c1; dr[p] = c2;
        }
        for (int c = tid; c <= k; c += TT) { dl[c] = c; dr[c] = c; }
        if (tid == 0) { jAD[root] = make_int2(-1, 0); if (weighted) jDD[root] = make_double2(0.0, 0.0); }
        team_sync<CSIZE>();
        phase_stamp(0, CSIZE == 0 && tid == 0);
        // root-path sums and leftmost / rightmost leaves by pointer jumping
        int fround = 0;
        int cur = 0;
        for (;; ++fround) {
            const int2* a_in = jAD + cur * N; int2* a_out = jAD + (cur ^ 1) * N;
            const double2* d_in = jDD + cur * N; double2* d_out = jDD + (cur ^ 1) * N;
            const int* dl_in = dl + cur * N; int* dl_out = dl + (cur ^ 1) * N;
            const int* dr_in = dr + cur * N; int* dr_out = dr + (cur ^ 1) * N;
            int live = 0;
#pragma unroll 2
            for (int u = tid; u < N; u += TT) {
                int nl = dl_in[u], nr = dr_in[u];               // leaves are fixed points
                int2 s = a_in[u];
                dd d = {0.0, 0.0};
                if (weighted) { const double2 x = d_in[u]; d = dd{x.x, x.y}; }
#pragma unroll
                for (int hop = 0; hop < 3; ++hop) {
                    nl = dl_in[nl]; nr = dr_in[nr];
                    if (s.x >= 0) {
                        const int a = s.x;
                        const int2 t = a_in[a];
                        if (weighted) { const double2 x = d_in[a]; d = dd_add(d, dd{x.x, x.y}); }
                        s.x = t.x; s.y += t.y;
                    }
                }
                live |= (nl > k) | (nr > k) | (s.x >= 0);
                dl_out[u] = nl; dr_out[u] = nr;
                a_out[u] = s;
                if (weighted) d_out[u] = make_double2(d.hi, d.lo);
            }
            cur ^= 1;
            const int any = team_any<CSIZE>(live, flags + min(fround, TAB_FLAGS - 1));
            if (!any) { ++fround; break; }
        }
        phase_stamp(1, CSIZE == 0 && tid == 0);
        if (CSIZE == 0 && tid == 0) g_phase_ns[16] = (unsigned long long)fround;
        dl += cur * N; dr += cur * N;
        {   // depth of every node, where the other kernels read it
            const int2* a_res = jAD + cur * N;
            const double2* d_res = jDD + cur * N;
            for (int u = tid; u < N; u += TT) {
                jDI[u] = a_res[u].y;
                if (weighted) { const double2 x = d_res[u]; jDH[u] = x.x; jDL[u] = x.y; }
            }
        }
        // The leaves as a linked list in DFS order: every internal node links the last leaf of its
        // first child's subtree to the first leaf of its second child's (and is the separator of that
        // pair); ranking the list by pointer jumping gives the DFS positions.
        for (int r = tid; r < k; r += TT) {
            const int a = dr[anc[3 * r]], b = dl[anc[3 * r + 1]];
            lnc[a] = make_int2(b, 1);
        }
        if (tid == 0) lnc[dr[root]] = make_int2(-1, 0);
        team_sync<CSIZE>();
        int lc = 0;
        for (;; ++fround) {
            const int2* in = lnc + lc * n; int2* out = lnc + (lc ^ 1) * n;
            int live = 0;
#pragma unroll 2
            for (int x = tid; x < n; x += TT) {
                int2 s = in[x];
#pragma unroll
                for (int hop = 0; hop < 3; ++hop) {
                    if (s.x >= 0) { const int2 t = in[s.x]; s.x = t.x; s.y += t.y; }
                }
                live |= (s.x >= 0);
                out[x] = s;
            }
            lc ^= 1;
            const int any = team_any<CSIZE>(live, flags + min(fround, TAB_FLAGS - 1));
            if (!any) break;
        }
        phase_stamp(2, CSIZE == 0 && tid == 0);
        if (CSIZE == 0 && tid == 0) g_phase_ns[17] = (unsigned long long)fround;
        const int2* after = lnc + lc * n;             // .y = number of leaves after this one
        // DFS positions of the leaves; separators by position
        for (int c = tid; c <= k; c += TT) {
            const int p = k - after[c].y;
            pos[c] = p;
            leaf_at[p] = c;
        }
        if (tid == 0) {
            sepk[0] = KEY_INF; sepk[n] = KEY_INF;
            if (weighted) { sephi[0] = 0.0; seplo[0] = 0.0; sephi[n] = 0.0; seplo[n] = 0.0; }
        }
        {
            const int2* a_res = jAD + cur * N;
            const double2* d_res = jDD + cur * N;
            for (int u = k + 1 + tid; u <= 2 * k; u += TT) {
                const int c1 = anc[3 * (u - k - 1) + 1];      // child 1: its first leaf opens the right part
                const int r = k - after[dl[c1]].y;
                sepk[r] = a_res[u].y;
                if (weighted) { const double2 x = d_res[u]; sephi[r] = x.x; seplo[r] = x.y; }
            }
        }
        team_sync<CSIZE>();
        phase_stamp(3, CSIZE == 0 && tid == 0);
    }
}

// The same tables for BATCHES of mid-size trees (2k + 1 < 65536 nodes that fit shared memory), one CTA per tree with ALL
// the jumping state in shared memory: node ids and topological depths as uint16, the leaf list as (next | count << 16).
// tables_kernel<1> spends its time in dependent gathers from L2 (three per radix-4 round, about a dozen rounds);
// here a round is shared-memory latency.  Writes exactly what the later kernels read: jDI / jDH / jDL by node,
// pos, leaf_at, sepk / sephi / seplo by position.
__host__ __device__ inline size_t tables_smem_bytes(int64_t k, bool weighted) {
    const size_t N = 2 * (size_t)k + 1, NP = (N + 7) & ~(size_t)7;
    return 8 * NP * sizeof(uint16_t) + (weighted ? 2 * N * sizeof(double2) : 0);
}

template <bool WEIGHTED>
__global__ void __launch_bounds__(TAB_THREADS)
tables_smem_kernel(unsigned char* __restrict__ tab, size_t tab_stride, const double* __restrict__ bl, int bl_stride,
                   int64_t bl_tree_stride, int64_t B, int k, int unrooted, const int32_t* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char tsm[];
    const CophLayout L = coph_layout(k);
    const int n = L.n, N = L.N, root = 2 * k;
    const int NP = (N + 7) & ~7;
    const int tid = threadIdx.x, TT = blockDim.x;
    constexpr unsigned NONE = 0xFFFFu;
    uint16_t* const a16 = reinterpret_cast<uint16_t*>(tsm);      // [2][NP] ancestor
    uint16_t* const dep = a16 + 2 * NP;                          // [2][NP] depth below that ancestor
    uint16_t* const dl0 = dep + 2 * NP;                          // [2][NP] leftmost leaf
    uint16_t* const dr0 = dl0 + 2 * NP;                          // [2][NP] rightmost leaf
    double2* const dd2 = reinterpret_cast<double2*>(dr0 + 2 * NP);   // [2][N] double-double depth          [WEIGHTED]
    uint32_t* const ln = reinterpret_cast<uint32_t*>(tsm);       // [2][n] leaf list, over a16 / dep once they are read
    for (int64_t tree = blockIdx.x; tree < B; tree += gridDim.x) {
        if (status && status[tree] != 0) continue;               // uniform per CTA
        unsigned char* T = tab + (size_t)tree * tab_stride;
        const int* anc = reinterpret_cast<const int*>(T + L.anc);
        int* jDI = reinterpret_cast<int*>(T + L.jDI);
        double* jDH = reinterpret_cast<double*>(T + L.jDH);
        double* jDL = reinterpret_cast<double*>(T + L.jDL);
        int* pos = reinterpret_cast<int*>(T + L.pos);
        int* leaf_at = reinterpret_cast<int*>(T + L.leaf_at);
        int* sepk = reinterpret_cast<int*>(T + L.sepk);
        double* sephi = reinterpret_cast<double*>(T + L.sephi);
        double* seplo = reinterpret_cast<double*>(T + L.seplo);
        const double* blt = WEIGHTED ? bl + tree * bl_tree_stride : nullptr;
        if (tid < TAB_FLAGS) reinterpret_cast<int*>(T + L.flags)[tid] = 0;     // as tables_kernel leaves them
        __syncthreads();                                         // the previous tree's readers are done
        for (int r = tid; r < k; r += TT) {
            const int c1 = anc[3 * r], c2 = anc[3 * r + 1], p = k + 1 + r;
            int w1 = 1, w2 = 1;
            if (unrooted) {                                      // graph.rs:46-50: the root edge to node 2k-1 has length 0
                if (c1 == 2 * k - 1) w1 = 0;
                if (c2 == 2 * k - 1) w2 = 0;
            }
            P2V_CHECK_INDEX(c1, N); P2V_CHECK_INDEX(c2, N); P2V_CHECK_INDEX(p, N);
            a16[c1] = (uint16_t)p; a16[c2] = (uint16_t)p;
            dep[c1] = (uint16_t)w1; dep[c2] = (uint16_t)w2;
            if (WEIGHTED) {
                dd2[c1] = make_double2(w1 ? blt[(size_t)r * bl_stride] : 0.0, 0.0);
                dd2[c2] = make_double2(w2 ? blt[(size_t)r * bl_stride + 1] : 0.0, 0.0);
            }
            dl0[p] = (uint16_t)c1; dr0[p] = (uint16_t)c2;
        }
        for (int c = tid; c <= k; c += TT) { dl0[c] = (uint16_t)c; dr0[c] = (uint16_t)c; }
        if (tid == 0) { a16[root] = (uint16_t)NONE; dep[root] = 0; if (WEIGHTED) dd2[root] = make_double2(0.0, 0.0); }
        __syncthreads();
        // root-path sums and leftmost / rightmost leaves by radix-4 pointer jumping
        int cur = 0;
        for (;;) {
            const uint16_t* a_in = a16 + cur * NP; uint16_t* a_out = a16 + (cur ^ 1) * NP;
            const uint16_t* p_in = dep + cur * NP; uint16_t* p_out = dep + (cur ^ 1) * NP;
            const uint16_t* l_in = dl0 + cur * NP; uint16_t* l_out = dl0 + (cur ^ 1) * NP;
            const uint16_t* r_in = dr0 + cur * NP; uint16_t* r_out = dr0 + (cur ^ 1) * NP;
            const double2* d_in = dd2 + cur * N; double2* d_out = dd2 + (cur ^ 1) * N;
            int live = 0;
            for (int u = tid; u < N; u += TT) {
                unsigned nl = l_in[u], nr = r_in[u];             // leaves are fixed points
                unsigned a = a_in[u], dsum = p_in[u];
                dd d = {0.0, 0.0};
                if (WEIGHTED) { const double2 x = d_in[u]; d = dd{x.x, x.y}; }
#pragma unroll
                for (int hop = 0; hop < 3; ++hop) {
                    P2V_CHECK_INDEX(nl, N); P2V_CHECK_INDEX(nr, N);
                    nl = l_in[nl]; nr = r_in[nr];
                    if (a != NONE) {
                        P2V_CHECK_INDEX(a, N);
                        if (WEIGHTED) { const double2 x = d_in[a]; d = dd_add(d, dd{x.x, x.y}); }
                        dsum += p_in[a];
                        a = a_in[a];
                    }
                }
                live |= ((int)nl > k) | ((int)nr > k) | (a != NONE);
                l_out[u] = (uint16_t)nl; r_out[u] = (uint16_t)nr;
                a_out[u] = (uint16_t)a; p_out[u] = (uint16_t)dsum;
                if (WEIGHTED) d_out[u] = make_double2(d.hi, d.lo);
            }
            cur ^= 1;
            if (!__syncthreads_or(live)) break;
        }
        const uint16_t* dl = dl0 + cur * NP;
        const uint16_t* dr = dr0 + cur * NP;
        {   // depth of every node, where the other kernels read it
            const uint16_t* p_res = dep + cur * NP;
            const double2* d_res = dd2 + cur * N;
            for (int u = tid; u < N; u += TT) {
                jDI[u] = (int)p_res[u];
                if (WEIGHTED) { const double2 x = d_res[u]; jDH[u] = x.x; jDL[u] = x.y; }
            }
        }
        __syncthreads();                                         // a16 / dep are read: the leaf list takes their place
        // the leaves as a linked list in DFS order (see tables_kernel), ranked by pointer jumping
        for (int r = tid; r < k; r += TT) {
            const unsigned a = dr[anc[3 * r]], b = dl[anc[3 * r + 1]];
            P2V_CHECK_INDEX(a, n); P2V_CHECK_INDEX(b, n);
            ln[a] = b | (1u << 16);
        }
        if (tid == 0) ln[dr[root]] = NONE;                       // the last leaf: no successor, nothing after it
        __syncthreads();
        int lc = 0;
        for (;;) {
            const uint32_t* in = ln + lc * n; uint32_t* out = ln + (lc ^ 1) * n;
            int live = 0;
            for (int x = tid; x < n; x += TT) {
                uint32_t s = in[x];
                unsigned nx = s & 0xFFFFu, cnt = s >> 16;
#pragma unroll
                for (int hop = 0; hop < 3; ++hop) {
                    if (nx != NONE) { const uint32_t t = in[nx]; nx = t & 0xFFFFu; cnt += t >> 16; }
                }
                live |= (nx != NONE);
                out[x] = nx | (cnt << 16);
            }
            lc ^= 1;
            if (!__syncthreads_or(live)) break;
        }
        const uint32_t* after = ln + lc * n;                     // >> 16 = number of leaves after this one
        for (int c = tid; c <= k; c += TT) {
            const int p = k - (int)(after[c] >> 16);
            P2V_CHECK_INDEX(p, n);
            pos[c] = p;
            leaf_at[p] = c;
        }
        if (tid == 0) {
            sepk[0] = KEY_INF; sepk[n] = KEY_INF;
            if (WEIGHTED) { sephi[0] = 0.0; seplo[0] = 0.0; sephi[n] = 0.0; seplo[n] = 0.0; }
        }
        for (int u = k + 1 + tid; u <= 2 * k; u += TT) {          // jDI / jDH / jDL: this CTA's own stores, two barriers ago
            const int c1 = anc[3 * (u - k - 1) + 1];             // child 1: its first leaf opens the right part
            const int r = k - (int)(after[dl[c1]] >> 16);
            sepk[r] = jDI[u];
            if (WEIGHTED) { sephi[r] = jDH[u]; seplo[r] = jDL[u]; }
        }
    }
}

// ---- 3. blocks for the requested rows ----------------------------------------------------------
// Block id of DFS position p:  p / pcap + max(0, c(p) - 1) / rcap  (coph_call),  c(p) = number of requested
// rows at positions <= p.  Hence every block holds <= rcap requested rows and <= pcap positions
// (ids that no position maps to stay empty).  hdr[0] = number of ids, hdr[1] = number of tiles
// (blocks that hold requested rows), listed in tilelist.
__device__ __forceinline__ unsigned long long pack_key(int key, int p) {
    return ((unsigned long long)(unsigned)key << 32) | (unsigned)p;
}

// prefix / suffix minima of sep inside block `id` of one tree: one warp, 32 positions per step, the
// next step's keys loaded while the current one is scanned.
__device__ __forceinline__ void block_minima_one(unsigned char* T, const CophLayout& L, int id, int lane) {
    const int* sepk = reinterpret_cast<const int*>(T + L.sepk);
    int* prek = reinterpret_cast<int*>(T + L.prek);
    int* prea = reinterpret_cast<int*>(T + L.prea);
    int* sufk = reinterpret_cast<int*>(T + L.sufk);
    int* sufa = reinterpret_cast<int*>(T + L.sufa);
    int* blkk = reinterpret_cast<int*>(T + L.blkk);
    int* blka = reinterpret_cast<int*>(T + L.blka);
    const int lo = reinterpret_cast<const int*>(T + L.bstart)[id];
    const int hi = reinterpret_cast<const int*>(T + L.bend)[id];
    if (lo < 0) return;
    unsigned long long carry = pack_key(KEY_INF, 0);
    int knext = (lo + lane < hi) ? sepk[lo + lane] : KEY_INF;
    for (int base = lo; base < hi; base += 32) {       // pre[p] = min sep[lo..p]
        const int p = base + lane;
        const int kcur = knext;
        knext = (p + 32 < hi) ? sepk[p + 32] : KEY_INF;
        unsigned long long m = (p < hi) ? pack_key(kcur, p) : pack_key(KEY_INF, 0);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long y = __shfl_up_sync(FULL, m, d);
            if (lane >= d) m = min(m, y);
        }
        m = min(m, carry);
        if (p < hi) { prek[p] = (int)(m >> 32); prea[p] = (int)(m & 0xFFFFFFFFu); }
        carry = __shfl_sync(FULL, m, 31);
    }
    if (lane == 0) { blkk[id] = (int)(carry >> 32); blka[id] = (int)(carry & 0xFFFFFFFFu); }
    carry = pack_key(KEY_INF, 0);
    const int last = lo + ((hi - lo - 1) / 32) * 32;
    knext = (last + lane < hi) ? sepk[last + lane] : KEY_INF;
    for (int base = last; base >= lo; base -= 32) {   // suf[p] = min sep[p+1..hi-1]
        const int p = base + lane;
        const int kcur = knext;
        knext = (base - 32 >= lo) ? sepk[p - 32] : KEY_INF;
        unsigned long long m = (p < hi) ? pack_key(kcur, p) : pack_key(KEY_INF, 0);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long y = __shfl_down_sync(FULL, m, d);
            if (lane + d < 32) m = min(m, y);
        }
        unsigned long long ex = __shfl_down_sync(FULL, m, 1);
        if (lane == 31) ex = pack_key(KEY_INF, 0);
        ex = min(ex, carry);
        if (p < hi) { sufk[p] = (int)(ex >> 32); sufa[p] = (int)(ex & 0xFFFFFFFFu); }
        carry = min(carry, __shfl_sync(FULL, m, 0));
    }
}

// Per-label copies of what rows_kernel needs about column j, so that its column setup is one level
// of coalesced loads: block of the leaf, and the two in-block candidates (prefix / suffix minimum
// at the leaf's position) with their double-double depths.
__device__ __forceinline__ void column_tables_one(unsigned char* T, const CophLayout& L, int j, int weighted) {
    const int p = reinterpret_cast<const int*>(T + L.pos)[j];
    const int pk = reinterpret_cast<const int*>(T + L.prek)[p], sk = reinterpret_cast<const int*>(T + L.sufk)[p];
    reinterpret_cast<int*>(T + L.c_blk)[j] = reinterpret_cast<const int*>(T + L.blk_of)[p];
    reinterpret_cast<int*>(T + L.c_prek)[j] = pk;
    reinterpret_cast<int*>(T + L.c_sufk)[j] = sk;
    if (weighted) {
        const double* sephi = reinterpret_cast<const double*>(T + L.sephi);
        const double* seplo = reinterpret_cast<const double*>(T + L.seplo);
        const int pa = reinterpret_cast<const int*>(T + L.prea)[p], sa = reinterpret_cast<const int*>(T + L.sufa)[p];
        const bool hp = pk < KEY_INF, hs = sk < KEY_INF;
        reinterpret_cast<double*>(T + L.c_prehi)[j] = hp ? sephi[pa] : 0.0;
        reinterpret_cast<double*>(T + L.c_prelo)[j] = hp ? seplo[pa] : 0.0;
        reinterpret_cast<double*>(T + L.c_sufhi)[j] = hs ? sephi[sa] : 0.0;
        reinterpret_cast<double*>(T + L.c_suflo)[j] = hs ? seplo[sa] : 0.0;
    }
}

// CSIZE > 1: a thread-block cluster per tree (few, large trees): every CTA takes a contiguous part
// of the positions, the per-warp counts of requested rows go through global scratch.
// CSIZE == 0 (the whole grid on one tree) also runs the block minima and the per-label column tables
// behind two more grid barriers instead of two more launches.
template <int CSIZE>
__global__ void __launch_bounds__(BLK_THREADS)
blocks_kernel(unsigned char* __restrict__ tab, size_t tab_stride, int64_t B, int k, int64_t row_begin,
              int64_t row_end, int pcap, int rcap, const int32_t* __restrict__ status, int weighted) {
    __shared__ int s_part[BLK_THREADS / 32];
    __shared__ int s_total;
    const CophLayout L = coph_layout(k);
    const int tid = threadIdx.x, lane = tid & 31;
    const int crank = team_rank<CSIZE>();
    constexpr int NWC = BLK_THREADS / 32;          // warps per CTA
    const int NW = team_ctas<CSIZE>() * NWC;       // warps per team
    const int warp = crank * NWC + (tid >> 5);     // warp index inside the team
    const int n = L.n;
    const int64_t team = team_index<CSIZE>(), nteams = team_count<CSIZE>();
    for (int64_t tree = team; tree < B; tree += nteams) {
        if (status && status[tree] != 0) continue;
        unsigned char* T = tab + (size_t)tree * tab_stride;
        const int* leaf_at = reinterpret_cast<const int*>(T + L.leaf_at);
        int* blk_of = reinterpret_cast<int*>(T + L.blk_of);
        int* rowlist = reinterpret_cast<int*>(T + L.rowlist);
        int* bstart = reinterpret_cast<int*>(T + L.bstart);
        int* bend = reinterpret_cast<int*>(T + L.bend);
        int* rowoff = reinterpret_cast<int*>(T + L.rowoff);
        int* cend = reinterpret_cast<int*>(T + L.cend);
        int* blkk = reinterpret_cast<int*>(T + L.blkk);
        int* blka = reinterpret_cast<int*>(T + L.blka);
        int* hdr = reinterpret_cast<int*>(T + L.hdr);

        int* wcount = reinterpret_cast<int*>(T + L.tilelist);   // per-warp counts (tilelist is filled last)
        team_sync<CSIZE>();   // shared scratch of the previous tree is no longer read
        phase_stamp(32, CSIZE == 0 && crank == 0 && tid == 0);
        for (int i = crank * BLK_THREADS + tid; i < L.nbmax; i += team_ctas<CSIZE>() * BLK_THREADS) {
            bstart[i] = -1; bend[i] = -1; rowoff[i] = 0; cend[i] = 0; blkk[i] = KEY_INF; blka[i] = -1;
        }
        // every warp owns a contiguous range of positions and reads it with coalesced loads
        const int per_warp = (((n + NW - 1) / NW) + 31) & ~31;
        const int w0 = min(n, warp * per_warp), w1 = min(n, w0 + per_warp);
        int cnt = 0;
#pragma unroll 8
        for (int base = w0; base < w1; base += 32) {     // independent loads: several in flight
            const int p = base + lane;
            const int lab = (p < w1) ? leaf_at[p] : -1;
            cnt += __popc(__ballot_sync(FULL, lab >= row_begin && lab < row_end));
        }
        int c0;                                      // requested rows at positions < w0
        if (CSIZE == 1) {
            if (lane == 0) s_part[warp] = cnt;
            __syncthreads();
            if (warp == 0) {
                const int v = s_part[lane];
                const int sc = (int)warp_incl_scan((uint32_t)v, lane);
                s_part[lane] = sc - v;
            }
            __syncthreads();
            c0 = s_part[warp];
        } else {
            if (lane == 0) wcount[warp] = cnt;
            team_sync<CSIZE>();
            int acc = 0;
            for (int w = lane; w < warp; w += 32) acc += wcount[w];
#pragma unroll
            for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(FULL, acc, d);
            c0 = acc;
        }
        phase_stamp(33, CSIZE == 0 && crank == 0 && tid == 0);
        int prev_last = (w0 > 0) ? (w0 - 1) / pcap + max(0, c0 - 1) / rcap : -1;   // id of position w0-1
        int lab_next = (w0 + lane < w1) ? leaf_at[w0 + lane] : -1;
        int lab_next2 = (w0 + 32 + lane < w1) ? leaf_at[w0 + 32 + lane] : -1;
        for (int base = w0; base < w1; base += 32) {
            const int p = base + lane;
            const bool in = p < w1;
            const int lab = lab_next;                     // loaded two iterations ahead
            lab_next = lab_next2;
            lab_next2 = (p + 64 < w1) ? leaf_at[p + 64] : -1;
            const bool own = lab >= row_begin && lab < row_end;
            const unsigned bm = __ballot_sync(FULL, own);
            const int c = c0 + __popc(bm & lanes_le(lane));          // requested rows at positions <= p
            const int id = in ? p / pcap + max(0, c - 1) / rcap : -1;
            int prev_id = __shfl_up_sync(FULL, id, 1);
            if (lane == 0) prev_id = prev_last;
            if (in) {
                blk_of[p] = id;
                if (own) rowlist[c - 1] = p;
                if (id != prev_id) {
                    bstart[id] = p; rowoff[id] = c - (own ? 1 : 0);
                    if (prev_id >= 0) { bend[prev_id] = p; cend[prev_id] = c - (own ? 1 : 0); }
                }
                if (p == n - 1) { bend[id] = n; cend[id] = c; hdr[0] = id + 1; }
            }
            c0 += __popc(bm);
            const int last = __shfl_sync(FULL, id, 31);
            if (last >= 0) prev_last = last;
        }
        team_sync<CSIZE>();
        phase_stamp(34, CSIZE == 0 && crank == 0 && tid == 0);
        // compact the blocks that hold requested rows into tilelist
        if (crank == 0) {
            const int warp = tid >> 5;
            int* tilelist = reinterpret_cast<int*>(T + L.tilelist);
            const int nid = hdr[0];
            int base_count = 0;
            for (int id0 = 0; id0 < nid; id0 += BLK_THREADS) {
                const int id = id0 + tid;
                const int has = (id < nid && bstart[id] >= 0 && cend[id] - rowoff[id] > 0) ? 1 : 0;
                const int inc = (int)warp_incl_scan((uint32_t)has, lane);
                __syncthreads();
                if (lane == 31) s_part[warp] = inc;
                __syncthreads();
                if (warp == 0) {
                    const int v = s_part[lane];
                    const int sc = (int)warp_incl_scan((uint32_t)v, lane);
                    s_part[lane] = sc - v;
                    if (lane == 31) s_total = sc;
                }
                __syncthreads();
                if (has) tilelist[base_count + s_part[warp] + inc - 1] = id;
                base_count += s_total;
            }
            if (tid == 0) hdr[1] = base_count;
        }
        if (CSIZE == 0) {
            team_sync<CSIZE>();                       // hdr, bstart / bend of every block are final
            phase_stamp(35, crank == 0 && tid == 0);
            const int nid = hdr[0];
            {   // blocks dealt round-robin over the CTAs first: every SM takes a few, none takes 32
                const int nct = team_ctas<CSIZE>();
                // Blocks are dealt round-robin over the CTAs (every SM takes a few), and never to warp 0:
                // measured on B200, the warp of the CTA's master thread runs shuffle-heavy code about ten
                // times slower after a grid barrier (63 us instead of 6 for a 512-position block;
                // profiles/r02_c4_notes.md), whatever reconverges it.
                const int wq = tid >> 5;
                if (wq > 0)
                    for (int id = (wq - 1) * nct + crank; id < nid; id += (NWC - 1) * nct) block_minima_one(T, L, id, lane);
            }
            if (CSIZE == 0) grid_barrier(reinterpret_cast<unsigned*>(T + L.flags) + 100);
            else team_sync<CSIZE>();
            phase_stamp(36, crank == 0 && tid == 0);
            for (int j = crank * BLK_THREADS + tid; j < n; j += team_ctas<CSIZE>() * BLK_THREADS)
                column_tables_one(T, L, j, weighted);
            phase_stamp(37, crank == 0 && tid == 0);
        }
    }
}

template <int CSIZE>
static bool launch_blocks_cluster(unsigned char* tab, size_t tab_stride, int64_t B, int k, int64_t row_begin,
                                  int64_t row_end, int pcap, int rcap, const int32_t* st, cudaStream_t stream) {
    auto kern = blocks_kernel<CSIZE>;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * CSIZE));
    cfg.blockDim = dim3(BLK_THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CSIZE; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess || max_clusters < 1) {
        cudaGetLastError();
        return false;
    }
    if (cudaLaunchKernelEx(&cfg, kern, tab, tab_stride, B, k, row_begin, row_end, pcap, rcap, st, 0) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return true;
}

// one warp per (tree, block), whole grid
__global__ void __launch_bounds__(256)
block_minima_kernel(unsigned char* __restrict__ tab, size_t tab_stride, int64_t B, int k,
                    const int32_t* __restrict__ status) {
    const CophLayout L = coph_layout(k);
    const int lane = threadIdx.x & 31;
    const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t w = gwarp; w < B * L.nbmax; w += nwarps) {
        const int64_t tree = w / L.nbmax;
        const int id = (int)(w % L.nbmax);
        if (status && status[tree] != 0) continue;
        unsigned char* T = tab + (size_t)tree * tab_stride;
        if (id >= reinterpret_cast<const int*>(T + L.hdr)[0]) continue;
        block_minima_one(T, L, id, lane);
    }
}

__global__ void __launch_bounds__(256)
column_tables_kernel(unsigned char* __restrict__ tab, size_t tab_stride, int64_t B, int k, int weighted,
                     const int32_t* __restrict__ status) {
    const CophLayout L = coph_layout(k);
    const int n = L.n;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < B * n; w += (int64_t)gridDim.x * blockDim.x) {
        const int64_t tree = w / n;
        const int j = (int)(w % n);
        if (status && status[tree] != 0) continue;
        column_tables_one(tab + (size_t)tree * tab_stride, L, j, weighted);
    }
}

// ---- 4. rows --------------------------------------------------------------------------------
__device__ __forceinline__ double int_to_f64(int v) {
    // exact for |v| < 2^31 : one DADD instead of an I2F.F64 on the slow pipe
    return __hiloint2double(0x43300000, (int)(0x80000000u ^ (unsigned)v)) - 4503601774854144.0;  // 2^52 + 2^31
}

template <bool VEC2>
__device__ __forceinline__ void store2(double* p, double a, double b) {
    if (VEC2) {
        double2 v; v.x = a; v.y = b;
        __stcs(reinterpret_cast<double2*>(p), v);
    } else {
        __stcs(p, a); __stcs(p + 1, b);
    }
}

// double-double x - y for x >= y >= 0 on one root path, rounded to one double
__device__ __forceinline__ double dd_sub_round(double xh, double xl, double yh, double yl) {
    const double s = __dadd_rn(xh, -yh);                     // Fast2Sum: xh >= yh >= 0
    const double e = __dadd_rn(__dadd_rn(xh, -s), -yh);
    return __dadd_rn(s, __dadd_rn(e, __dadd_rn(xl, -yl)));
}

// double-double a - b, rounded to one double
__device__ __forceinline__ double dd_diff_round(dd a, dd b) {
    return __dadd_rn(__dadd_rn(a.hi, -b.hi), __dadd_rn(a.lo, -b.lo));
}

struct RowShared {
    int label[ROW_R];          // leaf label of the row
    int qpos[ROW_R];           // its DFS position
    int di[ROW_R];
    double dhi[ROW_R], dlo[ROW_R];
    unsigned long long wtotR[8], wtotL[8]; // per-warp totals of the block scan
    long long roff[ROW_R];                 // (label - row_begin) * n: element offset of the row in the output
    // what the streaming loop reads per column, indexed [row][side] ([0]: columns left of the tile, [1]:
    // right) so that a column's side is an address offset instead of three 64-bit selects:
    // sv = (double-double depth of the row candidate hi, lo; d_i - d_rl rounded once),
    // si = (topological depth of the row candidate; d_i - that depth)
    double sv[ROW_R][2][3];
    int si[ROW_R][2][2];
    int incount;
};
// dynamic shared memory of rows_kernel: midhi[nidb], midlo[nidb] (doubles), midk[nidb], mida[nidb],
// ptab[rcap * nincap], sephi[bsz], seplo[bsz], inhi[nincap], inlo[nincap] (doubles), then kblk[bsz],
// inlist[bsz], inlab[bsz], indi[bsz] (ints), st[levels][bsz] (uint16)
inline size_t rows_dyn_smem(const CophCall& c) {
    const int levels = ilog2c(c.bsz) + 1;
    return (((size_t)24 * c.nidb + 15) & ~(size_t)15) + (size_t)8 * c.rcap * c.nincap + (size_t)20 * c.bsz +
           (size_t)28 * c.nincap +
           (size_t)2 * levels * c.bsz + 16;
}

__device__ __forceinline__ unsigned long long shfl_up64(unsigned long long x, int d) {
    const unsigned lo = __shfl_up_sync(FULL, (unsigned)x, d), hi = __shfl_up_sync(FULL, (unsigned)(x >> 32), d);
    return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ unsigned long long shfl_down64(unsigned long long x, int d) {
    const unsigned lo = __shfl_down_sync(FULL, (unsigned)x, d), hi = __shfl_down_sync(FULL, (unsigned)(x >> 32), d);
    return ((unsigned long long)hi << 32) | lo;
}

constexpr int rows_min_blocks(int threads) { return threads >= 256 ? 3 : (threads == 128 ? 6 : (threads == 64 ? 10 : 16)); }

// VCV: write depth(LCA) instead of the distance (vector/graph.rs:211-258: out[l][r] = root_to_x of
// the node where l and r meet, diagonal = root_to_x[leaf]) -- the same candidates, min instead of
// the three-term sum.
template <bool TOPO, bool VEC2, int THREADS, bool VCV>
__global__ void __launch_bounds__(THREADS, rows_min_blocks(THREADS))
rows_kernel(const unsigned char* __restrict__ tab, size_t tab_stride, int64_t B, int k,
            int64_t row_begin, int64_t row_end, double* __restrict__ out,
            const int32_t* __restrict__ status, int chunks_per_tree, int chunks_per_cta, int tiles_bound,
            int pmax /* block size cap */, int nidb, int rcap, int nincap) {
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ RowShared S;
    constexpr int CHUNK = THREADS * ROW_CPT;
    constexpr int NWARP = THREADS / 32;
    const CophLayout L = coph_layout(k);
    const int n = L.n;
    double* midhi = reinterpret_cast<double*>(dyn);   // nidb: dd depth of the minimum between the tile and block J
    double* midlo = midhi + nidb;
    int* midk = reinterpret_cast<int*>(midlo + nidb);    // nidb: its key
    int* mida = midk + nidb;                          // nidb: its position
    double* s_ptab = reinterpret_cast<double*>(dyn + (((size_t)24 * nidb + 15) & ~(size_t)15));   // rcap*nincap: row x in-block column
    double* s_sephi = s_ptab + (size_t)rcap * nincap; // pmax: dd depth of the block's separators
    double* s_seplo = s_sephi + pmax;
    double* s_inhi = s_seplo + pmax;                  // nincap: dd depth of the staged in-block leaves of this chunk
    double* s_inlo = s_inhi + nincap;
    int* s_kblk = reinterpret_cast<int*>(s_inlo + nincap);  // pmax: separator keys of the block
    int* s_inlist = s_kblk + pmax;                    // nincap: staged in-block leaves of this chunk (positions)
    int* s_indi = s_inlist + nincap;                  // nincap: their topological depths
    unsigned short* s_st = reinterpret_cast<unsigned short*>(s_indi + nincap);   // [levels][pmax] argmin offsets
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t rows = row_end - row_begin;

    const int groups = (chunks_per_tree + chunks_per_cta - 1) / chunks_per_cta;   // chunk groups per tile
    const int64_t tiles_per_tree = (int64_t)tiles_bound * groups;
    for (int64_t work = blockIdx.x; work < B * tiles_per_tree; work += gridDim.x) {
        const int64_t tree = work / tiles_per_tree;
        const int rem = (int)(work % tiles_per_tree);
        const int tix = rem / groups;
        const int chunk_begin = (rem % groups) * chunks_per_cta;
        const int chunk_end = min(chunks_per_tree, chunk_begin + chunks_per_cta);
        if (status && status[tree] != 0) continue;
        const unsigned char* T = tab + (size_t)tree * tab_stride;
        const int* hdr = reinterpret_cast<const int*>(T + L.hdr);
        if (tix >= hdr[1]) continue;                  // uniform per CTA
        const int nb = hdr[0];
        const int I = reinterpret_cast<const int*>(T + L.tilelist)[tix];
        const int lo = reinterpret_cast<const int*>(T + L.bstart)[I];
        const int hi = reinterpret_cast<const int*>(T + L.bend)[I];
        const int roff = reinterpret_cast<const int*>(T + L.rowoff)[I];
        const int nrow = reinterpret_cast<const int*>(T + L.cend)[I] - roff;
        const int* jDI = reinterpret_cast<const int*>(T + L.jDI);
        const double* jDH = reinterpret_cast<const double*>(T + L.jDH);
        const double* jDL = reinterpret_cast<const double*>(T + L.jDL);
        const int* pos = reinterpret_cast<const int*>(T + L.pos);
        const int* leaf_at = reinterpret_cast<const int*>(T + L.leaf_at);
        const int* sepk = reinterpret_cast<const int*>(T + L.sepk);
        const double* sephi = reinterpret_cast<const double*>(T + L.sephi);
        const double* seplo = reinterpret_cast<const double*>(T + L.seplo);
        const int* blk_of = reinterpret_cast<const int*>(T + L.blk_of);
        const int* rowlist = reinterpret_cast<const int*>(T + L.rowlist);
        const int* prek = reinterpret_cast<const int*>(T + L.prek);
        const int* prea = reinterpret_cast<const int*>(T + L.prea);
        const int* sufk = reinterpret_cast<const int*>(T + L.sufk);
        const int* sufa = reinterpret_cast<const int*>(T + L.sufa);
        const int* blkk = reinterpret_cast<const int*>(T + L.blkk);
        const int* blka = reinterpret_cast<const int*>(T + L.blka);
        const int* c_blk = reinterpret_cast<const int*>(T + L.c_blk);
        const int* c_prek = reinterpret_cast<const int*>(T + L.c_prek);
        const int* c_sufk = reinterpret_cast<const int*>(T + L.c_sufk);
        const double* c_prehi = reinterpret_cast<const double*>(T + L.c_prehi);
        const double* c_prelo = reinterpret_cast<const double*>(T + L.c_prelo);
        const double* c_sufhi = reinterpret_cast<const double*>(T + L.c_sufhi);
        const double* c_suflo = reinterpret_cast<const double*>(T + L.c_suflo);
        __syncthreads();   // previous iteration's readers of S / dyn are done

        // ---- row side of the tile -------------------------------------------------------
        for (int r = tid; r < nrow; r += THREADS) {
            const int q = rowlist[roff + r];
            const int i = leaf_at[q];
            S.label[r] = i; S.qpos[r] = q;
            S.roff[r] = (long long)(i - row_begin) * n;
            const int di = jDI[i];
            S.di[r] = di;
            const dd d = {jDH[i], jDL[i]};
            if (!TOPO) { S.dhi[r] = d.hi; S.dlo[r] = d.lo; }
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                const int ck = side ? sufk[q] : prek[q];
                const int ca = side ? sufa[q] : prea[q];
                const bool none = ck >= KEY_INF;         // no separator on that side inside the block
                if (TOPO) {
                    const int a = none ? di : ck;        // key == topological depth of the candidate
                    S.si[r][side][0] = a; S.si[r][side][1] = di - a;
                } else {
                    const dd a = none ? d : dd{sephi[ca], seplo[ca]};
                    const double ufv = none ? 0.0 : dd_diff_round(d, a);
                    S.sv[r][side][0] = a.hi; S.sv[r][side][1] = a.lo; S.sv[r][side][2] = ufv;
                }
            }
        }
        // separators of the block for the in-block patch
        for (int x = tid; x < hi - lo; x += THREADS) {
            s_kblk[x] = sepk[lo + x];
            s_st[x] = (unsigned short)x;
            if (!TOPO) { s_sephi[x] = sephi[lo + x]; s_seplo[x] = seplo[lo + x]; }
        }

        // ---- minima over the whole blocks between the tile and every other block --------
        // midk[J] / mida[J] = min / arg position of sep over the blocks strictly between I and J.
        // Every thread owns a contiguous segment of blocks; the segment minima go through a
        // block-wide exclusive prefix-min (blocks right of I) and suffix-min (left of I).
        {
            const int SEG = (nb + THREADS - 1) / THREADS;
            const int s0 = min(nb, tid * SEG), s1 = min(nb, s0 + SEG);
            const unsigned long long INF64 = ((unsigned long long)KEY_INF << 32);
            unsigned long long pr = INF64, pl = INF64;      // this segment's minimum right / left of I
            for (int J = s0; J < s1; ++J) {
                const unsigned long long e = ((unsigned long long)(unsigned)blkk[J] << 32) | (unsigned)blka[J];
                if (J > I) pr = min(pr, e);
                if (J < I) pl = min(pl, e);
            }
            unsigned long long ir = pr, il = pl;            // inclusive scans inside the warp
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long y = shfl_up64(ir, d), z = shfl_down64(il, d);
                if (lane >= d) ir = min(ir, y);
                if (lane + d < 32) il = min(il, z);
            }
            unsigned long long runR = shfl_up64(ir, 1), runL = shfl_down64(il, 1);
            if (lane == 0) runR = INF64;
            if (lane == 31) runL = INF64;
            if (NWARP > 1) {
                if (lane == 31) S.wtotR[warp] = ir;
                if (lane == 0) S.wtotL[warp] = il;
                __syncthreads();
                for (int w = 0; w < NWARP; ++w) {
                    if (w < warp) runR = min(runR, S.wtotR[w]);
                    if (w > warp) runL = min(runL, S.wtotL[w]);
                }
            }
            for (int J = s0; J < s1; ++J) {                  // blocks right of I, walking right
                if (J > I) {
                    midk[J] = (int)(runR >> 32); mida[J] = (int)(runR & 0xFFFFFFFFu);
                    if (!TOPO && runR < INF64) { midhi[J] = sephi[mida[J]]; midlo[J] = seplo[mida[J]]; }
                    runR = min(runR, ((unsigned long long)(unsigned)blkk[J] << 32) | (unsigned)blka[J]);
                }
            }
            for (int J = s1 - 1; J >= s0; --J) {             // blocks left of I, walking left
                if (J < I) {
                    midk[J] = (int)(runL >> 32); mida[J] = (int)(runL & 0xFFFFFFFFu);
                    if (!TOPO && runL < INF64) { midhi[J] = sephi[mida[J]]; midlo[J] = seplo[mida[J]]; }
                    runL = min(runL, ((unsigned long long)(unsigned)blkk[J] << 32) | (unsigned)blka[J]);
                }
                if (J == I) { midk[J] = KEY_INF; mida[J] = 0; }
            }
        }
        __syncthreads();

        // sparse table of argmin offsets over the block's separators (for the in-block patch)
        const int bs = hi - lo;
        for (int lev = 1; (1 << lev) <= bs; ++lev) {
            const int half = 1 << (lev - 1);
            for (int x = tid; x + (1 << lev) <= bs; x += THREADS) {
                const int a = s_st[(lev - 1) * pmax + x], b = s_st[(lev - 1) * pmax + x + half];
                s_st[lev * pmax + x] = (unsigned short)(s_kblk[a] <= s_kblk[b] ? a : b);
            }
            __syncthreads();
        }
        double* obase = out + (size_t)tree * rows * n;

        for (int chunk = chunk_begin; chunk < chunk_end; ++chunk) {
        const int cbase = chunk * CHUNK;
        if (tid == 0) S.incount = 0;
        __syncthreads();
        // ---- column side: 4 columns per thread (two adjacent pairs) ----------------------
        int col[ROW_CPT], side[ROW_CPT];
        unsigned inpack = 0;               // per column: slot + 1 in the in-block list, 0 = outside (8 bits each)
        unsigned inmask = 0;               // per column: 1 = in-block (whether staged or not)
        int bi[ROW_CPT], uji[ROW_CPT];
        double bhi[ROW_CPT], blo[ROW_CPT], ujf[ROW_CPT];
#pragma unroll
        for (int c = 0; c < ROW_CPT; ++c) {
            const int j = cbase + (c >> 1) * (THREADS * 2) + 2 * tid + (c & 1);
            col[c] = j;
            side[c] = 0; bi[c] = 0; uji[c] = 0; bhi[c] = blo[c] = ujf[c] = 0.0;
            if (j < n) {
                const int J = c_blk[j];
                const bool right = J > I;
                side[c] = right ? 1 : 0;
                int ck = right ? c_prek[j] : c_sufk[j];
                const int mk = midk[J];
                const bool usemid = mk < ck;
                if (usemid) ck = mk;
                const bool none = ck >= KEY_INF;
                int slot = -1;
                if (J == I) {                      // its LCA with the rows lies inside the block
                    slot = atomicAdd(&S.incount, 1);
                    inmask |= 1u << c;
                    if (slot < nincap) { s_inlist[slot] = pos[j]; inpack |= (unsigned)(slot + 1) << (8 * c); }
                    else slot = -1;
                }
                if (TOPO) {
                    const int dj = jDI[j];
                    if (slot >= 0) s_indi[slot] = dj;
                    bi[c] = none ? dj : ck;
                    uji[c] = dj - bi[c];
                } else {
                    const dd d = {jDH[j], jDL[j]};
                    if (slot >= 0) { s_inhi[slot] = d.hi; s_inlo[slot] = d.lo; }
                    dd b = right ? dd{c_prehi[j], c_prelo[j]} : dd{c_sufhi[j], c_suflo[j]};
                    if (usemid) b = dd{midhi[J], midlo[J]};
                    if (none) b = d;
                    bhi[c] = b.hi; blo[c] = b.lo;
                    ujf[c] = none ? 0.0 : dd_diff_round(d, b);
                }
            }
        }

        // ---- in-block columns (incl. the zero diagonal): their LCA lies inside the block and comes
        //      from the sparse table.  Normally (<= ROW_NINCAP of them in this chunk) the values are
        //      staged in shared memory and merged into the streamed stores; otherwise they are
        //      written by a patch pass after the streamed part.
        __syncthreads();
        const int nin = S.incount;
        const bool staged = nin <= nincap;
        auto inblock_value = [&](int r, int p, int dj_i, dd dj) -> double {
            const int q = S.qpos[r];
            if (p == q) return VCV ? (TOPO ? int_to_f64(S.di[r]) : S.dhi[r]) : 0.0;
            const int a = min(p, q) + 1 - lo, b = max(p, q) - lo;     // sep offsets a..b
            const int lev = 31 - __clz(b - a + 1);
            const int x = s_st[lev * pmax + a], y = s_st[lev * pmax + b - (1 << lev) + 1];
            const int arg = (s_kblk[x] <= s_kblk[y]) ? x : y;
            if (VCV) return TOPO ? int_to_f64(s_kblk[arg]) : s_sephi[arg];
            if (TOPO) return int_to_f64(S.di[r] + dj_i - 2 * s_kblk[arg]);
            // (d_i - d_lca) + (d_j - d_lca): each difference of double-doubles rounded once
            const double lh = s_sephi[arg], ll = s_seplo[arg];
            return __dadd_rn(dd_sub_round(S.dhi[r], S.dlo[r], lh, ll), dd_sub_round(dj.hi, dj.lo, lh, ll));
        };
        if (staged) {
            for (int w = tid; w < nin * nrow; w += THREADS) {
                const int r = w / nin, e = w - r * nin;
                s_ptab[r * nincap + e] = TOPO ? inblock_value(r, s_inlist[e], s_indi[e], dd{0.0, 0.0})
                                              : inblock_value(r, s_inlist[e], 0, dd{s_inhi[e], s_inlo[e]});
            }
            __syncthreads();
        }
        const bool warp_has_in = staged && __any_sync(FULL, inpack != 0);
        // a thread with ONE staged in-block column (the usual case) merges it with one shared-memory load
        // and four predicated moves per row; a warp where some thread has several takes the general form
        int in_c = -1, in_e = 0;
#pragma unroll
        for (int c = 0; c < ROW_CPT; ++c) {
            const unsigned e = (inpack >> (8 * c)) & 0xFFu;
            if (e) { in_c = (in_c < 0) ? c : 4; in_e = (int)e - 1; }
        }
        const bool warp_multi = warp_has_in && __any_sync(FULL, in_c == 4);
        const bool pin0 = in_c == 0, pin1 = in_c == 1, pin2 = in_c == 2, pin3 = in_c == 3;
        // store targets and predicates of this thread's two column pairs (VEC2: n is even, a pair is in or out)
        double* const cp0 = obase + col[0];
        double* const cp1 = obase + col[2];
        const bool full0 = col[0] + 1 < n, full1 = col[2] + 1 < n;
        const bool half0 = !full0 && col[0] < n, half1 = !full1 && col[2] < n;
        const int sx0 = side[0], sx1 = side[1], sx2 = side[2], sx3 = side[3];
        const int sxs[ROW_CPT] = {sx0, sx1, sx2, sx3};

        // ---- streamed part: requested rows x 4 columns per thread -----------------------------
#pragma unroll 2
        for (int r = 0; r < nrow; ++r) {
            const long long ro = S.roff[r];
            double val[ROW_CPT];
            if (TOPO) {
#pragma unroll
                for (int c = 0; c < ROW_CPT; ++c) {
                    const int2 au = *reinterpret_cast<const int2*>(&S.si[r][sxs[c]][0]);
                    val[c] = VCV ? int_to_f64(min(au.x, bi[c])) : int_to_f64(au.y + uji[c] + abs(au.x - bi[c]));
                }
            } else {
#pragma unroll
                for (int c = 0; c < ROW_CPT; ++c) {
                    const double* sv = &S.sv[r][sxs[c]][0];
                    const double ah = sv[0], al = sv[1];
                    if (VCV) {      // the shallower candidate; both lie on one root path
                        const bool rowmin = ah < bhi[c] || (ah == bhi[c] && al < blo[c]);
                        val[c] = rowmin ? ah : bhi[c];
                    } else {
                        const double u = sv[2];
                        const double m = __dadd_rn(__dadd_rn(ah, -bhi[c]), __dadd_rn(al, -blo[c]));
                        val[c] = __dadd_rn(__dadd_rn(u, ujf[c]), fabs(m));
                    }
                }
            }
            if (warp_has_in) {           // warp-uniform
                if (!warp_multi) {
                    const double x = s_ptab[r * nincap + in_e];
                    if (pin0) val[0] = x;
                    if (pin1) val[1] = x;
                    if (pin2) val[2] = x;
                    if (pin3) val[3] = x;
                } else {
#pragma unroll
                    for (int c = 0; c < ROW_CPT; ++c) {
                        const unsigned e = (inpack >> (8 * c)) & 0xFFu;
                        if (e) val[c] = s_ptab[r * nincap + e - 1];
                    }
                }
            }
            if (full0) store2<VEC2>(cp0 + ro, val[0], val[1]);
            else if (half0) __stcs(cp0 + ro, val[0]);
            if (full1) store2<VEC2>(cp1 + ro, val[2], val[3]);
            else if (half1) __stcs(cp1 + ro, val[2]);
        }

        if (!staged) {   // rare: more in-block columns in this chunk than the staging table holds
            __syncthreads();   // orders the streamed stores of this CTA before the patch stores
#pragma unroll
            for (int c = 0; c < ROW_CPT; ++c) {
                if (!((inmask >> c) & 1u)) continue;        // every thread patches its own columns
                const int j = col[c], p = pos[j];
                const int dj_i = TOPO ? jDI[j] : 0;
                const dd dj = TOPO ? dd{0.0, 0.0} : dd{jDH[j], jDL[j]};
                for (int r = 0; r < nrow; ++r)
                    obase[(size_t)(S.label[r] - row_begin) * n + j] = inblock_value(r, p, dj_i, dj);
            }
        }
        __syncthreads();   // the patch has read inlist / incount before the next chunk resets them
        }  // chunk
    }
}

// ---- small trees: one CTA per tree, decode + tables + matrix in one kernel --------------------
// For n <= 512 the general path above (five kernels with per-tree tables in global memory, CTAs of
// 1024 threads) spends most of its time in per-tree fixed costs.  Here a CTA of 128 threads keeps
// everything about one tree in shared memory:
//   1. warp 0 decodes v to the ancestry (the warp-per-tree decoder of p2v_decode_warp.cuh);
//   2. subtree leaf counts (rows are in child-before-parent order), then depths (int and
//      double-double) and DFS offsets of every node by double-buffered pointer jumping;
//   3. separators by DFS position, keyed (depth << 16 | position), and a sparse table over them;
//   4. each warp takes rows i (label order): the LCA of (i, leaf at position p) is one O(1)
//      range-minimum query, the row is built in DFS order in shared memory and leaves permuted to
//      label order as coalesced 128-bit streaming stores.
// Same arithmetic as the row kernel's in-block values: distances are (d_i - d_lca) + (d_j - d_lca),
// each difference of double-doubles rounded once.
constexpr int CS_N_MAX = 512;            // capacity of the fused kernel
constexpr int CS_MAXLC = 5;              // label-order column pairs per lane: 64 * 5 = 320 columns
constexpr int CS_N_DEFAULT = 320;        // used up to here (above, the general kernels are faster)
constexpr int CS_THREADS = 128;
constexpr unsigned CS_NONE = 0xFFFFu;

constexpr int CS_TREES = CS_THREADS / 32;       // trees decoded side by side (one per warp) per CTA pass

struct SmallLayout {
    size_t dec, anc, ja, jd, dl, dr, nxt, cnt, pos, leaf_at, dipos, sepd, ddpos, jh, jl, st, rowbuf, colk, total;
    size_t dec_stride, anc_stride;
    int levels, tpp;                    // tpp: trees decoded side by side per pass (one per warp)
    bool wt;                            // tiny trees: every warp takes its tree through ALL phases (own copy of the tables)
    size_t wt_stride;                   // bytes between the warps' table copies
};
__host__ __device__ inline SmallLayout small_layout(int k, bool topo, bool wt) {
    SmallLayout L;
    const size_t n = (size_t)k + 1, N = 2 * (size_t)k + 1, K = ((size_t)k + 31) & ~(size_t)31;
    L.levels = 1;                                       // floor(log2(n - 1)) + 1
    while (((size_t)2 << (L.levels - 1)) <= (n > 1 ? n - 1 : 1)) ++L.levels;
    size_t o = 0;
    L.dec_stride = al16(2 * ((size_t)3 * (K + 32) + 64));
    L.anc_stride = al16(4 * 3 * (size_t)k);
    L.tpp = (n <= 128) ? CS_TREES : 1;  // bigger trees: one at a time, fewer bytes per CTA
    L.dec = o; o += L.tpp * L.dec_stride;
    L.anc = o; o += L.tpp * L.anc_stride;
    L.ja = o; o = al16(o + 2 * 2 * N);
    L.jd = o; o = al16(o + 2 * 2 * N);
    L.dl = o; o = al16(o + 2 * 2 * N);
    L.dr = o; o = al16(o + 2 * 2 * N);
    L.nxt = o; o = al16(o + 2 * 2 * n);
    L.cnt = o; o = al16(o + 2 * 2 * n);
    L.pos = o; o = al16(o + 2 * (n + 1));
    L.leaf_at = o; o = al16(o + 2 * n);
    const size_t P32 = (n + 31) & ~(size_t)31;          // positions padded to whole chunks of 32
    L.dipos = o; o = al16(o + 2 * P32);
    L.sepd = o; o = al16(o + (topo ? 0 : 16 * (n + 1)));
    L.ddpos = o; o = al16(o + (topo ? 0 : 16 * P32));
    // sparse table: level 0 (the separator keys) is written while the jumping buffers of the
    // double-double depths are still being read; the upper levels and the row buffers are only
    // written after that and live on top of those buffers
    const size_t stn = (n + 1 + 3) & ~(size_t)3;        // entries per level, 16-byte multiple
    L.st = o; o = o + 4 * stn;
    const size_t a0 = o;
    L.jh = o; o = al16(o + (topo ? 0 : 8 * 2 * N));
    L.jl = o; o = al16(o + (topo ? 0 : 8 * 2 * N));
    const size_t a1 = o;
    o = a0 + 4 * stn * (size_t)(L.levels - 1);
    o = al16(o);
    L.rowbuf = o; o = al16(o + (size_t)(CS_THREADS / 32) * 8 * P32);
    L.colk = o; o = al16(o + (size_t)(CS_THREADS / 32) * 4 * P32);     // column candidates of a warp's tile
    L.total = al16(o > a1 ? o : a1);
    L.wt = wt;
    L.wt_stride = L.total - L.ja;
    if (L.wt) L.total = L.ja + (size_t)L.tpp * L.wt_stride;
    return L;
}


// TILED: rows by tiles of 32 DFS positions with per-lane label-order columns in registers (pays off
// from n ~ 150); otherwise one sparse-table query per element and a DFS-ordered row buffer.
template <bool TOPO, bool VCV, bool VEC2, bool TILED, bool STAGED, bool WT>
__global__ void __launch_bounds__(CS_THREADS, 8)
coph_small_kernel(const int* __restrict__ v32, const double* __restrict__ bl, int bl_stride, int64_t bl_tree_stride,
                  int64_t B, int k, int unrooted, int64_t row_begin, int64_t row_end, double* __restrict__ out,
                  int32_t* __restrict__ status, int use_tma) {
    constexpr bool stage_lengths = STAGED && !TOPO;
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ int s_bad[CS_TREES];
    __shared__ __align__(8) uint64_t s_bar[2];    // [1]: mbarrier of the staged branch lengths
    __shared__ uint8_t s_lut[2048];
    nth_bit_lut_init(s_lut, threadIdx.x, CS_THREADS);
    const SmallLayout L = small_layout(k, TOPO, WT);
    const int n = k + 1, N = 2 * k + 1, K = (k + 31) & ~31, root = 2 * k;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = CS_THREADS / 32;
    // tiny trees (n <= 32, L.wt): a 31-node tree keeps one warp busy at best, so each of the CTA's warps takes its
    // own tree through tables and rows, on its own copy of the tables, with warp-level barriers only
    unsigned char* const tb = dyn + (WT ? (size_t)warp * L.wt_stride : 0);
    const int t0 = WT ? lane : tid;                                       // the thread's index / count in its team
    constexpr int ts = WT ? 32 : CS_THREADS;
    auto bar = [&]() { if (WT) __syncwarp(); else __syncthreads(); };
    auto bar_or = [&](int x) -> int { return WT ? __any_sync(FULL, x) : __syncthreads_or(x); };
    uint16_t* ja = reinterpret_cast<uint16_t*>(tb + L.ja);
    uint16_t* jd = reinterpret_cast<uint16_t*>(tb + L.jd);
    uint16_t* dl = reinterpret_cast<uint16_t*>(tb + L.dl);
    uint16_t* dr = reinterpret_cast<uint16_t*>(tb + L.dr);
    uint16_t* nxt = reinterpret_cast<uint16_t*>(tb + L.nxt);
    uint16_t* cnt = reinterpret_cast<uint16_t*>(tb + L.cnt);
    uint16_t* pos = reinterpret_cast<uint16_t*>(tb + L.pos);
    uint16_t* leaf_at = reinterpret_cast<uint16_t*>(tb + L.leaf_at);
    uint16_t* dipos = reinterpret_cast<uint16_t*>(tb + L.dipos);
    double2* sepd = reinterpret_cast<double2*>(tb + L.sepd);
    double2* ddpos = reinterpret_cast<double2*>(tb + L.ddpos);
    double* jh = reinterpret_cast<double*>(tb + L.jh);
    double* jl = reinterpret_cast<double*>(tb + L.jl);
    uint32_t* st = reinterpret_cast<uint32_t*>(tb + L.st);
    const int rb_stride = (n + 31) & ~31;
    double* rowbuf = reinterpret_cast<double*>(tb + L.rowbuf) + (size_t)warp * rb_stride;
    uint32_t* colk = reinterpret_cast<uint32_t*>(tb + L.colk) + (size_t)warp * rb_stride;
    const int64_t rows = row_end - row_begin;
    const int stn = (n + 1 + 3) & ~3;             // entries per sparse-table level

    const int tpp = L.tpp;
    // ---- staged branch lengths (TMA): while the rows of a tree are being written, the lengths of the NEXT tree
    //      (its matrix rows, 24k bytes) arrive by one bulk copy (cp.async.bulk, completion on an mbarrier).
    //      They land in the jumping buffers ja .. cnt, which are dead during the row phase (40k + 24 bytes >=
    //      24k + 32), so staging costs no shared memory.  A copy takes the whole 16-byte pieces that cover the
    //      source range; where those would reach outside the caller's array (an unaligned first or last
    //      tree) the threads copy the range themselves.  Measured (profiles/r02_coph_small_tma.md): staging the
    //      v rows the same way (4k bytes per tree, into the tree's ancestry slot) made the kernel 1 - 5 %
    //      slower, so v is read straight from global memory; the staged lengths cost 2 - 4 % (the other CTAs
    //      of the SM already hide these loads) and are a template flag, off for n <= 128.
    const int bl_first = (bl_stride == 3) ? 1 : 0;        // matrix rows [v, bl1, bl2]: a tree's block starts one double earlier
    const double* const bl_lo = TOPO ? nullptr : bl - bl_first;
    double* const sbl_stage = reinterpret_cast<double*>(dyn + L.ja);
    struct Piece { const char* a0; uint32_t bytes, lead; bool tma; };
    auto piece = [use_tma](const char* g, size_t len, const char* lo, const char* hi) {
        Piece q;
        q.a0 = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(g) & ~(uintptr_t)15);
        const char* e = reinterpret_cast<const char*>((reinterpret_cast<uintptr_t>(g) + len + 15) & ~(uintptr_t)15);
        q.bytes = (uint32_t)(e - q.a0); q.lead = (uint32_t)(g - q.a0);
        q.tma = use_tma && q.a0 >= lo && e <= hi;
        return q;
    };
    auto bl_piece = [&](int64_t t) {
        return piece(reinterpret_cast<const char*>(bl_lo + t * bl_tree_stride), (size_t)bl_tree_stride * 8,
                     reinterpret_cast<const char*>(bl_lo), reinterpret_cast<const char*>(bl_lo + B * bl_tree_stride));
    };
    // every thread calls this at a point where the destination is dead
    auto stage_bl = [&](int64_t t) {
        if (TOPO || !stage_lengths || t >= B) return;
        const Piece q = bl_piece(t);
        if (q.tma) { if (tid == 0) { fence_proxy_async_smem(); bulk_load(sbl_stage, q.a0, q.bytes, &s_bar[1]); } }
        else for (int i = tid; i < (int)bl_tree_stride; i += CS_THREADS) sbl_stage[q.lead / 8 + i] = __ldg(bl_lo + t * bl_tree_stride + i);
    };
    uint32_t par_bl = 0;
    if (stage_lengths) {
        if (tid == 0) { mbar_init(&s_bar[1], 1); fence_proxy_async_smem(); }
        __syncthreads();
        stage_bl((int64_t)blockIdx.x * tpp);
    }
    for (int64_t base = (int64_t)blockIdx.x * tpp; base < B; base += (int64_t)gridDim.x * tpp) {
        const int64_t base_next = base + (int64_t)gridDim.x * tpp;
        __syncthreads();                          // the previous pass's readers are done
        // ---- 1. decode: every warp its own tree ---------------------------------------------
        if (warp < tpp && base + warp < B) {
            const int64_t tree = base + warp;
            const int bad = decode_tree_warp<int, int, MODE_ANCESTRY, 1>(
                v32 + tree * k, k, K, reinterpret_cast<int*>(dyn + L.anc + warp * L.anc_stride),
                reinterpret_cast<uint16_t*>(dyn + L.dec + warp * L.dec_stride), lane, s_lut);
            if (lane == 0) { s_bad[warp] = bad; status[tree] = bad + 1; }
        }
        __syncthreads();
        // tiny trees (WT): every warp takes its own tree through all the phases, with warp-level barriers only
        for (int sub = WT ? warp : 0; sub < (WT ? warp + 1 : tpp) && base + sub < B; ++sub) {
        const int64_t tree = base + sub;
        const int64_t tree_next = (sub + 1 < tpp && tree + 1 < B) ? tree + 1 : base_next;
        Piece qb; qb.lead = 0; qb.tma = false;
        if (!TOPO && stage_lengths) {
            qb = bl_piece(tree);
            if (qb.tma) { mbar_wait(&s_bar[1], par_bl); par_bl ^= 1u; }
        }
        if (s_bad[sub] >= 0) {                    // uniform: skip the tree, keep the staging going
            if (!TOPO && stage_lengths) { bar(); stage_bl(tree_next); bar(); }
            continue;
        }
        const int* anc = reinterpret_cast<const int*>(dyn + L.anc + sub * L.anc_stride);
        // ---- 2. parents, edge weights, first / last child pointers -----------------------------
        if (!TOPO && stage_lengths) {             // the staged lengths first: they sit where ja .. cnt are about to go
            const double* blt = sbl_stage + qb.lead / 8 + bl_first;
            for (int r = t0; r < k; r += ts) {
                const int c1 = anc[3 * r], c2 = anc[3 * r + 1];
                double b1 = blt[(size_t)r * bl_stride], b2 = blt[(size_t)r * bl_stride + 1];
                if (unrooted) {                   // graph.rs:46-50: the root edge to node 2k-1 has length 0
                    if (c1 == 2 * k - 1) b1 = 0.0;
                    if (c2 == 2 * k - 1) b2 = 0.0;
                }
                jh[c1] = b1; jh[c2] = b2; jl[c1] = 0.0; jl[c2] = 0.0;
            }
            if (t0 == 0) { jh[root] = 0.0; jl[root] = 0.0; }
            bar();
        }
        for (int c = t0; c <= k; c += ts) { dl[c] = (uint16_t)c; dr[c] = (uint16_t)c; }
        const bool direct_bl = !TOPO && !stage_lengths;                 // lengths straight from global memory
        const double* blg = direct_bl ? bl + tree * bl_tree_stride : nullptr;
        for (int r = t0; r < k; r += ts) {
            const int c1 = anc[3 * r], c2 = anc[3 * r + 1], p = k + 1 + r;
            double b1 = 1.0, b2 = 1.0;
            if (direct_bl) { b1 = blg[(size_t)r * bl_stride]; b2 = blg[(size_t)r * bl_stride + 1]; }
            int w1 = 1, w2 = 1;
            if (unrooted) {
                if (c1 == 2 * k - 1) { w1 = 0; b1 = 0.0; }
                if (c2 == 2 * k - 1) { w2 = 0; b2 = 0.0; }
            }
            P2V_CHECK_INDEX(c1, N); P2V_CHECK_INDEX(c2, N); P2V_CHECK_INDEX(p, N);
            ja[c1] = (uint16_t)p; ja[c2] = (uint16_t)p;
            jd[c1] = (uint16_t)w1; jd[c2] = (uint16_t)w2;
            dl[p] = (uint16_t)c1; dr[p] = (uint16_t)c2;
            if (direct_bl) { jh[c1] = b1; jh[c2] = b2; jl[c1] = 0.0; jl[c2] = 0.0; }
        }
        if (t0 == 0) {
            ja[root] = (uint16_t)CS_NONE; jd[root] = 0;
            if (!TOPO && !stage_lengths) { jh[root] = 0.0; jl[root] = 0.0; }
        }
        bar();
        // ---- pointer jumping (buffers 0 / 1): root-path sums upwards, leftmost / rightmost leaf
        //      downwards (leaves are fixed points) ---------------------------------------------------
        int cur = 0;
        for (;;) {
            const uint16_t* a_in = ja + cur * N; uint16_t* a_out = ja + (cur ^ 1) * N;
            const uint16_t* d_in = jd + cur * N; uint16_t* d_out = jd + (cur ^ 1) * N;
            const uint16_t* l_in = dl + cur * N; uint16_t* l_out = dl + (cur ^ 1) * N;
            const uint16_t* r_in = dr + cur * N; uint16_t* r_out = dr + (cur ^ 1) * N;
            const double* h_in = jh + cur * N; double* h_out = jh + (cur ^ 1) * N;
            const double* w_in = jl + cur * N; double* w_out = jl + (cur ^ 1) * N;
            int live = 0;
            for (int x = t0; x < N; x += ts) {
                const unsigned a = a_in[x];
                unsigned d = d_in[x], na = CS_NONE;
                dd f = {0.0, 0.0};
                if (!TOPO) f = dd{h_in[x], w_in[x]};
                if (a != CS_NONE) {
                    d += d_in[a];
                    if (!TOPO) f = dd_add(f, dd{h_in[a], w_in[a]});
                    na = a_in[a];
                    live |= (na != CS_NONE);
                }
                const unsigned nl = l_in[l_in[x]], nr = r_in[r_in[x]];
                live |= ((int)nl > k) | ((int)nr > k);
                a_out[x] = (uint16_t)na; d_out[x] = (uint16_t)d;
                l_out[x] = (uint16_t)nl; r_out[x] = (uint16_t)nr;
                if (!TOPO) { h_out[x] = f.hi; w_out[x] = f.lo; }
            }
            cur ^= 1;
            if (!bar_or(live)) break;
        }
        const uint16_t* dres = jd + cur * N;
        const uint16_t* lres = dl + cur * N;
        const uint16_t* rres = dr + cur * N;
        // ---- leaves as a linked list in DFS order: every internal node links the last leaf of its
        //      first child's subtree to the first leaf of its second child's; ranking by jumping -----
        for (int r = t0; r < k; r += ts) {
            const int a = rres[anc[3 * r]], b = lres[anc[3 * r + 1]];
            P2V_CHECK_INDEX(a, n); P2V_CHECK_INDEX(b, n);
            nxt[a] = (uint16_t)b; cnt[a] = 1;
        }
        if (t0 == 0) { const int e = rres[root]; nxt[e] = (uint16_t)CS_NONE; cnt[e] = 0; }
        bar();
        int lc = 0;
        for (;;) {
            const uint16_t* n_in = nxt + lc * n; uint16_t* n_out = nxt + (lc ^ 1) * n;
            const uint16_t* c_in = cnt + lc * n; uint16_t* c_out = cnt + (lc ^ 1) * n;
            int live = 0;
            for (int x = t0; x < n; x += ts) {
                const unsigned a = n_in[x];
                unsigned c = c_in[x], na = CS_NONE;
                if (a != CS_NONE) { c += c_in[a]; na = n_in[a]; live |= (na != CS_NONE); }
                n_out[x] = (uint16_t)na; c_out[x] = (uint16_t)c;
            }
            lc ^= 1;
            if (!bar_or(live)) break;
        }
        // ---- 3. positions, separators, sparse table -------------------------------------------
        {
            const uint16_t* cres = cnt + lc * n;          // leaves after this one
            const double* hres = jh + cur * N;
            const double* wres = jl + cur * N;
            for (int c = t0; c <= k; c += ts) {
                const int p = k - cres[c];
                P2V_CHECK_INDEX(p, n);
                pos[c] = (uint16_t)p; leaf_at[p] = (uint16_t)c; dipos[p] = dres[c];
                if (!TOPO) ddpos[p] = make_double2(hres[c], wres[c]);
            }
            for (int r = t0; r < k; r += ts) {
                const int u = k + 1 + r;
                const int rp = k - cres[lres[anc[3 * r + 1]]];     // first position of the second child
                P2V_CHECK_INDEX(rp, n + 1);
                st[rp] = ((uint32_t)dres[u] << 16) | (uint32_t)rp;
                if (!TOPO) sepd[rp] = make_double2(hres[u], wres[u]);
            }
            if (t0 == 0) { st[0] = 0xFFFFFFFFu; st[n] = 0xFFFFFFFFu; }
        }
        bar();
        stage_bl(tree_next);                      // the jumping buffers are dead from here on: the next lengths land in them
        for (int lev = 1; lev < L.levels; ++lev) {
            const int half = 1 << (lev - 1);
            for (int r = 1 + t0; r + (1 << lev) <= n; r += ts)
                st[lev * stn + r] = min(st[(lev - 1) * stn + r], st[(lev - 1) * stn + r + half]);
            bar();
        }
        if (!TILED) {
        // ---- 4. rows: every lane owns pairs of columns in LABEL order, (j, j + 1) with j = 64 lc + 2 lane, keeps
        //      their DFS positions (and, for n <= 128, their double-double depths) in registers for all the rows
        //      of its warp, answers one sparse-table query per element and stores whole 16-byte pairs.  (Round 1
        //      built each row in DFS order in shared memory and permuted it on the way out: 25 trips through
        //      the shared-memory pipe per 32 elements -- the pipe ran at 87 % -- against 11 now.) -----------
        auto rows_direct = [&](auto lc_tag, auto cache_tag) {
            constexpr int LC = decltype(lc_tag)::value;
            constexpr bool CACHE = decltype(cache_tag)::value;
            uint32_t cpos[LC], cdi[LC];
            double2 cd0[CACHE ? LC : 1], cd1[CACHE ? LC : 1];
#pragma unroll
            for (int lc = 0; lc < LC; ++lc) {
                const int j = 64 * lc + 2 * lane;
                cpos[lc] = 0; cdi[lc] = 0;
                if (CACHE) { cd0[lc] = make_double2(0.0, 0.0); cd1[lc] = cd0[lc]; }
                if (j < n) {
                    const uint32_t pp = *reinterpret_cast<const uint32_t*>(pos + j);   // pos is padded to even n
                    const uint32_t p0 = pp & 0xFFFFu, p1 = (j + 1 < n) ? (pp >> 16) : p0;
                    cpos[lc] = p0 | (p1 << 16);
                    cdi[lc] = (uint32_t)dipos[p0] | ((uint32_t)dipos[p1] << 16);
                    if (!TOPO && CACHE) { cd0[lc] = ddpos[p0]; cd1[lc] = ddpos[p1]; }
                }
            }
            double* orow = out + ((size_t)tree * rows + (WT ? 0 : warp)) * n;
            const int i_end = (int)row_end;
            for (int i = (int)row_begin + (WT ? 0 : warp); i < i_end; i += (WT ? 1 : NW), orow += (size_t)(WT ? 1 : NW) * n) {
                const int q = pos[i];
                const int di = dipos[q];
                double2 dq = make_double2(0.0, 0.0);
                if (!TOPO) dq = ddpos[q];
                auto element = [&](int p, int dj_i, double2 dj) -> double {
                    if (p == q) return VCV ? (TOPO ? int_to_f64(di) : dq.x) : 0.0;
                    const int a = min(p, q) + 1, b = max(p, q);
                    const int lev = 31 - __clz(b - a + 1);
                    P2V_CHECK_INDEX(lev, L.levels); P2V_CHECK_INDEX(a, stn); P2V_CHECK_INDEX(b - (1 << lev) + 1, stn);
                    const uint32_t m = min(st[lev * stn + a], st[lev * stn + b - (1 << lev) + 1]);
                    if (TOPO) {
                        const int dlca = (int)(m >> 16);
                        return VCV ? int_to_f64(dlca) : int_to_f64(di + dj_i - 2 * dlca);
                    }
                    P2V_CHECK_INDEX(m & 0xFFFFu, n + 1);
                    const double2 sd = sepd[m & 0xFFFFu];
                    if (VCV) return sd.x;
                    return __dadd_rn(dd_sub_round(dq.x, dq.y, sd.x, sd.y), dd_sub_round(dj.x, dj.y, sd.x, sd.y));
                };
#pragma unroll
                for (int lc = 0; lc < LC; ++lc) {
                    const int j = 64 * lc + 2 * lane;
                    if (j < n) {
                        const int p0 = (int)(cpos[lc] & 0xFFFFu), p1 = (int)(cpos[lc] >> 16);
                        double2 d0 = make_double2(0.0, 0.0), d1 = d0;
                        if (!TOPO && !VCV) {
                            if (CACHE) { d0 = cd0[lc]; d1 = cd1[lc]; }
                            else { d0 = ddpos[p0]; d1 = ddpos[p1]; }
                        }
                        const double v0 = element(p0, (int)(cdi[lc] & 0xFFFFu), d0);
                        const double v1 = element(p1, (int)(cdi[lc] >> 16), d1);
                        if (j + 1 < n) store2<VEC2>(orow + j, v0, v1);
                        else __stcs(orow + j, v0);
                    }
                }
            }
        };
        if (n > 128) rows_direct(std::integral_constant<int, CS_MAXLC>{}, std::false_type{});
        else if (n > 40) rows_direct(std::integral_constant<int, 2>{}, std::true_type{});
        else {
        // tiny trees (a pair of columns per lane leaves most lanes idle): the round-1 form, each row built in DFS
        // order in shared memory, one position per lane, and permuted on the way out
        double* orow = out + ((size_t)tree * rows + (WT ? 0 : warp)) * n;
        const int i_end = (int)row_end;
        for (int i = (int)row_begin + (WT ? 0 : warp); i < i_end; i += (WT ? 1 : NW), orow += (size_t)(WT ? 1 : NW) * n) {
            const int q = pos[i];
            const int di = dipos[q];
            double2 dq = make_double2(0.0, 0.0);
            if (!TOPO) dq = ddpos[q];
            // positions past n - 1 (last chunk) compute a harmless value into the padding of rowbuf
#pragma unroll 1
            for (int c0 = 0; c0 < n; c0 += 32) {
                const int p = c0 + lane;
                const int pc = min(p, k);
                const int a = min(pc, q) + 1, b = max(pc, q);
                const int lev = 31 - __clz(max(b - a + 1, 1));
                const uint32_t m = min(st[lev * stn + a], st[lev * stn + b - (1 << lev) + 1]);
                double val;
                if (TOPO) {
                    const int dlca = (int)(m >> 16);
                    val = VCV ? int_to_f64(dlca) : int_to_f64(di + (int)dipos[p] - 2 * dlca);
                    if (p == q) val = VCV ? int_to_f64(di) : 0.0;
                } else {
                    const double2 sd = sepd[m & 0xFFFFu];
                    if (VCV) val = sd.x;
                    else {
                        const double2 dj = ddpos[p];
                        val = __dadd_rn(dd_sub_round(dq.x, dq.y, sd.x, sd.y), dd_sub_round(dj.x, dj.y, sd.x, sd.y));
                    }
                    if (p == q) val = VCV ? dq.x : 0.0;
                }
                rowbuf[p] = val;
            }
            __syncwarp();
            if (VEC2) {                           // n even: whole pairs, positions read two at a time
#pragma unroll 1
                for (int j = 2 * lane; j < n; j += 64) {
                    const uint32_t pp = *reinterpret_cast<const uint32_t*>(pos + j);
                    double2 v2;
                    v2.x = rowbuf[pp & 0xFFFFu]; v2.y = rowbuf[pp >> 16];
                    __stcs(reinterpret_cast<double2*>(orow + j), v2);
                }
            } else {
                for (int j = lane; j < n; j += 32) __stcs(orow + j, rowbuf[pos[j]]);
            }
            __syncwarp();
        }
        }
        } else {
        // ---- 4. rows, a tile of 32 consecutive DFS positions per warp ---------------------------------
        // For a row at position q of tile t and a column at position p of another chunk c the range
        // of separators (min(p,q), max(p,q)] splits at the tile boundary: the row's part (inside the
        // tile, one value per row and side) and the column's part (from the tile boundary to p, one
        // value per column and tile, from chunk-wise min scans) -- the LCA key is the min of the two.
        // Columns inside the tile itself take one sparse-table query each.
        double* obase = out + (size_t)tree * rows * n;
        const int nch = (n + 31) >> 5;
        const uint32_t KINF = 0xFFFFFFFFu;
        // fewer tiles than warps (n <= 96): the rows of a tile are dealt round-robin to NS warps
        const int NS = (nch < NW) ? NW / nch : 1;
        const int rsub = (nch < NW) ? warp / nch : 0;
        for (int t = (nch < NW) ? warp % nch : warp; t < nch && rsub < NS; t += NW) {
            const int base = t << 5;
            // does the tile hold any requested row?  (labels of its positions)
            const int qmine = base + lane;
            const int lab_mine = (qmine < n) ? (int)leaf_at[qmine] : -1;
            const bool want = lab_mine >= (int)row_begin && lab_mine < (int)row_end;
            const unsigned wantmask = __ballot_sync(FULL, want);
            if (!wantmask) continue;              // warp-uniform
            // column parts: right of the tile ...
            uint32_t carry = KINF;
            for (int c = t + 1; c < nch; ++c) {
                const int p = (c << 5) + lane;
                uint32_t x = (p < n) ? st[p] : KINF;          // key[p]: separator between p - 1 and p
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(FULL, x, d);
                    if (lane >= d) x = min(x, y);
                }
                x = min(x, carry);
                colk[p] = x;
                carry = __shfl_sync(FULL, x, 31);
            }
            // ... and left of it (separators p + 1 .. base)
            carry = KINF;
            for (int c = t - 1; c >= 0; --c) {
                const int p = (c << 5) + lane;
                uint32_t x = st[p + 1];                        // p + 1 <= base <= n - 1
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_down_sync(FULL, x, d);
                    if (lane + d < 32) x = min(x, y);
                }
                x = min(x, carry);
                colk[p] = x;
                carry = __shfl_sync(FULL, x, 0);
            }
            // row parts of lane's own row: separators q + 1 .. base + 31 and base + 1 .. q
            uint32_t rowR, rowL;
            {
                const int qn = qmine + 1;
                uint32_t x = (lane < 31 && qn < n) ? st[qn] : KINF;      // key[q + 1], not past the tile
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_down_sync(FULL, x, d);
                    if (lane + d < 32) x = min(x, y);
                }
                rowR = x;
                uint32_t z = (lane > 0 && qmine < n) ? st[qmine] : KINF; // key[q], not before the tile
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(FULL, z, d);
                    if (lane >= d) z = min(z, y);
                }
                rowL = z;
            }
            __syncwarp();
            // per-lane columns in LABEL order: pairs (j, j + 1), j = 64 * lc + 2 * lane.  Their
            // positions, column candidates and depths stay in registers for all rows of the tile.
            uint32_t cpos[CS_MAXLC], ckey0[CS_MAXLC], ckey1[CS_MAXLC], cdi[CS_MAXLC];
#pragma unroll
            for (int lc = 0; lc < CS_MAXLC; ++lc) {
                const int j = 64 * lc + 2 * lane;
                cpos[lc] = 0; ckey0[lc] = KINF; ckey1[lc] = KINF; cdi[lc] = 0;
                if (j < n) {
                    const uint32_t pp = *reinterpret_cast<const uint32_t*>(pos + j);   // pos is padded to even n
                    const uint32_t p0 = pp & 0xFFFFu, p1 = (j + 1 < n) ? (pp >> 16) : p0;
                    cpos[lc] = p0 | (p1 << 16);
                    ckey0[lc] = colk[p0]; ckey1[lc] = colk[p1];          // (unused for in-tile columns)
                    cdi[lc] = (uint32_t)dipos[p0] | ((uint32_t)dipos[p1] << 16);
                }
            }
            double* tilebuf = rowbuf;             // 32 in-tile values of the current row
            for (int r = 0; r < 32; ++r) {
                if (!((wantmask >> r) & 1u) || (r % NS) != rsub) continue;        // warp-uniform
                const int q = base + r;
                const int i = __shfl_sync(FULL, lab_mine, r);
                const uint32_t rR = __shfl_sync(FULL, rowR, r), rL = __shfl_sync(FULL, rowL, r);
                const int di = dipos[q];
                double2 dq = make_double2(0.0, 0.0);
                if (!TOPO) dq = ddpos[q];
                auto value = [&](uint32_t m, uint32_t p, int dj_i) -> double {
                    if (TOPO) {
                        const int dlca = (int)(m >> 16);
                        return VCV ? int_to_f64(dlca) : int_to_f64(di + dj_i - 2 * dlca);
                    }
                    const double2 sd = sepd[min(m & 0xFFFFu, (uint32_t)n)];
                    if (VCV) return sd.x;
                    const double2 dj = ddpos[p];
                    return __dadd_rn(dd_sub_round(dq.x, dq.y, sd.x, sd.y), dd_sub_round(dj.x, dj.y, sd.x, sd.y));
                };
                {   // inside the tile: one sparse-table query per position
                    const int p = base + lane;
                    const int pc = min(p, k);
                    const int a = min(pc, q) + 1, b = max(pc, q);
                    const int lev = 31 - __clz(max(b - a + 1, 1));
                    const uint32_t m = min(st[lev * stn + a], st[lev * stn + b - (1 << lev) + 1]);
                    double val = value(m, (uint32_t)pc, (int)dipos[pc]);
                    if (p == q) val = VCV ? (TOPO ? int_to_f64(di) : dq.x) : 0.0;
                    tilebuf[lane] = val;
                }
                __syncwarp();
                double* orow = obase + (size_t)(i - (int)row_begin) * n;
#pragma unroll
                for (int lc = 0; lc < CS_MAXLC; ++lc) {
                    const int j = 64 * lc + 2 * lane;
                    if (j < n) {
                        const uint32_t p0 = cpos[lc] & 0xFFFFu, p1 = cpos[lc] >> 16;
                        const uint32_t o0 = p0 - (uint32_t)base, o1 = p1 - (uint32_t)base;   // < 32: in the tile
                        const uint32_t m0 = min(p0 > (uint32_t)base ? rR : rL, ckey0[lc]);
                        const uint32_t m1 = min(p1 > (uint32_t)base ? rR : rL, ckey1[lc]);
                        double v0 = value(m0, p0, (int)(cdi[lc] & 0xFFFFu));
                        double v1 = value(m1, p1, (int)(cdi[lc] >> 16));
                        if (o0 < 32u) v0 = tilebuf[o0];
                        if (o1 < 32u) v1 = tilebuf[o1];
                        if (j + 1 < n) store2<VEC2>(orow + j, v0, v1);
                        else __stcs(orow + j, v0);
                    }
                }
                __syncwarp();
            }
        }
        }
        bar();                          // rows done before the next tree's tables are built
        }  // sub
    }
}

__global__ void fill_f64_kernel(double* __restrict__ out, int64_t count, double value) {
    const int64_t n2 = count >> 1;
    double2 v; v.x = value; v.y = value;
    double2* o2 = reinterpret_cast<double2*>(out);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n2; i += 4 * stride) {     // four 128-bit streaming stores in flight per thread
        __stcs(o2 + i, v); __stcs(o2 + i + stride, v); __stcs(o2 + i + 2 * stride, v); __stcs(o2 + i + 3 * stride, v);
    }
    for (; i < n2; i += stride) __stcs(o2 + i, v);
    if ((count & 1) && blockIdx.x == 0 && threadIdx.x == 0) out[count - 1] = value;
}

// --------------------------------------------------------------------------------------------
// workspace: [ status scratch (B int32) | int32 v (B*k) | decode workspace | tables (B * total) ]
struct CophWs { size_t status, v32, dec, tab, total; };
static CophWs coph_ws(int64_t B, int64_t k) {
    CophWs w;
    size_t o = 0;
    w.status = o; o = al16(o + 4 * (size_t)B);
    w.v32 = o; o = al16(o + 4 * (size_t)B * (size_t)(k > 0 ? k : 1));
    w.dec = o; o = al16(o + decode_workspace_bytes(B, k));
    w.tab = o; o = al16(o + (size_t)B * (k >= 2 ? coph_layout(k).total : 16));
    w.total = o;
    return w;
}

size_t cophenetic_workspace_bytes(int64_t B, int64_t k) {
    if (B <= 0) return 0;
    return coph_ws(B, k).total;
}

template <int CSIZE>
static bool launch_tables_cluster(unsigned char* tab, size_t tab_stride, const double* bl, int bl_stride,
                                  int64_t bl_tree_stride, int64_t B, int k, int unrooted, const int32_t* st,
                                  cudaStream_t stream) {
    auto kern = tables_kernel<CSIZE>;
    if (CSIZE > 8 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * CSIZE));
    cfg.blockDim = dim3(TAB_THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CSIZE; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess || max_clusters < 1) {
        cudaGetLastError();
        return false;
    }
    if (cudaLaunchKernelEx(&cfg, kern, tab, tab_stride, bl, bl_stride, bl_tree_stride, B, k, unrooted, st) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return true;
}

// Phase 1+2: decode and build the per-tree tables inside the workspace.
cudaError_t launch_cophenetic_prepare(const void* v_or_matrix, int idx_bytes, const double* bls, int64_t B,
                                      int64_t k, int unrooted, int32_t* status, void* ws, size_t ws_bytes,
                                      cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    if (k > LARGE_K_MAX) return cudaErrorNotSupported;
    const CophWs W = coph_ws(B, k);
    if (!ws || ws_bytes < W.total) return cudaErrorInvalidValue;
    unsigned char* wsb = static_cast<unsigned char*>(ws);
    int32_t* st = reinterpret_cast<int32_t*>(wsb + W.status);
    int* v32 = reinterpret_cast<int*>(wsb + W.v32);
    const int grid1d = sm_count() * 8;
    cudaError_t e;

    // branch lengths: bls (B,k,2) or columns 1..2 of the (B,k,3) matrix
    const double* bl = bls;
    int bl_stride = 2;
    int64_t bl_tree_stride = 2 * k;
    const void* vin = v_or_matrix;
    int vin_bytes = idx_bytes;
    if (idx_bytes == 0) {  // matrix input
        const double* m = static_cast<const double*>(v_or_matrix);
        bl = m + 1; bl_stride = 3; bl_tree_stride = 3 * k;
        if (k > 0) {
            matrix_v_kernel<<<grid1d, 256, 0, stream>>>(m, B * k, v32);
            count_launch();
        }
        vin = v32; vin_bytes = 4;
    } else if (idx_bytes == 8) {
        if (k > 0) {
            narrow_v_kernel<<<grid1d, 256, 0, stream>>>(static_cast<const long long*>(v_or_matrix), B * k, v32);
            count_launch();
        }
        vin = v32; vin_bytes = 4;
    } else if (idx_bytes != 4) {
        return cudaErrorInvalidValue;
    }
    if (k < 2) {
        if (k == 1) e = launch_check_v(vin, vin_bytes, B, k, st, stream);
        else e = cudaMemsetAsync(st, 0, sizeof(int32_t) * B, stream);
        if (e != cudaSuccess) return e;
        // k == 1: the rows phase needs bl0 + bl1 per tree; keep it in the table area (16 B per tree)
        if (k == 1) {
            double* sums = reinterpret_cast<double*>(wsb + W.tab);
            tiny_sum_kernel<<<(int)((B + 255) / 256 < 1024 ? (B + 255) / 256 : 1024), 256, 0, stream>>>(
                bl, bl_tree_stride, B, sums);
            count_launch();
            if (bl) {
                flag_negative_bl_kernel<<<(int)((B + 255) / 256 < 1024 ? (B + 255) / 256 : 1024), 256, 0, stream>>>(
                    bl, bl_stride, bl_tree_stride, B, (int)k, st);
                count_launch();
            }
        }
    } else {
        const CophLayout L = coph_layout(k);
        unsigned char* tab = wsb + W.tab;
        // decode straight into every tree's table area (per-tree output stride)
        e = launch_decode(MODE_ANCESTRY, vin, vin_bytes, B, k, tab + L.anc, st, wsb + W.dec, W.tab - W.dec,
                          stream, (int64_t)(L.total / sizeof(int)));
        if (e != cudaSuccess) return e;
        if (bl) {
            const int64_t want = (B * k + 255) / 256, capb = (int64_t)sm_count() * 8;
            flag_negative_bl_kernel<<<(int)(want < capb ? want : capb), 256, 0, stream>>>(bl, bl_stride, bl_tree_stride,
                                                                                       B, (int)k, st);
            count_launch();
        }
        const int64_t cap = (int64_t)sm_count() * 2;
        bool launched = false;
        if (grid_team_trees(B, k)) {
            // one to four huge trees: the whole grid works on one tree at a time (cooperative launch)
            launched = launch_grid_team(tables_kernel<0>, TAB_THREADS, 0, grid_team_ctas(2 * k + 1, TAB_THREADS), stream,
                                        tab, (size_t)L.total, bl, bl_stride, bl_tree_stride, B, (int)k, unrooted,
                                        (const int32_t*)st);
        }
        if (!launched && B * 8 <= sm_count() && k >= 4096) {
            // few large trees: a thread-block cluster per tree (16 CTAs, else the portable 8)
            launched = launch_tables_cluster<TAB_CLUSTER>(tab, L.total, bl, bl_stride, bl_tree_stride, B, (int)k,
                                                          unrooted, st, stream) ||
                       launch_tables_cluster<8>(tab, L.total, bl, bl_stride, bl_tree_stride, B, (int)k, unrooted,
                                                st, stream);
        }
        if (!launched && 2 * k + 1 < 65535 && tables_smem_bytes(k, bl != nullptr) <= (size_t)220 * 1024 && tables_smem_on()) {
            // batches of mid-size trees: the jumping state of a tree fits shared memory
            const size_t smem = tables_smem_bytes(k, bl != nullptr);
            auto kern = bl ? tables_smem_kernel<true> : tables_smem_kernel<false>;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            int per_sm = 0;
            static const int th_env = [] { const char* e2 = getenv("P2V_TABLES_SMEM_THREADS"); return e2 ? atoi(e2) : TAB_THREADS; }();
            const int th = th_env;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, th, smem);
            if (e != cudaSuccess) return e;
            const int64_t capt = (int64_t)sm_count() * (per_sm < 1 ? 1 : per_sm);
            kern<<<(int)(B < capt ? B : capt), th, smem, stream>>>(tab, L.total, bl, bl_stride, bl_tree_stride, B,
                                                                  (int)k, unrooted, st);
            launched = true;
        }
        if (!launched) {
            const int grid = (int)(B < cap ? B : cap);
            tables_kernel<1><<<grid, TAB_THREADS, 0, stream>>>(tab, L.total, bl, bl_stride, bl_tree_stride, B,
                                                              (int)k, unrooted, st);
        }
        count_launch();
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (status) {
        e = cudaMemcpyAsync(status, st, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, stream);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

// Phase 3+4: rows [row_begin,row_end) of every prepared tree.
cudaError_t launch_cophenetic_rows(void* ws, size_t ws_bytes, int64_t B, int64_t k, int weighted,
                                   int64_t row_begin, int64_t row_end, double* out, cudaStream_t stream,
                                   int what) {
    if (B <= 0) return cudaSuccess;
    const int64_t n = k + 1;
    if (row_begin < 0 || row_end > n || row_begin > row_end) return cudaErrorInvalidValue;
    if (row_begin == row_end) return cudaSuccess;
    const CophWs W = coph_ws(B, k);
    if (!ws || ws_bytes < W.total) return cudaErrorInvalidValue;
    unsigned char* wsb = static_cast<unsigned char*>(ws);
    const int32_t* st = reinterpret_cast<const int32_t*>(wsb + W.status);
    cudaError_t e;
    if (k < 2) {
        tiny_kernel<<<(int)((B + 255) / 256 < 1024 ? (B + 255) / 256 : 1024), 256, 0, stream>>>(
            reinterpret_cast<const double*>(wsb + W.tab), B, (int)k, row_begin, row_end, out, what);
        count_launch();
        return cudaGetLastError();
    }
    const CophLayout L = coph_layout(k);
    unsigned char* tab = wsb + W.tab;
    const CophCall call = coph_call(n, row_end - row_begin);
    {
        const int64_t cap = (int64_t)sm_count() * 2;
        const int grid = (int)(B < cap ? B : cap);
        bool launched = false, fused = false;
        if (grid_team_trees(B, k)) {
            // one to four huge trees: blocks, block minima and column tables in one cooperative launch
            int ctas = grid_team_ctas(n, BLK_THREADS);
            if (ctas * (BLK_THREADS / 32) > L.nbmax) ctas = L.nbmax / (BLK_THREADS / 32);   // per-warp counts live in tilelist
            if (ctas >= 1)
                launched = fused = launch_grid_team(blocks_kernel<0>, BLK_THREADS, 0, ctas, stream, tab, (size_t)L.total, B,
                                                    (int)k, row_begin, row_end, call.pcap, call.rcap, st, weighted);
        }
        if (!launched && B * 8 <= sm_count() && k >= 4096)    // few large trees: a cluster of 8 CTAs per tree
            launched = launch_blocks_cluster<8>(tab, L.total, B, (int)k, row_begin, row_end, call.pcap, call.rcap, st, stream);
        if (!launched)
            blocks_kernel<1><<<grid, BLK_THREADS, 0, stream>>>(tab, L.total, B, (int)k, row_begin, row_end, call.pcap, call.rcap, st, 0);
        count_launch();
        if (!fused) {
            const int64_t warps = B * (int64_t)L.nbmax;
            const int64_t want = (warps + 7) / 8, cap2 = (int64_t)sm_count() * 8;
            block_minima_kernel<<<(int)(want < cap2 ? want : cap2), 256, 0, stream>>>(tab, L.total, B, (int)k, st);
            count_launch();
            const int64_t want3 = (B * n + 255) / 256;
            column_tables_kernel<<<(int)(want3 < cap2 * 4 ? want3 : cap2 * 4), 256, 0, stream>>>(tab, L.total, B, (int)k,
                                                                                               weighted, st);
            count_launch();
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    const int threads = coph_row_threads((int)n);
    const int chunk_cols = threads * ROW_CPT;
    const int chunks = (int)((n + chunk_cols - 1) / chunk_cols);
    int64_t tiles_bound = call.nidb;
    if (tiles_bound > L.nbmax) tiles_bound = L.nbmax;
    // small matrices: one CTA walks all column chunks of its tile (row-side setup done once)
    // and bigger ones as many as still leave every SM a few CTAs: the per-tile set-up (row side,
    // minima over the blocks in between, sparse table of the block) is then paid once per group
    static const int cta_factor = [] { const char* e = getenv("P2V_ROWS_CTA_FACTOR"); return e ? atoi(e) : 18; }();
    int chunks_per_cta = (chunks <= 4) ? chunks : 1;
    if (chunks > 4) {
        for (int c = 8; c > 1; c >>= 1)
            if (B * tiles_bound * ((chunks + c - 1) / c) >= (int64_t)sm_count() * cta_factor) { chunks_per_cta = c; break; }
    }
    const int64_t work = B * tiles_bound * ((chunks + chunks_per_cta - 1) / chunks_per_cta);
    const size_t smem = rows_dyn_smem(call);
    const bool topo = !weighted;
    const bool vec2 = (n % 2 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const int64_t cap = (int64_t)sm_count() * 96 * (256 / threads);
    const int grid = (int)(work < cap ? work : cap);
#define P2V_ROWS4(T_, V_, TH_)                                                                        \
    do {                                                                                              \
        auto kern = what ? rows_kernel<T_, V_, TH_, true> : rows_kernel<T_, V_, TH_, false>;          \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
        if (e != cudaSuccess) return e;                                                               \
        kern<<<grid, TH_, smem, stream>>>(tab, L.total, B, (int)k, row_begin, row_end, out, st, chunks, \
                                          chunks_per_cta, (int)tiles_bound, call.bsz, call.nidb, call.rcap,   \
                                          call.nincap);                                               \
    } while (0)
#define P2V_ROWS2(T_, V_)                                                                             \
    do {                                                                                              \
        if (threads == 256) P2V_ROWS4(T_, V_, 256);                                                    \
        else if (threads == 128) P2V_ROWS4(T_, V_, 128);                                               \
        else if (threads == 64) P2V_ROWS4(T_, V_, 64);                                                 \
        else P2V_ROWS4(T_, V_, 32);                                                                    \
    } while (0)
    if (topo) { if (vec2) P2V_ROWS2(true, true); else P2V_ROWS2(true, false); }
    else { if (vec2) P2V_ROWS2(false, true); else P2V_ROWS2(false, false); }
#undef P2V_ROWS2
#undef P2V_ROWS4
    count_launch();
    return cudaGetLastError();
}

static int coph_small_n_max() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("P2V_COPH_SMALL_NMAX");       // tuning knob (comparison runs)
        v = e ? atoi(e) : CS_N_DEFAULT;
        if (v > CS_N_MAX) v = CS_N_MAX;
        if (v < 0) v = 0;
    }
    return v;
}

// One-shot path for small trees (3 <= n <= 512): input adapters + the fused kernel.
static cudaError_t launch_cophenetic_small(const void* v_or_matrix, int idx_bytes, const double* bls, int64_t B,
                                           int64_t k, int unrooted, int64_t row_begin, int64_t row_end, double* out,
                                           int32_t* status, void* ws, size_t ws_bytes, cudaStream_t stream,
                                           int what) {
    const CophWs W = coph_ws(B, k);
    if (!ws || ws_bytes < W.total) return cudaErrorInvalidValue;
    unsigned char* wsb = static_cast<unsigned char*>(ws);
    int32_t* st = reinterpret_cast<int32_t*>(wsb + W.status);
    int* v32 = reinterpret_cast<int*>(wsb + W.v32);
    const int grid1d = sm_count() * 8;
    const double* bl = bls;
    int bl_stride = 2;
    int64_t bl_tree_stride = 2 * k;
    const int* vin = static_cast<const int*>(v_or_matrix);
    if (idx_bytes == 0) {
        const double* m = static_cast<const double*>(v_or_matrix);
        bl = m + 1; bl_stride = 3; bl_tree_stride = 3 * k;
        matrix_v_kernel<<<grid1d, 256, 0, stream>>>(m, B * k, v32);
        count_launch();
        vin = v32;
    } else if (idx_bytes == 8) {
        narrow_v_kernel<<<grid1d, 256, 0, stream>>>(static_cast<const long long*>(v_or_matrix), B * k, v32);
        count_launch();
        vin = v32;
    } else if (idx_bytes != 4) {
        return cudaErrorInvalidValue;
    }
    const bool topo = (bl == nullptr);
    const int64_t n = k + 1;
    static const int wt_max = [] { const char* e = getenv("P2V_COPH_SMALL_WT_MAX"); return e ? atoi(e) : 40; }();
    // tiny trees: a warp per tree through all phases (measured: +35 % at n = 16, +15 % at 40, -8 % at 64)
    const bool wt = n <= wt_max && n <= 128;
    const SmallLayout SL = small_layout((int)k, topo, wt);
    const size_t smem = SL.total;
    const int tpp = SL.tpp;
    const bool vec2 = (n % 2 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    // measured: fp64 distance rows are bound by the double-double math either way, so only the
    // integer and vcv rows take the tiled path
    const bool tiled = (n > 160) && (topo || what != 0);
    static const int use_tma = [] { const char* e = getenv("P2V_COPH_SMALL_TMA"); return e ? atoi(e) : 1; }();
    // Branch lengths staged by TMA (see the kernel): measured against direct loads (profiles/r02_coph_small_tma.md)
    // it costs 4 % at n = 100, 2 % at n = 128..320 and more below -- the other CTAs of the SM already hide
    // those loads.  It is on for the one-tree-per-pass sizes, where it costs least; P2V_COPH_SMALL_STAGE_BL=0/1 forces it.
    static const int stage_env = [] { const char* e = getenv("P2V_COPH_SMALL_STAGE_BL"); return e ? atoi(e) : -1; }();
    const int stage_lengths = stage_env >= 0 ? stage_env : (n > 128);
    cudaError_t e = cudaSuccess;
#define P2V_SMALL(T_, W_, V_)                                                                          \
    do {                                                                                               \
        auto kern = tiled ? coph_small_kernel<T_, W_, V_, true, false, false> : coph_small_kernel<T_, W_, V_, false, false, false>; \
        if (!T_ && stage_lengths) kern = tiled ? coph_small_kernel<T_, W_, V_, true, !T_, false> : coph_small_kernel<T_, W_, V_, false, !T_, false>; \
        if (wt) kern = coph_small_kernel<T_, W_, V_, false, false, true>;                                  \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
        if (e != cudaSuccess) return e;                                                                \
        int per_sm = 0;                                                                                \
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CS_THREADS, smem);            \
        if (e != cudaSuccess) return e;                                                                \
        if (per_sm < 1) per_sm = 1;                                                                    \
        const int64_t cap = (int64_t)sm_count() * per_sm, want = (B + tpp - 1) / tpp;                  \
        kern<<<(int)(want < cap ? want : cap), CS_THREADS, smem, stream>>>(vin, bl, bl_stride, bl_tree_stride, B, (int)k, \
                                                                     unrooted, row_begin, row_end, out, st, use_tma); \
    } while (0)
#define P2V_SMALL2(T_, W_) do { if (vec2) P2V_SMALL(T_, W_, true); else P2V_SMALL(T_, W_, false); } while (0)
    if (topo) { if (what) P2V_SMALL2(true, true); else P2V_SMALL2(true, false); }
    else { if (what) P2V_SMALL2(false, true); else P2V_SMALL2(false, false); }
#undef P2V_SMALL2
#undef P2V_SMALL
    count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (bl) {      // the matrices of such trees are already written and meaningless; the status says so
        const int64_t want = (B * k + 255) / 256, capb = (int64_t)sm_count() * 8;
        flag_negative_bl_kernel<<<(int)(want < capb ? want : capb), 256, 0, stream>>>(bl, bl_stride, bl_tree_stride, B,
                                                                                   (int)k, st);
        count_launch();
    }
    if (status) e = cudaMemcpyAsync(status, st, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, stream);
    return e;
}

cudaError_t launch_cophenetic(const void* v_or_matrix, int idx_bytes, const double* bls, int64_t B,
                              int64_t k, int unrooted, int64_t row_begin, int64_t row_end,
                              double* out, int32_t* status, void* ws, size_t ws_bytes,
                              cudaStream_t stream, int what) {
    if (B <= 0) return cudaSuccess;
    if (row_begin < 0 || row_end > k + 1 || row_begin > row_end) return cudaErrorInvalidValue;
    if (k >= 2 && k + 1 <= coph_small_n_max())
        return launch_cophenetic_small(v_or_matrix, idx_bytes, bls, B, k, unrooted, row_begin, row_end, out, status,
                                       ws, ws_bytes, stream, what);
    cudaError_t e = launch_cophenetic_prepare(v_or_matrix, idx_bytes, bls, B, k, unrooted, status, ws, ws_bytes, stream);
    if (e != cudaSuccess) return e;
    const int weighted = (idx_bytes == 0) || (bls != nullptr);
    return launch_cophenetic_rows(ws, ws_bytes, B, k, weighted, row_begin, row_end, out, stream, what);
}

// Depths of all 2k+1 nodes of every tree, (B, 2k+1) float64: prepare + one gather.
cudaError_t launch_node_depths(const void* v_or_matrix, int idx_bytes, const double* bls, int64_t B, int64_t k,
                               double* out, int32_t* status, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    cudaError_t e = launch_cophenetic_prepare(v_or_matrix, idx_bytes, bls, B, k, 0, status, ws, ws_bytes, stream);
    if (e != cudaSuccess) return e;
    const CophWs W = coph_ws(B, k);
    unsigned char* wsb = static_cast<unsigned char*>(ws);
    const int weighted = (idx_bytes == 0) || (bls != nullptr);
    const int64_t want = (B * (2 * k + 1) + 255) / 256, cap = (int64_t)sm_count() * 16;
    node_depths_kernel<<<(int)(want < cap ? want : cap), 256, 0, stream>>>(
        wsb + W.tab, k >= 2 ? coph_layout(k).total : 16, B, (int)k, weighted, out,
        reinterpret_cast<const int32_t*>(wsb + W.status));
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_fill_f64(double* out, int64_t count, double value, cudaStream_t stream) {
    if (count <= 0) return cudaSuccess;
    fill_f64_kernel<<<sm_count() * 6, 256, 0, stream>>>(out, count, value);
    count_launch();
    return cudaGetLastError();
}

}  // namespace p2v
