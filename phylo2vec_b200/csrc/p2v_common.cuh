// Shared device helpers for the phylo2vec B200 kernels (sm_100a).
#pragma once
#include <cstdio>
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace p2v {

constexpr unsigned FULL = 0xFFFFFFFFu;
// -DP2V_DEBUG_BOUNDS (python -m phylo2vec_b200.build --debug-bounds -> libphylo2vec_b200_dbg.so): every index into a
// shared-memory table or a parked scratch word of the decode / encode kernels is checked; a bad one prints its
// location and traps, which fails the launch (tests/test_gpu_debug_bounds.py runs the boundary sizes this way;
// compute-sanitizer is closed on the GPU pool).
#ifdef P2V_DEBUG_BOUNDS
#define P2V_CHECK_INDEX(i, n)                                                                                   \
    do {                                                                                                        \
        if (!((unsigned long long)(long long)(i) < (unsigned long long)(long long)(n))) {                       \
            printf("p2v bounds: %s:%d index %lld, limit %lld\n", __FILE__, __LINE__, (long long)(i), (long long)(n)); \
            __trap();                                                                                           \
        }                                                                                                       \
    } while (0)
#else
#define P2V_CHECK_INDEX(i, n) do { } while (0)
#endif

constexpr uint32_t NONE16 = 0xFFFFu;
constexpr uint32_t NONE32 = 0xFFFFFFFFu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// lanes strictly above / at-or-below `lane`
__device__ __forceinline__ uint32_t lanes_gt(int lane) { return ~((2u << lane) - 1u); }
__device__ __forceinline__ uint32_t lanes_le(int lane) { return (2u << lane) - 1u; }
__device__ __forceinline__ uint32_t lanes_lt(int lane) { return (1u << lane) - 1u; }

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t x, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t y = __shfl_up_sync(FULL, x, d);
        if (lane >= d) x += y;
    }
    return x;
}

// position of the r-th (0-based) set bit of wd; r < popc(wd)
__device__ __forceinline__ uint32_t nth_set_bit(uint32_t wd, uint32_t r) {
    uint32_t pos = 0, c;
    c = __popc(wd & 0xFFFFu);          if (r >= c) { r -= c; pos = 16; }
    c = __popc((wd >> pos) & 0xFFu);   if (r >= c) { r -= c; pos += 8; }
    c = __popc((wd >> pos) & 0xFu);    if (r >= c) { r -= c; pos += 4; }
    c = __popc((wd >> pos) & 0x3u);    if (r >= c) { r -= c; pos += 2; }
    c = (wd >> pos) & 1u;              if (r >= c) { pos += 1; }
    return pos;
}

// The same through a 2 KB table in shared memory for the last three levels:
// lut[byte * 8 + r] = position of the r-th set bit of `byte`.
__device__ __forceinline__ void nth_bit_lut_init(uint8_t* lut, int tid, int nthreads) {
    for (int e = tid; e < 2048; e += nthreads) {
        uint32_t byte = (uint32_t)e >> 3, r = (uint32_t)e & 7u, pos = 0;
        for (uint32_t i = 0; i < 8; ++i) {
            if ((byte >> i) & 1u) {
                if (r == 0) { pos = i; break; }
                --r;
            }
        }
        lut[e] = (uint8_t)pos;
    }
}
__device__ __forceinline__ uint32_t nth_set_bit_lut(uint32_t wd, uint32_t r, const uint8_t* __restrict__ lut) {
    uint32_t pos = 0, c;
    c = __popc(wd & 0xFFFFu);          if (r >= c) { r -= c; pos = 16; }
    c = __popc((wd >> pos) & 0xFFu);   if (r >= c) { r -= c; pos += 8; }
    return pos + lut[(((wd >> pos) & 0xFFu) << 3) | (r & 7u)];
}

// One step of the in-batch rank scan: lane i looks at lane i+d (if it exists)
// and bumps R when that later element was inserted at or before R.
// SHFL.DOWN delivers the "source lane in range" predicate for free.
__device__ __forceinline__ void rank_step(uint32_t& R, uint32_t ins, int d) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 x;\n\t"
        "shfl.sync.down.b32 x|p, %1, %2, 31, 0xffffffff;\n\t"
        "setp.le.and.u32 p, x, %0, p;\n\t"
        "@p add.u32 %0, %0, 1;\n\t"
        "}\n"
        : "+r"(R)
        : "r"(ins), "r"(d));
}

// ---- TMA 1-D bulk copies global -> shared, completion on an mbarrier (cp.async.bulk; SASS UBLKCP) -------------
// One thread arms the barrier with the byte count and issues the copy; everybody waits on the phase parity.
// Source and destination 16-byte aligned, size a multiple of 16.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_global, uint32_t bytes, uint64_t* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_global), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra W_%=;\n\t"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// lane i counts earlier lanes (i-d) whose value is smaller (order-free count)
__device__ __forceinline__ void count_smaller_step(uint32_t& cnt, uint32_t mine, int d) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 x;\n\t"
        "shfl.sync.up.b32 x|p, %1, %2, 0, 0xffffffff;\n\t"
        "setp.lt.and.u32 p, x, %1, p;\n\t"
        "@p add.u32 %0, %0, 1;\n\t"
        "}\n"
        : "+r"(cnt)
        : "r"(mine), "r"(d));
}

// ---- phase stamps of the team kernels (profiling aid) -----------------------------------------
// Thread 0 of team CTA 0 stores the global nanosecond timer at phase boundaries; p2v_debug_phase_ns()
// copies the stamps out.  One predicated store per phase: left in release builds.
constexpr int PHASE_SLOTS = 256;
static __device__ unsigned long long g_phase_ns[PHASE_SLOTS];    // one copy per translation unit
__device__ __forceinline__ void phase_stamp(int slot, bool who) {
    if (who) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_phase_ns[slot] = t;
    }
}

// ---- teams: the CTAs that work on one tree ------------------------------------------------
// CSIZE = 1: one CTA; CSIZE > 1: a thread-block cluster of CSIZE CTAs (hardware cluster barrier);
// CSIZE = 0: the whole grid of a cooperative launch (one huge tree on all SMs, grid barrier).
template <int CSIZE>
__device__ __forceinline__ void team_sync() {
    if (CSIZE == 1) __syncthreads();
    else if (CSIZE == 0) {
        cooperative_groups::this_grid().sync();
        // the warp of the CTA's master thread comes out of the barrier's spin DIVERGED (measured: its
        // shuffle-heavy code then runs ten times slower until something reconverges it)
        __syncwarp();
    }
    else cooperative_groups::this_cluster().sync();
}
template <int CSIZE>
__device__ __forceinline__ int team_rank() {
    if (CSIZE == 1) return 0;
    if (CSIZE == 0) return (int)blockIdx.x;
    return (int)cooperative_groups::this_cluster().block_rank();
}
// CTAs per team, number of teams in the grid, index of this CTA's team
template <int CSIZE>
__device__ __forceinline__ int team_ctas() { return CSIZE == 0 ? (int)gridDim.x : CSIZE; }
template <int CSIZE>
__device__ __forceinline__ int64_t team_count() { return CSIZE == 0 ? 1 : (int64_t)(gridDim.x / CSIZE); }
template <int CSIZE>
__device__ __forceinline__ int64_t team_index() { return CSIZE == 0 ? 0 : (int64_t)(blockIdx.x / CSIZE); }
// Whole-grid barrier that also ORs a predicate: every CTA adds (any << 16) + 1 to a fresh zero word
// and waits until the low half counts all CTAs -- one atomic and one spin instead of a flag
// update, a grid barrier and a flag read.  Same fence / barrier structure as the cooperative-groups
// grid barrier, so data written before it is visible after it.
__device__ __forceinline__ int grid_barrier_any(int any_cta, unsigned* word) {
    __shared__ unsigned s_seen;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(word, (any_cta ? 0x10000u : 0u) + 1u);
        unsigned v;
        unsigned spins = 0;         // bounded: a lost CTA must not hang the device (the result is then garbage, never a hang)
        do { v = *reinterpret_cast<volatile unsigned*>(word); } while ((v & 0xFFFFu) < gridDim.x && ++spins < (1u << 22));
        __threadfence();
        s_seen = (spins >= (1u << 22)) ? 0u : (v >> 16);
    }
    __syncwarp();
    __syncthreads();
    const int r = (int)s_seen;
    __syncthreads();          // s_seen may be rewritten by the next round
    return r;
}
// Whole-grid barrier on a monotone counter word: every barrier on the word adds gridDim.x arrivals, so
// the word never needs a reset between barriers or launches -- it only has to START at a multiple of
// gridDim.x (zero).  Bounded spin like grid_barrier_any.
__device__ __forceinline__ void grid_barrier(unsigned* word) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned G = gridDim.x;
        const unsigned old = atomicAdd(word, 1u);
        const unsigned target = old - old % G + G;
        unsigned spins = 0;
        while ((int)(*reinterpret_cast<volatile unsigned*>(word) - target) < 0 && ++spins < (1u << 22)) {}
        __threadfence();
    }
    __syncwarp();
    __syncthreads();
}
// OR of `live` over the whole team; `flag` is a zero-initialised word owned by the team
template <int CSIZE>
__device__ __forceinline__ int team_any(int live, int* flag) {
    const int any_cta = __syncthreads_or(live);
    if (CSIZE == 1) return any_cta;
    if (CSIZE == 0) return grid_barrier_any(any_cta, reinterpret_cast<unsigned*>(flag));
    if (threadIdx.x == 0 && any_cta) atomicOr(flag, 1);
    team_sync<CSIZE>();
    return *reinterpret_cast<volatile int*>(flag);
}

// Cooperative launch of a CSIZE = 0 kernel: at most `want_ctas` CTAs, never more than are co-resident
// (the grid barrier needs every CTA on an SM).  false: not launched (the caller falls back).
template <typename Kern, typename... Args>
static inline bool launch_grid_team(Kern kern, int threads, size_t smem, int want_ctas, cudaStream_t stream,
                                    Args... args) {
    int dev = 0, coop = 0, sms = 0, per_sm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return false; }
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (!coop || sms <= 0) return false;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        return false;
    }
    int grid = want_ctas < 1 ? 1 : want_ctas;
    if (grid > sms * per_sm) grid = sms * per_sm;
    void* argv[] = {(void*)&args...};
    if (cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)grid), dim3((unsigned)threads), argv, smem,
                                    stream) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return true;
}

// ---- double-double helpers (fp64 branch-length depths) --------------------
struct dd { double hi, lo; };

__device__ __forceinline__ dd two_sum(double a, double b) {
    double s = __dadd_rn(a, b);
    double bb = __dadd_rn(s, -a);
    double e = __dadd_rn(__dadd_rn(a, -__dadd_rn(s, -bb)), __dadd_rn(b, -bb));
    return {s, e};
}
__device__ __forceinline__ dd quick_two_sum(double a, double b) {
    double s = __dadd_rn(a, b);
    double e = __dadd_rn(b, -__dadd_rn(s, -a));
    return {s, e};
}
__device__ __forceinline__ dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    dd t = two_sum(a.lo, b.lo);
    s.lo = __dadd_rn(s.lo, t.hi);
    s = quick_two_sum(s.hi, s.lo);
    s.lo = __dadd_rn(s.lo, t.lo);
    return quick_two_sum(s.hi, s.lo);
}
__device__ __forceinline__ dd dd_neg2(dd a) { return {-2.0 * a.hi, -2.0 * a.lo}; }

}  // namespace p2v
