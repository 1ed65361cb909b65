// C ABI of the library (declared in include/phylo2vec_b200.h).
#include <emmintrin.h>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/phylo2vec_b200.h"
#include "p2v_internal.h"

namespace p2v {

static std::atomic<int64_t> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
    static thread_local int cached_dev = -1;
    static thread_local int cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached = v; cached_dev = dev;
    }
    return cached;
}

static thread_local std::string g_err;

static int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
    g_err = what;
    if (e != cudaSuccess) { g_err += ": "; g_err += cudaGetErrorString(e); }
    return code;
}

static int from_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return P2V_OK;
    if (e == cudaErrorNotSupported) return fail(P2V_ERR_UNSUPPORTED, what, e);
    if (e == cudaErrorInvalidValue) return fail(P2V_ERR_INVALID_ARG, what, e);
    return fail(P2V_ERR_CUDA, what, e);
}

static int check_common(const void* in, const void* out, int idx_bytes, int64_t B, int64_t k) {
    if (B < 0 || k < 0) return fail(P2V_ERR_INVALID_ARG, "negative batch or length");
    if (idx_bytes != 4 && idx_bytes != 8) return fail(P2V_ERR_INVALID_ARG, "idx_bytes must be 4 or 8");
    if (B > 0 && k > 0 && (!in || !out)) return fail(P2V_ERR_INVALID_ARG, "null buffer");
    if (k > LARGE_K_MAX) return fail(P2V_ERR_UNSUPPORTED, "k above the supported maximum (2^19)");
    return P2V_OK;
}

static int decode_entry(int mode, const void* v, int idx_bytes, int64_t B, int64_t k, void* out,
                        int32_t* status, void* ws, size_t ws_bytes, p2v_stream_t stream) {
    int rc = check_common(v, out, idx_bytes, B, k);
    if (rc) return rc;
    if (ws_bytes < decode_workspace_bytes(B, k)) return fail(P2V_ERR_WORKSPACE, "decode workspace too small");
    // pairs and edges are written as (x, y) vector stores: 2 * idx_bytes alignment (see the header)
    if (mode != MODE_ANCESTRY && B > 0 && k > 0 && (reinterpret_cast<uintptr_t>(out) % (2 * (size_t)idx_bytes)) != 0)
        return fail(P2V_ERR_INVALID_ARG, "pairs / edges output must be aligned to 2 * idx_bytes bytes");
    if ((reinterpret_cast<uintptr_t>(v) % (size_t)idx_bytes) != 0 || (reinterpret_cast<uintptr_t>(out) % (size_t)idx_bytes) != 0)
        return fail(P2V_ERR_INVALID_ARG, "index buffers must be aligned to idx_bytes");
    return from_cuda(launch_decode(mode, v, idx_bytes, B, k, out, status, ws, ws_bytes, (cudaStream_t)stream),
                     "decode launch");
}

static int encode_entry(int mode, const void* in, int idx_bytes, int64_t B, int64_t k, void* v,
                        int32_t* status, void* ws, size_t ws_bytes, p2v_stream_t stream) {
    int rc = check_common(in, v, idx_bytes, B, k);
    if (rc) return rc;
    if (ws_bytes < encode_workspace_bytes(B, k)) return fail(P2V_ERR_WORKSPACE, "encode workspace too small");
    return from_cuda(launch_encode(mode, in, idx_bytes, B, k, v, status, ws, ws_bytes, (cudaStream_t)stream),
                     "encode launch");
}

// ---- host-buffer plumbing --------------------------------------------------------------
// The batch is cut into chunks that rotate over three streams, so that (with pinned host memory)
// the H2D copy, the kernels and the D2H copy of neighbouring chunks overlap.  Streams and device
// staging buffers are cached per host thread and grow on demand; p2v_host_release() frees them.
constexpr int HOST_STREAMS = 3;
// The thread-local caches below free their memory when a worker thread ends.  The main thread's destructors
// only run while the process exits, when the CUDA driver may already be gone: nothing is called then (the
// Python layer releases the main thread's cache from an atexit handler, which runs earlier; the OS has the rest).
static bool at_process_exit() { return (long)syscall(SYS_gettid) == (long)getpid(); }

struct HostCtx {
    int dev = -1;
    cudaStream_t st[HOST_STREAMS] = {};
    void* din[HOST_STREAMS] = {};
    void* dout[HOST_STREAMS] = {};
    void* dws[HOST_STREAMS] = {};
    int32_t* dstat[HOST_STREAMS] = {};
    size_t cap_in = 0, cap_out = 0, cap_ws = 0, cap_stat = 0;
    void release() {
        for (int i = 0; i < HOST_STREAMS; ++i) {
            if (st[i]) cudaStreamSynchronize(st[i]);
            if (din[i]) cudaFree(din[i]);
            if (dout[i]) cudaFree(dout[i]);
            if (dws[i]) cudaFree(dws[i]);
            if (dstat[i]) cudaFree(dstat[i]);
            if (st[i]) cudaStreamDestroy(st[i]);
            st[i] = nullptr; din[i] = dout[i] = dws[i] = nullptr; dstat[i] = nullptr;
        }
        cap_in = cap_out = cap_ws = cap_stat = 0;
        dev = -1;
    }
    size_t bytes() const { return HOST_STREAMS * (cap_in + cap_out + cap_ws + cap_stat); }
    // thread exit: give the memory back (see at_process_exit)
    ~HostCtx() { if (!at_process_exit()) { release(); cudaGetLastError(); } }
};
static thread_local HostCtx g_host;
// Staging the *_host entry points may keep between calls, per host thread (device + pinned bytes).
// A call that leaves more than this behind frees its staging before it returns.
static std::atomic<int64_t> g_host_cache_limit{(int64_t)512 << 20};

static cudaError_t host_ensure(HostCtx& H, size_t in_b, size_t out_b, size_t ws_b, size_t stat_b) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (H.dev != dev) { if (H.dev >= 0) H.release(); H.dev = dev; }
    for (int i = 0; i < HOST_STREAMS; ++i) {
        if (!H.st[i]) { e = cudaStreamCreateWithFlags(&H.st[i], cudaStreamNonBlocking); if (e != cudaSuccess) return e; }
        if (in_b > H.cap_in || (in_b && !H.din[i])) {
            if (H.din[i]) cudaFree(H.din[i]);
            H.din[i] = nullptr;
            e = cudaMalloc(&H.din[i], in_b); if (e != cudaSuccess) return e;
        }
        if (out_b > H.cap_out || (out_b && !H.dout[i])) {
            if (H.dout[i]) cudaFree(H.dout[i]);
            H.dout[i] = nullptr;
            e = cudaMalloc(&H.dout[i], out_b); if (e != cudaSuccess) return e;
        }
        if (ws_b > H.cap_ws || (ws_b && !H.dws[i])) {
            if (H.dws[i]) cudaFree(H.dws[i]);
            H.dws[i] = nullptr;
            e = cudaMalloc(&H.dws[i], ws_b); if (e != cudaSuccess) return e;
        }
        if (stat_b > H.cap_stat || !H.dstat[i]) {
            if (H.dstat[i]) cudaFree(H.dstat[i]);
            H.dstat[i] = nullptr;
            e = cudaMalloc((void**)&H.dstat[i], stat_b); if (e != cudaSuccess) return e;
        }
    }
    if (in_b > H.cap_in) H.cap_in = in_b;
    if (out_b > H.cap_out) H.cap_out = out_b;
    if (ws_b > H.cap_ws) H.cap_ws = ws_b;
    if (stat_b > H.cap_stat) H.cap_stat = stat_b;
    return cudaSuccess;
}

static int host_thread_count();
static bool is_pageable(const void* p);
static bool bind_thread_near_gpu();

// Pinned staging of the generic host path: a pageable buffer (a plain numpy array) would be staged
// by the driver on one thread at ~10 GB/s; here worker threads copy it through pinned buffers of
// the library, chunk by chunk, while the other chunks are in flight.
struct StageCtx {
    void* hin[HOST_STREAMS] = {};
    void* hout[HOST_STREAMS] = {};
    int32_t* hstat[HOST_STREAMS] = {};
    cudaEvent_t ev[HOST_STREAMS] = {};
    size_t cap_in = 0, cap_out = 0, cap_stat = 0;
    size_t bytes() const { return HOST_STREAMS * (cap_in + cap_out + cap_stat); }
    ~StageCtx() { if (!at_process_exit()) { release(); cudaGetLastError(); } }
    void release() {
        for (int i = 0; i < HOST_STREAMS; ++i) {
            if (hin[i]) cudaFreeHost(hin[i]);
            if (hout[i]) cudaFreeHost(hout[i]);
            if (hstat[i]) cudaFreeHost(hstat[i]);
            if (ev[i]) cudaEventDestroy(ev[i]);
            hin[i] = hout[i] = nullptr; hstat[i] = nullptr; ev[i] = nullptr;
        }
        cap_in = cap_out = cap_stat = 0;
    }
};
static thread_local StageCtx g_stage;

struct CopyJob { const void* src; void* dst; size_t bytes; int chunk; int kind; };   // kind 0: out, 1: in

template <typename Launch>
static int host_chunked_staged(const void* in, size_t in_tree_bytes, void* out, size_t out_tree_bytes,
                               int32_t* status, int64_t B, int64_t chunk, size_t wsb, bool page_in, bool page_out,
                               Launch launch) {
    HostCtx& H = g_host;
    StageCtx& S = g_stage;
    cudaError_t e = cudaSuccess;
    const size_t in_b = in_tree_bytes * chunk, out_b = out_tree_bytes * chunk, st_b = sizeof(int32_t) * (size_t)chunk;
    for (int i = 0; i < HOST_STREAMS && e == cudaSuccess; ++i) {
        if (page_in && (in_b > S.cap_in || !S.hin[i])) {
            if (S.hin[i]) cudaFreeHost(S.hin[i]);
            S.hin[i] = nullptr;
            e = cudaHostAlloc(&S.hin[i], in_b, cudaHostAllocDefault);
        }
        if (e == cudaSuccess && page_out && (out_b > S.cap_out || !S.hout[i])) {
            if (S.hout[i]) cudaFreeHost(S.hout[i]);
            S.hout[i] = nullptr;
            e = cudaHostAlloc(&S.hout[i], out_b, cudaHostAllocDefault);
        }
        if (e == cudaSuccess && (st_b > S.cap_stat || !S.hstat[i])) {
            if (S.hstat[i]) cudaFreeHost(S.hstat[i]);
            S.hstat[i] = nullptr;
            e = cudaHostAlloc((void**)&S.hstat[i], st_b, cudaHostAllocDefault);
        }
        if (e == cudaSuccess && !S.ev[i]) e = cudaEventCreateWithFlags(&S.ev[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { S.release(); return from_cuda(e, "host entry: staging allocation"); }
    if (page_in && in_b > S.cap_in) S.cap_in = in_b;
    if (page_out && out_b > S.cap_out) S.cap_out = out_b;
    if (st_b > S.cap_stat) S.cap_stat = st_b;

    const int nthreads = host_thread_count();
    int dev_of_call = 0;
    cudaGetDevice(&dev_of_call);
    const int64_t nchunks = (B + chunk - 1) / chunk;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::deque<CopyJob> jobs;
    std::vector<int> pend_out((size_t)nchunks, 0), pend_in((size_t)nchunks, 0);
    bool stop = false;
    std::vector<std::thread> pool;
    pool.reserve(nthreads);
    for (int t = 0; t < nthreads; ++t) {
        pool.emplace_back([&, dev_of_call]() {
            cudaSetDevice(dev_of_call);
            bind_thread_near_gpu();
            for (;;) {
                CopyJob j;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv_job.wait(lk, [&] { return stop || !jobs.empty(); });
                    if (jobs.empty()) return;
                    j = jobs.front(); jobs.pop_front();
                }
                memcpy(j.dst, j.src, j.bytes);
                {
                    std::lock_guard<std::mutex> lk(mu);
                    int& left = j.kind ? pend_in[(size_t)j.chunk] : pend_out[(size_t)j.chunk];
                    if (--left == 0) cv_done.notify_all();
                }
            }
        });
    }
    auto enqueue_copy = [&](const void* src, void* dst, size_t bytes, int64_t c, int kind) {
        const size_t per = ((bytes + nthreads - 1) / nthreads + 63) & ~(size_t)63;
        std::lock_guard<std::mutex> lk(mu);
        for (size_t o = 0; o < bytes; o += per) {
            jobs.push_back(CopyJob{static_cast<const char*>(src) + o, static_cast<char*>(dst) + o,
                                   (bytes - o < per) ? (bytes - o) : per, (int)c, kind});
            ++(kind ? pend_in : pend_out)[(size_t)c];
        }
        cv_job.notify_all();
    };
    auto wait_for = [&](std::vector<int>& pend, int64_t c) {
        if (c < 0) return;
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pend[(size_t)c] == 0; });
    };
    auto chunk_trees = [&](int64_t c) { const int64_t b0 = c * chunk; return (B - b0 < chunk) ? (B - b0) : chunk; };
    auto landed = [&](int64_t c) {              // chunk c is complete on the host side of the wire
        const int s = (int)(c % HOST_STREAMS);
        const int64_t b0 = c * chunk, nb = chunk_trees(c);
        if (status) memcpy(status + b0, S.hstat[s], sizeof(int32_t) * nb);
        if (page_out && out_tree_bytes)
            enqueue_copy(S.hout[s], static_cast<char*>(out) + (size_t)b0 * out_tree_bytes, out_tree_bytes * nb, c, 0);
    };

    int rc = P2V_OK;
    if (page_in && in_tree_bytes) enqueue_copy(in, S.hin[0], in_tree_bytes * chunk_trees(0), 0, 1);
    for (int64_t c = 0; c < nchunks; ++c) {
        const int s = (int)(c % HOST_STREAMS);
        const int64_t b0 = c * chunk, nb = chunk_trees(c);
        if (page_out) wait_for(pend_out, c - HOST_STREAMS);           // staging s has been copied out
        const void* src = static_cast<const char*>(in) + (size_t)b0 * in_tree_bytes;
        if (page_in && in_tree_bytes) {
            wait_for(pend_in, c);
            if (c + 1 < nchunks)                 // staging (c+1)%3 was read by chunk c-2, complete since the last iteration
                enqueue_copy(static_cast<const char*>(in) + (size_t)(c + 1) * chunk * in_tree_bytes,
                             S.hin[(c + 1) % HOST_STREAMS], in_tree_bytes * chunk_trees(c + 1), c + 1, 1);
            src = S.hin[s];
        }
        e = cudaSuccess;
        if (in_tree_bytes) e = cudaMemcpyAsync(H.din[s], src, in_tree_bytes * nb, cudaMemcpyHostToDevice, H.st[s]);
        if (e == cudaSuccess) e = launch(H.din[s], H.dout[s], H.dstat[s], nb, H.dws[s], wsb, H.st[s]);
        if (e == cudaSuccess && out_tree_bytes)
            e = cudaMemcpyAsync(page_out ? S.hout[s] : static_cast<void*>(static_cast<char*>(out) + (size_t)b0 * out_tree_bytes),
                                H.dout[s], out_tree_bytes * nb, cudaMemcpyDeviceToHost, H.st[s]);
        if (e == cudaSuccess && status)
            e = cudaMemcpyAsync(S.hstat[s], H.dstat[s], sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, H.st[s]);
        if (e == cudaSuccess) e = cudaEventRecord(S.ev[s], H.st[s]);
        if (e == cudaSuccess && c >= 1) {
            e = cudaEventSynchronize(S.ev[(c - 1) % HOST_STREAMS]);
            if (e == cudaSuccess) landed(c - 1);
        }
        if (e != cudaSuccess) { rc = from_cuda(e, "host entry: chunk"); break; }
    }
    if (rc == P2V_OK) {
        e = cudaEventSynchronize(S.ev[(nchunks - 1) % HOST_STREAMS]);
        if (e == cudaSuccess) landed(nchunks - 1);
        else rc = from_cuda(e, "host entry: synchronize");
    }
    if (rc == P2V_OK && page_out)
        for (int64_t c = (nchunks > HOST_STREAMS ? nchunks - HOST_STREAMS : 0); c < nchunks; ++c) wait_for(pend_out, c);
    { std::lock_guard<std::mutex> lk(mu); stop = true; }
    cv_job.notify_all();
    for (auto& th : pool) th.join();
    for (int i = 0; i < HOST_STREAMS; ++i) cudaStreamSynchronize(H.st[i]);
    return rc;
}

template <typename Launch, typename WsBytes>
static int host_chunked(const void* in, size_t in_tree_bytes, void* out, size_t out_tree_bytes,
                        int32_t* status, int64_t B, WsBytes ws_bytes_of, Launch launch) {
    if (B <= 0) return P2V_OK;
    const size_t budget = (size_t)1 << 30;  // device bytes per chunk (in + out)
    size_t per = in_tree_bytes + out_tree_bytes;
    if (per == 0) per = 1;
    int64_t chunk = (int64_t)(budget / per);
    if (chunk < 1) chunk = 1;
    if (chunk > B) chunk = B;
    HostCtx& H = g_host;
    // the workspace of a short tail chunk can exceed the full chunk's (a few large trees take the
    // whole-GPU kernels with per-worker tables): size for both
    size_t wsb = ws_bytes_of(chunk);
    if (B % chunk != 0) { const size_t wt = ws_bytes_of(B % chunk); if (wt > wsb) wsb = wt; }
    cudaError_t e = host_ensure(H, in_tree_bytes * chunk, out_tree_bytes * chunk, wsb, sizeof(int32_t) * chunk);
    if (e != cudaSuccess) { H.release(); return from_cuda(e, "host entry: device allocation"); }
    // big calls on pageable buffers: threaded staging through pinned memory
    if ((in_tree_bytes + out_tree_bytes) * (size_t)B >= ((size_t)32 << 20)) {
        const bool page_in = in_tree_bytes && in && is_pageable(in);
        const bool page_out = out_tree_bytes && out && is_pageable(out);
        if (page_in || page_out) {
            int64_t cs = (int64_t)(((size_t)256 << 20) / per);      // smaller chunks: the staging is pinned memory
            if (cs < 1) cs = 1;
            if (cs > chunk) cs = chunk;
            return host_chunked_staged(in, in_tree_bytes, out, out_tree_bytes, status, B, cs, wsb, page_in, page_out, launch);
        }
    }
    int c = 0;
    for (int64_t b0 = 0; b0 < B; b0 += chunk, ++c) {
        const int s = c % HOST_STREAMS;
        const int64_t nb = (B - b0 < chunk) ? (B - b0) : chunk;
        e = cudaSuccess;
        if (in_tree_bytes)
            e = cudaMemcpyAsync(H.din[s], (const char*)in + (size_t)b0 * in_tree_bytes, in_tree_bytes * nb,
                                cudaMemcpyHostToDevice, H.st[s]);
        if (e == cudaSuccess) e = launch(H.din[s], H.dout[s], H.dstat[s], nb, H.dws[s], wsb, H.st[s]);
        if (e == cudaSuccess && out_tree_bytes)
            e = cudaMemcpyAsync((char*)out + (size_t)b0 * out_tree_bytes, H.dout[s], out_tree_bytes * nb,
                                cudaMemcpyDeviceToHost, H.st[s]);
        if (e == cudaSuccess && status)
            e = cudaMemcpyAsync(status + b0, H.dstat[s], sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, H.st[s]);
        if (e != cudaSuccess) { for (int i = 0; i < HOST_STREAMS; ++i) cudaStreamSynchronize(H.st[i]); return from_cuda(e, "host entry: chunk"); }
    }
    for (int i = 0; i < HOST_STREAMS; ++i) {
        e = cudaStreamSynchronize(H.st[i]);
        if (e != cudaSuccess) return from_cuda(e, "host entry: synchronize");
    }
    return P2V_OK;
}

static int decode_host(int mode, int cols, const void* v, int idx_bytes, int64_t B, int64_t k, void* out,
                       int32_t* status) {
    int rc = check_common(v, out, idx_bytes, B, k);
    if (rc) return rc;
    if (k == 0) { if (status && B > 0) memset(status, 0, sizeof(int32_t) * B); return P2V_OK; }
    return host_chunked(
        v, (size_t)k * idx_bytes, out, (size_t)k * cols * idx_bytes, status, B,
        [&](int64_t nb) { return decode_workspace_bytes(nb, k); },
        [&](void* din, void* dout, int32_t* dst, int64_t nb, void* ws, size_t wsb, cudaStream_t st) {
            return launch_decode(mode, din, idx_bytes, nb, k, dout, dst, ws, wsb, st);
        });
}

// ---- to_ancestry through host buffers, int64: compact transport ------------------------------
// The (k,3) int64 ancestry is 24 bytes per row on the wire, but its third column is always
// k+1+row and the first two fit 32 bits.  For big batches the device therefore produces the two
// informative columns as int32 (MODE_ANC2, 8 bytes per row), they cross PCIe into a pinned
// staging buffer of the library, and host threads widen them into the caller's int64 rows with
// streaming stores while the next chunks are in flight.  PCIe carries a third of the bytes; the
// widening runs at host-memory speed (profiles/host_widen_probe.c: 115 GB/s written with 16
// threads on the GPU box).  P2V_HOST_COMPACT=0 switches it off, P2V_HOST_THREADS sets the threads.
struct CompactCtx {
    void* stage[HOST_STREAMS] = {};
    int32_t* vstage[HOST_STREAMS] = {};     // pinned: int32 copy of a pageable input chunk
    int32_t* hstat[HOST_STREAMS] = {};      // pinned: the status of a chunk (the caller's array may be pageable)
    int32_t* v32[HOST_STREAMS] = {};
    cudaEvent_t ev[HOST_STREAMS] = {};
    size_t cap_stage = 0, cap_v32 = 0, cap_hstat = 0, cap_vstage = 0;
    size_t bytes() const { return HOST_STREAMS * (cap_stage + cap_v32 + cap_hstat + cap_vstage); }
    ~CompactCtx() { if (!at_process_exit()) { release(); cudaGetLastError(); } }
    void release() {
        for (int i = 0; i < HOST_STREAMS; ++i) {
            if (stage[i]) cudaFreeHost(stage[i]);
            if (vstage[i]) cudaFreeHost(vstage[i]);
            if (hstat[i]) cudaFreeHost(hstat[i]);
            if (v32[i]) cudaFree(v32[i]);
            if (ev[i]) cudaEventDestroy(ev[i]);
            stage[i] = nullptr; vstage[i] = nullptr; hstat[i] = nullptr; v32[i] = nullptr; ev[i] = nullptr;
        }
        cap_stage = cap_v32 = cap_hstat = cap_vstage = 0;
    }
};
static thread_local CompactCtx g_compact;

// End of every *_host entry point: staging above the limit goes back to the driver.
static void host_trim() {
    const int64_t lim = g_host_cache_limit.load(std::memory_order_relaxed);
    if (lim < 0) return;
    if ((int64_t)(g_host.bytes() + g_stage.bytes() + g_compact.bytes()) > lim) {
        g_host.release(); g_compact.release(); g_stage.release();
        cudaGetLastError();
    }
}
struct HostTrim { ~HostTrim() { host_trim(); } };

// CPUs next to the current device: /sys/bus/pci/devices/<bus id>/local_cpulist (the cores of the GPU's NUMA
// node) intersected with the process affinity.  Empty set = unknown (no sysfs, no NUMA): callers then leave
// the affinity alone.  P2V_HOST_NUMA=0 switches the binding off.
static bool gpu_local_cpus(cpu_set_t* out) {
    CPU_ZERO(out);
    const char* off = getenv("P2V_HOST_NUMA");
    if (off && off[0] == '0') return false;
    int dev = 0;
    char bus[32] = {};
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetPCIBusId(bus, sizeof(bus), dev) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    for (char* c = bus; *c; ++c) if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
    std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return false;
    char line[4096] = {};
    const bool got = fgets(line, sizeof(line), f) != nullptr;
    fclose(f);
    if (!got) return false;
    cpu_set_t have;
    if (sched_getaffinity(0, sizeof(have), &have) != 0) return false;
    int n = 0;
    for (const char* q = line; *q;) {                 // "0-15,32-47"
        while (*q == ',' || *q == ' ' || *q == '\n') ++q;
        if (!*q) break;
        char* end = nullptr;
        long a = strtol(q, &end, 10), b = a;
        if (end == q) break;
        q = end;
        if (*q == '-') { b = strtol(q + 1, &end, 10); q = end; }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c)
            if (c >= 0 && CPU_ISSET((int)c, &have)) { CPU_SET((int)c, out); ++n; }
    }
    return n > 0 && n < CPU_COUNT(&have);            // every CPU local = nothing to bind
}
// binds the calling thread (workers call it when they start); false = left alone
static bool bind_thread_near_gpu() {
    cpu_set_t set;
    if (!gpu_local_cpus(&set)) return false;
    return sched_setaffinity(0, sizeof(set), &set) == 0;
}

static int host_thread_count() {
    const char* e = getenv("P2V_HOST_THREADS");
    if (e && atoi(e) > 0) return atoi(e);
    cpu_set_t set;
    int n = 0;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = CPU_COUNT(&set);
    if (n <= 0) n = (int)std::thread::hardware_concurrency();
    const char* lw = getenv("LOCAL_WORLD_SIZE");         // one process per GPU (torchrun): share the cores
    if (lw && atoi(lw) > 1) n /= atoi(lw);
    if (n > 16) n = 16;
    return n < 2 ? 2 : n;
}
static bool host_compact_enabled() {
    const char* e = getenv("P2V_HOST_COMPACT");
    return !(e && e[0] == '0');
}

// (c1, c2) int32 pairs -> uint16 pairs on the device (labels < 65536): half the PCIe bytes again
__global__ void pack_pairs_u16_kernel(const int2* __restrict__ in, int64_t count, ushort2* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const int2 x = in[i];
        out[i] = make_ushort2((unsigned short)x.x, (unsigned short)x.y);
    }
}
static void widen_trees_u16(const uint16_t* src, long long* dst, int64_t ntrees, int64_t k) {
    for (int64_t t = 0; t < ntrees; ++t) {
        for (int64_t r = 0; r < k; ++r) {
            _mm_stream_si64(dst, (long long)src[0]);
            _mm_stream_si64(dst + 1, (long long)src[1]);
            _mm_stream_si64(dst + 2, (long long)(k + 1 + r));
            src += 2; dst += 3;
        }
    }
}

// whole trees: (c1, c2) int32 -> (c1, c2, k+1+row) int64, non-temporal stores
static void widen_trees(const int32_t* src, long long* dst, int64_t ntrees, int64_t k) {
    for (int64_t t = 0; t < ntrees; ++t) {
        for (int64_t r = 0; r < k; ++r) {
            _mm_stream_si64(dst, (long long)src[0]);
            _mm_stream_si64(dst + 1, (long long)src[1]);
            _mm_stream_si64(dst + 2, (long long)(k + 1 + r));
            src += 2; dst += 3;
        }
    }
}

// kind 0: widen `ntrees` trees from the wire format into dst; kind 1: narrow `ntrees * k` int64 inputs at
// src into int32 at dst (values that do not fit become -1 and fail check_v on the device)
struct WidenJob { const void* src; void* dst; int64_t ntrees; int chunk; int kind; };

static void narrow_inputs(const long long* src, int32_t* dst, int64_t count) {
    for (int64_t i = 0; i < count; ++i) {
        const long long x = src[i];
        dst[i] = (x < 0 || x > 0x7FFFFFFFLL) ? -1 : (int32_t)x;
    }
}
static bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

static int ancestry_host_compact(const void* v, int64_t B, int64_t k, void* out, int32_t* status) {
    HostCtx& H = g_host;
    CompactCtx& C = g_compact;
    // chunk: ~48 K trees of n = 1000 (v 8 B + v32 4 B + pairs 8 B per element on the device)
    int64_t chunk = (int64_t)(((size_t)960 << 20) / ((size_t)k * 20));
    if (chunk < 1) chunk = 1;
    if (chunk > B) chunk = B;
    const bool u16 = 2 * k + 1 <= 65535;                 // labels fit 16 bits: 4 bytes per row on the wire
    const size_t wire_row = u16 ? 4 : 8;
    // device per element: v 8 B, v32 4 B (reused for the packed pairs), int32 pairs 8 B
    const size_t in_b = (size_t)k * 8 * chunk, out_b = (size_t)k * 8 * chunk, v32_b = (size_t)k * 4 * chunk;
    size_t wsb = decode_workspace_bytes(chunk, k);
    if (B % chunk != 0) { const size_t wt = decode_workspace_bytes(B % chunk, k); if (wt > wsb) wsb = wt; }
    // a pageable input (a plain numpy array) would be staged by the driver on one thread: the workers
    // narrow it into pinned int32 staging instead, and that goes to the device
    const bool pageable_in = is_pageable(v);
    cudaError_t e = host_ensure(H, pageable_in ? 0 : in_b, out_b, wsb, sizeof(int32_t) * chunk);
    if (e != cudaSuccess) { H.release(); return from_cuda(e, "host entry: device allocation"); }
    for (int i = 0; i < HOST_STREAMS && e == cudaSuccess; ++i) {
        if (out_b > C.cap_stage || !C.stage[i]) {
            if (C.stage[i]) cudaFreeHost(C.stage[i]);
            C.stage[i] = nullptr;
            e = cudaHostAlloc(&C.stage[i], out_b, cudaHostAllocDefault);
        }
        if (e == cudaSuccess && (sizeof(int32_t) * (size_t)chunk > C.cap_hstat || !C.hstat[i])) {
            if (C.hstat[i]) cudaFreeHost(C.hstat[i]);
            C.hstat[i] = nullptr;
            e = cudaHostAlloc((void**)&C.hstat[i], sizeof(int32_t) * (size_t)chunk, cudaHostAllocDefault);
        }
        if (e == cudaSuccess && (v32_b > C.cap_v32 || !C.v32[i])) {
            if (C.v32[i]) cudaFree(C.v32[i]);
            C.v32[i] = nullptr;
            e = cudaMalloc((void**)&C.v32[i], v32_b);
        }
        if (e == cudaSuccess && pageable_in && (v32_b > C.cap_vstage || !C.vstage[i])) {
            if (C.vstage[i]) cudaFreeHost(C.vstage[i]);
            C.vstage[i] = nullptr;
            e = cudaHostAlloc((void**)&C.vstage[i], v32_b, cudaHostAllocDefault);
        }
        if (e == cudaSuccess && !C.ev[i]) e = cudaEventCreateWithFlags(&C.ev[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) { C.release(); return from_cuda(e, "host entry: staging allocation"); }
    if (pageable_in && v32_b > C.cap_vstage) C.cap_vstage = v32_b;
    if (out_b > C.cap_stage) C.cap_stage = out_b;
    if (sizeof(int32_t) * (size_t)chunk > C.cap_hstat) C.cap_hstat = sizeof(int32_t) * (size_t)chunk;
    if (v32_b > C.cap_v32) C.cap_v32 = v32_b;

    // widening workers
    const int nthreads = host_thread_count();
    int dev_of_call = 0;
    cudaGetDevice(&dev_of_call);
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::deque<WidenJob> jobs;
    const int64_t nchunks = (B + chunk - 1) / chunk;
    std::vector<int> pending((size_t)nchunks, 0);      // slices of a chunk not yet widened
    std::vector<int> pending_in((size_t)nchunks, 0);   // slices of an input chunk not yet narrowed
    bool stop = false;
    std::vector<std::thread> pool;
    pool.reserve(nthreads);
    for (int t = 0; t < nthreads; ++t) {
        pool.emplace_back([&, dev_of_call]() {
            cudaSetDevice(dev_of_call);
            bind_thread_near_gpu();
            for (;;) {
                WidenJob j;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv_job.wait(lk, [&] { return stop || !jobs.empty(); });
                    if (jobs.empty()) return;
                    j = jobs.front(); jobs.pop_front();
                }
                if (j.kind == 1) narrow_inputs(static_cast<const long long*>(j.src), static_cast<int32_t*>(j.dst), j.ntrees * k);
                else if (u16) widen_trees_u16(static_cast<const uint16_t*>(j.src), static_cast<long long*>(j.dst), j.ntrees, k);
                else widen_trees(static_cast<const int32_t*>(j.src), static_cast<long long*>(j.dst), j.ntrees, k);
                _mm_sfence();
                {
                    std::lock_guard<std::mutex> lk(mu);
                    int& left = (j.kind == 1) ? pending_in[(size_t)j.chunk] : pending[(size_t)j.chunk];
                    if (--left == 0) cv_done.notify_all();
                }
            }
        });
    }
    auto enqueue = [&](int64_t c) {     // chunk c has landed in its staging buffer
        const int s = (int)(c % HOST_STREAMS);
        const int64_t b0 = c * chunk, nb = (B - b0 < chunk) ? (B - b0) : chunk;
        // slices of whole trees so that the row number inside a tree starts at 0 in every slice
        const int64_t per = (nb + nthreads - 1) / nthreads;
        std::lock_guard<std::mutex> lk(mu);
        for (int64_t t0 = 0; t0 < nb; t0 += per) {
            const int64_t nt = (nb - t0 < per) ? (nb - t0) : per;
            jobs.push_back(WidenJob{static_cast<const char*>(C.stage[s]) + (size_t)t0 * k * wire_row,
                                    static_cast<long long*>(out) + (size_t)(b0 + t0) * k * 3, nt, (int)c, 0});
            ++pending[(size_t)c];
        }
        cv_job.notify_all();
    };
    auto enqueue_narrow = [&](int64_t c) {     // input chunk c: caller's int64 -> pinned int32 staging
        const int s = (int)(c % HOST_STREAMS);
        const int64_t b0 = c * chunk, nb = (B - b0 < chunk) ? (B - b0) : chunk;
        const int64_t per = (nb + nthreads - 1) / nthreads;
        std::lock_guard<std::mutex> lk(mu);
        for (int64_t t0 = 0; t0 < nb; t0 += per) {
            const int64_t nt = (nb - t0 < per) ? (nb - t0) : per;
            jobs.push_back(WidenJob{static_cast<const long long*>(v) + (size_t)(b0 + t0) * k,
                                    C.vstage[s] + (size_t)t0 * k, nt, (int)c, 1});
            ++pending_in[(size_t)c];
        }
        cv_job.notify_all();
    };
    auto wait_narrowed = [&](int64_t c) {
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending_in[(size_t)c] == 0; });
    };
    auto wait_done = [&](int64_t c) {
        if (c < 0) return;
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending[(size_t)c] == 0; });
    };
    auto finish = [&]() {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv_job.notify_all();
        for (auto& th : pool) th.join();
    };

    int rc = P2V_OK;
    if (pageable_in) enqueue_narrow(0);
    for (int64_t c = 0; c < nchunks; ++c) {
        const int s = (int)(c % HOST_STREAMS);
        const int64_t b0 = c * chunk, nb = (B - b0 < chunk) ? (B - b0) : chunk;
        if (c >= HOST_STREAMS) {                         // staging buffer s is free once chunk c - 3 is widened
            // (its jobs were enqueued two iterations ago)
            wait_done(c - HOST_STREAMS);
        }
        if (pageable_in) {
            wait_narrowed(c);
            // staging (c + 1) % 3 was last read by the H2D of chunk c - 2, whose event was waited for
            // in the previous iteration
            if (c + 1 < nchunks) enqueue_narrow(c + 1);
            e = cudaMemcpyAsync(C.v32[s], C.vstage[s], (size_t)k * 4 * nb, cudaMemcpyHostToDevice, H.st[s]);
        } else {
            e = cudaMemcpyAsync(H.din[s], (const char*)v + (size_t)b0 * k * 8, (size_t)k * 8 * nb, cudaMemcpyHostToDevice, H.st[s]);
            if (e == cudaSuccess) e = launch_narrow_v(static_cast<const long long*>(H.din[s]), nb * k, C.v32[s], H.st[s]);
        }
        if (e == cudaSuccess) e = launch_decode(MODE_ANC2, C.v32[s], 4, nb, k, H.dout[s], H.dstat[s], H.dws[s], wsb, H.st[s], 0);
        const void* wire = H.dout[s];
        if (e == cudaSuccess && u16) {                  // the int32 v is dead after the decode: its buffer takes the packed pairs
            pack_pairs_u16_kernel<<<sm_count() * 8, 256, 0, H.st[s]>>>(static_cast<const int2*>(H.dout[s]), nb * k,
                                                                       reinterpret_cast<ushort2*>(C.v32[s]));
            count_launch();
            e = cudaGetLastError();
            wire = C.v32[s];
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(C.stage[s], wire, (size_t)k * wire_row * nb, cudaMemcpyDeviceToHost, H.st[s]);
        if (e == cudaSuccess && status)
            e = cudaMemcpyAsync(C.hstat[s], H.dstat[s], sizeof(int32_t) * nb, cudaMemcpyDeviceToHost, H.st[s]);
        if (e == cudaSuccess) e = cudaEventRecord(C.ev[s], H.st[s]);
        if (e == cudaSuccess && c >= 1) {               // the previous chunk: wait for it, hand it to the workers
            e = cudaEventSynchronize(C.ev[(c - 1) % HOST_STREAMS]);
            if (e == cudaSuccess) {
                if (status) memcpy(status + (c - 1) * chunk, C.hstat[(c - 1) % HOST_STREAMS], sizeof(int32_t) * chunk);
                enqueue(c - 1);
            }
        }
        if (e != cudaSuccess) { rc = from_cuda(e, "host entry: chunk"); break; }
    }
    if (rc == P2V_OK) {
        e = cudaEventSynchronize(C.ev[(nchunks - 1) % HOST_STREAMS]);
        if (e == cudaSuccess) {
            const int64_t b0 = (nchunks - 1) * chunk;
            if (status) memcpy(status + b0, C.hstat[(nchunks - 1) % HOST_STREAMS], sizeof(int32_t) * (B - b0));
            enqueue(nchunks - 1);
        } else rc = from_cuda(e, "host entry: synchronize");
    }
    if (rc == P2V_OK) for (int64_t c = (nchunks > HOST_STREAMS ? nchunks - HOST_STREAMS : 0); c < nchunks; ++c) wait_done(c);
    finish();
    for (int i = 0; i < HOST_STREAMS; ++i) cudaStreamSynchronize(H.st[i]);
    return rc;
}

static int encode_host(int mode, int cols, const void* in, int idx_bytes, int64_t B, int64_t k, void* v,
                       int32_t* status) {
    int rc = check_common(in, v, idx_bytes, B, k);
    if (rc) return rc;
    if (k == 0) { if (status && B > 0) memset(status, 0, sizeof(int32_t) * B); return P2V_OK; }
    return host_chunked(
        in, (size_t)k * cols * idx_bytes, v, (size_t)k * idx_bytes, status, B,
        [&](int64_t nb) { return encode_workspace_bytes(nb, k); },
        [&](void* din, void* dout, int32_t* dst, int64_t nb, void* ws, size_t wsb, cudaStream_t st) {
            return launch_encode(mode, din, idx_bytes, nb, k, dout, dst, ws, wsb, st);
        });
}

static int coph_check(const void* in, const double* out, int64_t B, int64_t k, int64_t rb, int64_t re) {
    if (B < 0 || k < 0) return fail(P2V_ERR_INVALID_ARG, "negative batch or length");
    if (rb < 0 || re > k + 1 || rb > re) return fail(P2V_ERR_INVALID_ARG, "row block outside [0, n]");
    if (B > 0 && ((re > rb && !out) || (k > 0 && !in))) return fail(P2V_ERR_INVALID_ARG, "null buffer");
    if (k > LARGE_K_MAX) return fail(P2V_ERR_UNSUPPORTED, "k above the supported maximum (2^19)");
    return P2V_OK;
}

// Host entry for distances.  in_a: v (idx_bytes 4/8) or the (k,3) matrix (idx_bytes 0); bls optional.
// what: 0 distances, 1 variance-covariance, 2 node depths (out (B, 2k+1), rb/re ignored)
static int coph_host(const void* in_a, int idx_bytes, const double* bls, int64_t B, int64_t k, int unrooted,
                     int64_t rb, int64_t re, double* out, int32_t* status, int what = 0) {
    int rc = coph_check(in_a, out, B, k, rb, re);
    if (rc) return rc;
    if (B == 0 || (re == rb && what != 2)) return P2V_OK;
    const size_t a_bytes = (idx_bytes == 0) ? (size_t)k * 3 * sizeof(double) : (size_t)k * idx_bytes;
    const size_t b_bytes = bls ? (size_t)k * 2 * sizeof(double) : 0;
    const size_t out_tree = (what == 2) ? (size_t)(2 * k + 1) * sizeof(double)
                                        : (size_t)(re - rb) * (size_t)(k + 1) * sizeof(double);
    auto run = [&](void* din, const double* dbls, void* dout, int32_t* dst, int64_t nb, void* ws, size_t wsb,
                   cudaStream_t st) {
        if (what == 2) return launch_node_depths(din, idx_bytes, dbls, nb, k, (double*)dout, dst, ws, wsb, st);
        return launch_cophenetic(din, idx_bytes, dbls, nb, k, unrooted, rb, re, (double*)dout, dst, ws, wsb, st, what);
    };
    // one packed input record per tree: [a | bls]
    const size_t in_tree = a_bytes + b_bytes;
    // pack on the fly only when bls is present and separate; otherwise pass through
    if (!bls) {
        return host_chunked(
            in_a, a_bytes, out, out_tree, status, B,
            [&](int64_t nb) { return cophenetic_workspace_bytes(nb, k); },
            [&](void* din, void* dout, int32_t* dst, int64_t nb, void* ws, size_t wsb, cudaStream_t st) {
                return run(din, nullptr, dout, dst, nb, ws, wsb, st);
            });
    }
    // v and bls are separate host arrays: copy bls per chunk inside the launch lambda
    (void)in_tree;
    int64_t done = 0;  // trees consumed so far (chunks are issued in order)
    double* dbl[HOST_STREAMS] = {};
    int flip = 0;
    int64_t cap_chunk = 0;
    auto cleanup = [&]() { for (int i = 0; i < HOST_STREAMS; ++i) if (dbl[i]) cudaFree(dbl[i]); };
    int r = host_chunked(
        in_a, a_bytes, out, out_tree, status, B,
        [&](int64_t nb) { cap_chunk = nb; return cophenetic_workspace_bytes(nb, k); },
        [&](void* din, void* dout, int32_t* dst, int64_t nb, void* ws, size_t wsb, cudaStream_t st) {
            cudaError_t e = cudaSuccess;
            const int s = flip; flip = (flip + 1) % HOST_STREAMS;
            if (!dbl[s]) e = cudaMalloc((void**)&dbl[s], b_bytes * (size_t)cap_chunk);
            if (e != cudaSuccess) return e;
            e = cudaMemcpyAsync(dbl[s], (const char*)bls + (size_t)done * b_bytes, b_bytes * nb,
                                cudaMemcpyHostToDevice, st);
            done += nb;
            if (e != cudaSuccess) return e;
            return run(din, dbl[s], dout, dst, nb, ws, wsb, st);
        });
    cudaDeviceSynchronize();
    cleanup();
    return r;
}

}  // namespace p2v

using namespace p2v;

extern "C" {

int p2v_version(void) { return 100; }
const char* p2v_last_error(void) { return g_err.c_str(); }
int64_t p2v_launch_count(void) { return g_launches.load(); }

size_t p2v_decode_workspace_bytes(int64_t B, int64_t k) { return decode_workspace_bytes(B, k); }
size_t p2v_encode_workspace_bytes(int64_t B, int64_t k) { return encode_workspace_bytes(B, k); }
size_t p2v_cophenetic_workspace_bytes(int64_t B, int64_t k) { return cophenetic_workspace_bytes(B, k); }

int p2v_to_pairs_batch(const void* v, int idx_bytes, int64_t B, int64_t k, void* pairs, int32_t* status,
                       void* ws, size_t ws_bytes, p2v_stream_t stream) {
    return decode_entry(MODE_PAIRS, v, idx_bytes, B, k, pairs, status, ws, ws_bytes, stream);
}
int p2v_to_ancestry_batch(const void* v, int idx_bytes, int64_t B, int64_t k, void* anc, int32_t* status,
                          void* ws, size_t ws_bytes, p2v_stream_t stream) {
    return decode_entry(MODE_ANCESTRY, v, idx_bytes, B, k, anc, status, ws, ws_bytes, stream);
}
int p2v_to_edges_batch(const void* v, int idx_bytes, int64_t B, int64_t k, void* edges, int32_t* status,
                       void* ws, size_t ws_bytes, p2v_stream_t stream) {
    return decode_entry(MODE_EDGES, v, idx_bytes, B, k, edges, status, ws, ws_bytes, stream);
}
int p2v_from_pairs_batch(const void* pairs, int idx_bytes, int64_t B, int64_t k, void* v, int32_t* status,
                         void* ws, size_t ws_bytes, p2v_stream_t stream) {
    return encode_entry(FROM_PAIRS, pairs, idx_bytes, B, k, v, status, ws, ws_bytes, stream);
}
int p2v_from_ancestry_batch(const void* anc, int idx_bytes, int64_t B, int64_t k, void* v, int32_t* status,
                            void* ws, size_t ws_bytes, p2v_stream_t stream) {
    return encode_entry(FROM_ANCESTRY, anc, idx_bytes, B, k, v, status, ws, ws_bytes, stream);
}
int p2v_from_edges_batch(const void* edges, int idx_bytes, int64_t B, int64_t k, void* v, int32_t* status,
                         void* ws, size_t ws_bytes, p2v_stream_t stream) {
    return encode_entry(FROM_EDGES, edges, idx_bytes, B, k, v, status, ws, ws_bytes, stream);
}

int p2v_check_v_batch(const void* v, int idx_bytes, int64_t B, int64_t k, int32_t* status, p2v_stream_t stream) {
    if (!status) return fail(P2V_ERR_INVALID_ARG, "status is required");
    int rc = check_common(v, status, idx_bytes, B, k);
    if (rc) return rc;
    if (k == 0) return from_cuda(cudaMemsetAsync(status, 0, sizeof(int32_t) * B, (cudaStream_t)stream), "memset");
    return from_cuda(launch_check_v(v, idx_bytes, B, k, status, (cudaStream_t)stream), "check_v launch");
}

int p2v_cophenetic_batch(const void* v, int idx_bytes, const double* bls, int64_t B, int64_t k, int unrooted,
                         int64_t row_begin, int64_t row_end, double* out, int32_t* status, void* ws,
                         size_t ws_bytes, p2v_stream_t stream) {
    if (idx_bytes != 4 && idx_bytes != 8) return fail(P2V_ERR_INVALID_ARG, "idx_bytes must be 4 or 8");
    int rc = coph_check(v, out, B, k, row_begin, row_end);
    if (rc) return rc;
    if (ws_bytes < cophenetic_workspace_bytes(B, k)) return fail(P2V_ERR_WORKSPACE, "cophenetic workspace too small");
    return from_cuda(launch_cophenetic(v, idx_bytes, bls, B, k, unrooted, row_begin, row_end, out, status, ws,
                                       ws_bytes, (cudaStream_t)stream), "cophenetic launch");
}

int p2v_cophenetic_matrix_batch(const double* m, int64_t B, int64_t k, int unrooted, int64_t row_begin,
                                int64_t row_end, double* out, int32_t* status, void* ws, size_t ws_bytes,
                                p2v_stream_t stream) {
    int rc = coph_check(m, out, B, k, row_begin, row_end);
    if (rc) return rc;
    if (ws_bytes < cophenetic_workspace_bytes(B, k)) return fail(P2V_ERR_WORKSPACE, "cophenetic workspace too small");
    return from_cuda(launch_cophenetic(m, 0, nullptr, B, k, unrooted, row_begin, row_end, out, status, ws,
                                       ws_bytes, (cudaStream_t)stream), "cophenetic launch");
}

int p2v_cophenetic_prepare_batch(const void* v_or_matrix, int idx_bytes, const double* bls, int64_t B, int64_t k,
                                 int unrooted, int32_t* status, void* ws, size_t ws_bytes, p2v_stream_t stream) {
    if (idx_bytes != 0 && idx_bytes != 4 && idx_bytes != 8) return fail(P2V_ERR_INVALID_ARG, "idx_bytes must be 0, 4 or 8");
    if (B < 0 || k < 0) return fail(P2V_ERR_INVALID_ARG, "negative batch or length");
    if (B > 0 && k > 0 && !v_or_matrix) return fail(P2V_ERR_INVALID_ARG, "null buffer");
    if (k > LARGE_K_MAX) return fail(P2V_ERR_UNSUPPORTED, "k above the supported maximum (2^19)");
    if (ws_bytes < cophenetic_workspace_bytes(B, k)) return fail(P2V_ERR_WORKSPACE, "cophenetic workspace too small");
    return from_cuda(launch_cophenetic_prepare(v_or_matrix, idx_bytes, bls, B, k, unrooted, status, ws, ws_bytes,
                                               (cudaStream_t)stream), "cophenetic prepare");
}
int p2v_cophenetic_rows_batch(void* ws, size_t ws_bytes, int64_t B, int64_t k, int weighted,
                              int64_t row_begin, int64_t row_end, double* out, p2v_stream_t stream) {
    int rc = coph_check(ws, out, B, k, row_begin, row_end);
    if (rc) return rc;
    if (ws_bytes < cophenetic_workspace_bytes(B, k)) return fail(P2V_ERR_WORKSPACE, "cophenetic workspace too small");
    return from_cuda(launch_cophenetic_rows(ws, ws_bytes, B, k, weighted, row_begin, row_end, out,
                                            (cudaStream_t)stream), "cophenetic rows");
}

// ---- variance-covariance and node depths -----------------------------------------------
static int vcv_check(int64_t k) {
    if (k < 1) return fail(P2V_ERR_INVALID_ARG, "Variance-covariance matrix not supported for trees with < 2 leaves.");
    return P2V_OK;
}
int p2v_vcv_batch(const void* v, int idx_bytes, const double* bls, int64_t B, int64_t k, int64_t row_begin,
                  int64_t row_end, double* out, int32_t* status, void* ws, size_t ws_bytes, p2v_stream_t stream) {
    if (idx_bytes != 4 && idx_bytes != 8) return fail(P2V_ERR_INVALID_ARG, "idx_bytes must be 4 or 8");
    int rc = coph_check(v, out, B, k, row_begin, row_end);
    if (!rc) rc = vcv_check(k);
    if (rc) return rc;
    if (ws_bytes < cophenetic_workspace_bytes(B, k)) return fail(P2V_ERR_WORKSPACE, "cophenetic workspace too small");
    return from_cuda(launch_cophenetic(v, idx_bytes, bls, B, k, 0, row_begin, row_end, out, status, ws, ws_bytes,
                                       (cudaStream_t)stream, 1), "vcv launch");
}
int p2v_vcv_matrix_batch(const double* m, int64_t B, int64_t k, int64_t row_begin, int64_t row_end, double* out,
                         int32_t* status, void* ws, size_t ws_bytes, p2v_stream_t stream) {
    int rc = coph_check(m, out, B, k, row_begin, row_end);
    if (!rc) rc = vcv_check(k);
    if (rc) return rc;
    if (ws_bytes < cophenetic_workspace_bytes(B, k)) return fail(P2V_ERR_WORKSPACE, "cophenetic workspace too small");
    return from_cuda(launch_cophenetic(m, 0, nullptr, B, k, 0, row_begin, row_end, out, status, ws, ws_bytes,
                                       (cudaStream_t)stream, 1), "vcv launch");
}
int p2v_vcv_rows_batch(void* ws, size_t ws_bytes, int64_t B, int64_t k, int weighted, int64_t row_begin,
                       int64_t row_end, double* out, p2v_stream_t stream) {
    int rc = coph_check(ws, out, B, k, row_begin, row_end);
    if (!rc) rc = vcv_check(k);
    if (rc) return rc;
    if (ws_bytes < cophenetic_workspace_bytes(B, k)) return fail(P2V_ERR_WORKSPACE, "cophenetic workspace too small");
    return from_cuda(launch_cophenetic_rows(ws, ws_bytes, B, k, weighted, row_begin, row_end, out,
                                            (cudaStream_t)stream, 1), "vcv rows");
}
int p2v_node_depths_batch(const void* v_or_matrix, int idx_bytes, const double* bls, int64_t B, int64_t k,
                          double* out, int32_t* status, void* ws, size_t ws_bytes, p2v_stream_t stream) {
    if (idx_bytes != 0 && idx_bytes != 4 && idx_bytes != 8) return fail(P2V_ERR_INVALID_ARG, "idx_bytes must be 0, 4 or 8");
    if (B < 0 || k < 0) return fail(P2V_ERR_INVALID_ARG, "negative batch or length");
    if (B > 0 && (!out || (k > 0 && !v_or_matrix))) return fail(P2V_ERR_INVALID_ARG, "null buffer");
    if (k > LARGE_K_MAX) return fail(P2V_ERR_UNSUPPORTED, "k above the supported maximum (2^19)");
    if (ws_bytes < cophenetic_workspace_bytes(B, k)) return fail(P2V_ERR_WORKSPACE, "cophenetic workspace too small");
    return from_cuda(launch_node_depths(v_or_matrix, idx_bytes, bls, B, k, out, status, ws, ws_bytes,
                                        (cudaStream_t)stream), "node depths launch");
}
int p2v_vcv_host(const void* v, int idx_bytes, const double* bls, int64_t B, int64_t k, int64_t row_begin,
                 int64_t row_end, double* out, int32_t* status) {
    HostTrim trim_on_return;
    if (idx_bytes != 4 && idx_bytes != 8) return fail(P2V_ERR_INVALID_ARG, "idx_bytes must be 4 or 8");
    int rc = vcv_check(k);
    if (rc) return rc;
    return coph_host(v, idx_bytes, bls, B, k, 0, row_begin, row_end, out, status, 1);
}
int p2v_vcv_matrix_host(const double* m, int64_t B, int64_t k, int64_t row_begin, int64_t row_end, double* out,
                        int32_t* status) {
    HostTrim trim_on_return;
    int rc = vcv_check(k);
    if (rc) return rc;
    return coph_host(m, 0, nullptr, B, k, 0, row_begin, row_end, out, status, 1);
}
int p2v_node_depths_host(const void* v_or_matrix, int idx_bytes, const double* bls, int64_t B, int64_t k,
                         double* out, int32_t* status) {
    HostTrim trim_on_return;
    if (idx_bytes != 0 && idx_bytes != 4 && idx_bytes != 8) return fail(P2V_ERR_INVALID_ARG, "idx_bytes must be 0, 4 or 8");
    return coph_host(v_or_matrix, idx_bytes, bls, B, k, 0, 0, k + 1, out, status, 2);
}

// ---- balance indices (third "next" row, first half): Sackin, leaf-depth variance, B2 -----------
// vector/balance.rs:14-56: all three are sums over the topological depths of the leaves
// (get_node_depths), so they run on the prepared tables of the distance path: node depths into the
// workspace, then one warp per tree.  sum d and sum d^2 are integers (exact in fp64 below 2^53, in
// any order); B2 = sum d * 2^-d is a sum of positive terms added pairwise.
__global__ void balance_kernel(const double* __restrict__ depths, int64_t B, int k, double* __restrict__ out,
                               const int32_t* __restrict__ status) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int N = 2 * k + 1, n = k + 1;
    for (int64_t t = warp; t < B; t += nwarps) {
        if (status[t] != 0) {
            if (lane < 3) out[3 * t + lane] = 0.0;
            continue;
        }
        const double* d = depths + t * N;
        double s = 0.0, s2 = 0.0, b2 = 0.0;
        for (int i = lane; i < n; i += 32) {
            const double x = d[i];
            s += x; s2 += x * x;
            b2 += x * ldexp(1.0, -(int)x);        // 2f64.powi(-d): 0 once it underflows, like the reference
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
            s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o);
            b2 += __shfl_xor_sync(0xFFFFFFFFu, b2, o);
        }
        if (lane == 0) {
            const double nn = (double)n, m = __ddiv_rn(s, nn);
            out[3 * t] = s;                       // sackin (balance.rs:14-18)
            out[3 * t + 1] = __dsub_rn(__ddiv_rn(s2, nn), __dmul_rn(m, m));     // leaf_depth_variance (:20-33), no FMA like Rust
            out[3 * t + 2] = b2;                  // b2 (:49-58)
        }
    }
}
// 2 <= k <= 4095: one warp per tree on the two informative ancestry columns (MODE_ANC2, int32), all
// state in shared memory (8 bytes per leaf).  Every internal child records its parent row; the depth
// of the internal nodes comes top-down, 32 rows at a time from the root, parents inside the batch by
// pointer doubling (the sweep of the to_newick kernel with unit weights); a row then adds depth + 1
// for each of its leaf children.  The two integer sums are accumulated as integers.
constexpr int BAL_WARP_K_MAX = 4095;
static size_t bal_warp_bytes(int64_t k) { return (((size_t)8 * k) + 15) & ~(size_t)15; }
__global__ void __launch_bounds__(128)
balance_warp_kernel(const int* __restrict__ anc2, int64_t B, int k, double* __restrict__ out,
                    const int32_t* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char bal_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    unsigned char* mine = bal_dyn + (size_t)warp * ((((size_t)8 * k) + 15) & ~(size_t)15);
    uint32_t* ch = reinterpret_cast<uint32_t*>(mine);                       // c1 | c2 << 16 of row r
    uint16_t* par = reinterpret_cast<uint16_t*>(mine + 4 * (size_t)k);      // parent row of the node of row r
    uint16_t* dep = par + k;                                                 // depth of the node of row r
    const int nb = (k + 31) >> 5;
    for (int64_t t = (int64_t)blockIdx.x * nwarps + warp; t < B; t += (int64_t)gridDim.x * nwarps) {
        if (status[t] != 0) {
            if (lane < 3) out[3 * t + lane] = 0.0;
            continue;
        }
        const int2* anc = reinterpret_cast<const int2*>(anc2 + (size_t)t * 2 * k);
        for (int r = lane; r < k; r += 32) {
            const int2 x = anc[r];
            ch[r] = (uint32_t)x.x | ((uint32_t)x.y << 16);
            if (x.x > k) par[x.x - k - 1] = (uint16_t)r;
            if (x.y > k) par[x.y - k - 1] = (uint16_t)r;
        }
        __syncwarp();
        unsigned long long s = 0, s2 = 0;
        double b2 = 0.0;
        for (int b = nb - 1; b >= 0; --b) {
            const int base = b * 32, r = base + lane;
            const bool act = r < k;
            uint32_t acc = 0;
            int ptr = lane;
            bool fin = true;
            if (act && r != k - 1) {
                const int pr = (int)par[r];
                acc = 1;
                if (pr >= base + 32) acc += dep[pr];
                else { ptr = pr - base; fin = false; }
            }
#pragma unroll
            for (int it = 0; it < 5; ++it) {
                const uint32_t a = __shfl_sync(0xFFFFFFFFu, acc | (fin ? 0x80000000u : 0u), ptr);
                const int pp = __shfl_sync(0xFFFFFFFFu, ptr, ptr);
                if (!fin) { acc += a & 0x7FFFFFFFu; fin = (a >> 31) != 0u; ptr = pp; }
            }
            if (act) {
                dep[r] = (uint16_t)acc;
                const uint32_t cc = ch[r];
                const int leaves = ((int)(cc & 0xFFFFu) <= k) + ((int)(cc >> 16) <= k);
                const unsigned long long d = acc + 1u;                      // depth of this row's leaf children
                s += (unsigned long long)leaves * d;
                s2 += (unsigned long long)leaves * d * d;
                b2 += (double)leaves * ((double)d * ldexp(1.0, -(int)d));
            }
            __syncwarp();
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
            s2 += __shfl_xor_sync(0xFFFFFFFFu, s2, o);
            b2 += __shfl_xor_sync(0xFFFFFFFFu, b2, o);
        }
        if (lane == 0) {
            const double nn = (double)(k + 1), sd = (double)s, m = __ddiv_rn(sd, nn);
            out[3 * t] = sd;
            out[3 * t + 1] = __dsub_rn(__ddiv_rn((double)s2, nn), __dmul_rn(m, m));
            out[3 * t + 2] = b2;
        }
    }
}

static size_t bal_al(size_t x) { return (x + 255) & ~(size_t)255; }
static bool bal_warp_path(int64_t k) { return k >= 2 && k <= BAL_WARP_K_MAX; }
size_t p2v_balance_workspace_bytes(int64_t B, int64_t k) {
    if (B <= 0 || k < 0) return 0;
    if (bal_warp_path(k))      // [ status | int32 v | int32 (k,2) ancestry columns | decode scratch ]
        return bal_al(sizeof(int32_t) * (size_t)B) + bal_al(4 * (size_t)B * (size_t)k) + bal_al(8 * (size_t)B * (size_t)k) +
               bal_al(decode_workspace_bytes(B, k));
    return bal_al(cophenetic_workspace_bytes(B, k)) + bal_al(sizeof(double) * (size_t)B * (size_t)(2 * k + 1)) +
           bal_al(sizeof(int32_t) * (size_t)B);
}
static cudaError_t launch_balance(const void* v, int idx_bytes, int64_t B, int64_t k, double* out, int32_t* status,
                                  void* ws, size_t ws_bytes, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    unsigned char* wsb = static_cast<unsigned char*>(ws);
    if (bal_warp_path(k)) {
        int32_t* dst = reinterpret_cast<int32_t*>(wsb);
        int* v32 = reinterpret_cast<int*>(wsb + bal_al(sizeof(int32_t) * (size_t)B));
        int* anc2 = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(v32) + bal_al(4 * (size_t)B * (size_t)k));
        unsigned char* dws = reinterpret_cast<unsigned char*>(anc2) + bal_al(8 * (size_t)B * (size_t)k);
        const size_t dwb = decode_workspace_bytes(B, k);
        const void* vin = v;
        cudaError_t e = cudaSuccess;
        if (idx_bytes == 8) {
            e = launch_narrow_v(static_cast<const long long*>(v), B * k, v32, st);
            if (e != cudaSuccess) return e;
            vin = v32;
        }
        e = launch_decode(MODE_ANC2, vin, 4, B, k, anc2, dst, dwb ? dws : nullptr, dwb, st, 0);
        if (e != cudaSuccess) return e;
        const size_t per = bal_warp_bytes(k);
        int wpc = 1, best = 0;
        for (int w = 4; w >= 1; w >>= 1) {
            size_t ctas = ((size_t)228 * 1024) / (per * w + 1024);
            if (ctas > 32) ctas = 32;
            if (ctas > (size_t)(64 / w)) ctas = 64 / w;
            if ((int)ctas * w > best) { best = (int)ctas * w; wpc = w; }
        }
        const size_t smem = per * wpc;
        e = cudaFuncSetAttribute(balance_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        const int64_t cap = (int64_t)sm_count() * (best / wpc), want = (B + wpc - 1) / wpc;
        balance_warp_kernel<<<(int)(want < cap ? want : cap), wpc * 32, smem, st>>>(anc2, B, (int)k, out, dst);
        count_launch();
        e = cudaGetLastError();
        if (e == cudaSuccess && status) e = cudaMemcpyAsync(status, dst, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, st);
        return e;
    }
    const size_t cw = bal_al(cophenetic_workspace_bytes(B, k));
    double* depths = reinterpret_cast<double*>(wsb + cw);
    int32_t* dst = reinterpret_cast<int32_t*>(wsb + cw + bal_al(sizeof(double) * (size_t)B * (size_t)(2 * k + 1)));
    cudaError_t e = cudaMemsetAsync(dst, 0, sizeof(int32_t) * B, st);
    if (e != cudaSuccess) return e;
    e = launch_node_depths(v, idx_bytes, nullptr, B, k, depths, dst, ws, cw, st);
    if (e != cudaSuccess) return e;
    const int64_t want = (B * 32 + 255) / 256, cap = (int64_t)sm_count() * 8;
    balance_kernel<<<(int)(want < cap ? want : cap), 256, 0, st>>>(depths, B, (int)k, out, dst);
    count_launch();
    e = cudaGetLastError();
    if (e == cudaSuccess && status) e = cudaMemcpyAsync(status, dst, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, st);
    return e;
}
static int balance_check(const void* v, const double* out, int idx_bytes, int64_t B, int64_t k) {
    if (idx_bytes != 4 && idx_bytes != 8) return fail(P2V_ERR_INVALID_ARG, "idx_bytes must be 4 or 8");
    if (B < 0 || k < 0) return fail(P2V_ERR_INVALID_ARG, "negative batch or length");
    if (k > LARGE_K_MAX) return fail(P2V_ERR_UNSUPPORTED, "k above the supported maximum (2^19)");
    if (B > 0 && (!out || (k > 0 && !v))) return fail(P2V_ERR_INVALID_ARG, "null buffer");
    return P2V_OK;
}
int p2v_balance_batch(const void* v, int idx_bytes, int64_t B, int64_t k, double* out, int32_t* status, void* ws,
                      size_t ws_bytes, p2v_stream_t stream) {
    int rc = balance_check(v, out, idx_bytes, B, k);
    if (rc) return rc;
    if (B == 0) return P2V_OK;
    if (!ws || ws_bytes < p2v_balance_workspace_bytes(B, k)) return fail(P2V_ERR_WORKSPACE, "balance workspace too small");
    return from_cuda(launch_balance(v, idx_bytes, B, k, out, status, ws, ws_bytes, (cudaStream_t)stream), "balance launch");
}
int p2v_balance_host(const void* v, int idx_bytes, int64_t B, int64_t k, double* out, int32_t* status) {
    HostTrim trim_on_return;
    int rc = balance_check(v, out, idx_bytes, B, k);
    if (rc) return rc;
    return host_chunked(
        v, (size_t)k * idx_bytes, out, 3 * sizeof(double), status, B,
        [&](int64_t nb) { return p2v_balance_workspace_bytes(nb, k); },
        [&](void* din, void* dout, int32_t* dst, int64_t nb, void* ws, size_t wsb, cudaStream_t st) {
            return launch_balance(din, idx_bytes, nb, k, (double*)dout, dst, ws, wsb, st);
        });
}

// ---- Robinson-Foulds (third "next" row, second half) -------------------------------------------
size_t p2v_robinson_foulds_workspace_bytes(int64_t B1, int64_t B2, int64_t k) {
    return robinson_foulds_workspace_bytes(B1, B2, k);
}
static int rf_check(const void* v1, const void* v2, const double* out, int idx_bytes, int64_t B1, int64_t B2, int64_t k) {
    if (idx_bytes != 4 && idx_bytes != 8) return fail(P2V_ERR_INVALID_ARG, "idx_bytes must be 4 or 8");
    if (B1 < 0 || B2 < 0 || k < 0) return fail(P2V_ERR_INVALID_ARG, "negative batch or length");
    if (B1 != B2 && B1 != 1 && B2 != 1) return fail(P2V_ERR_INVALID_ARG, "batches must have equal sizes, or one of them one tree");
    if (k > 65535) return fail(P2V_ERR_UNSUPPORTED, "robinson_foulds: more than 65536 leaves");
    if (B1 > 0 && B2 > 0 && (!out || (k > 0 && (!v1 || !v2)))) return fail(P2V_ERR_INVALID_ARG, "null buffer");
    return P2V_OK;
}
int p2v_robinson_foulds_batch(const void* v1, const void* v2, int idx_bytes, int64_t B1, int64_t B2, int64_t k,
                              int normalize, double* out, int32_t* status, void* ws, size_t ws_bytes,
                              p2v_stream_t stream) {
    int rc = rf_check(v1, v2, out, idx_bytes, B1, B2, k);
    if (rc) return rc;
    if (B1 == 0 || B2 == 0) return P2V_OK;
    if (ws_bytes < robinson_foulds_workspace_bytes(B1, B2, k) || (!ws && robinson_foulds_workspace_bytes(B1, B2, k) > 0))
        return fail(P2V_ERR_WORKSPACE, "robinson_foulds workspace too small");
    return from_cuda(launch_robinson_foulds(v1, v2, idx_bytes, B1, B2, k, normalize, out, status, ws, ws_bytes,
                                            (cudaStream_t)stream), "robinson_foulds launch");
}
int p2v_robinson_foulds_host(const void* v1, const void* v2, int idx_bytes, int64_t B1, int64_t B2, int64_t k,
                             int normalize, double* out, int32_t* status) {
    int rc = rf_check(v1, v2, out, idx_bytes, B1, B2, k);
    if (rc) return rc;
    if (B1 == 0 || B2 == 0) return P2V_OK;
    const int64_t P = B1 > B2 ? B1 : B2;
    const size_t tree_b = (size_t)(k > 0 ? k : 1) * idx_bytes;
    int64_t chunk = (int64_t)(((size_t)256 << 20) / tree_b);
    if (chunk < 1) chunk = 1;
    if (chunk > P) chunk = P;
    const int64_t c1 = B1 == 1 ? 1 : chunk, c2 = B2 == 1 ? 1 : chunk;
    const size_t wsb = robinson_foulds_workspace_bytes(c1, c2, k);
    void *d1 = nullptr, *d2 = nullptr, *dws = nullptr;
    double* dout = nullptr;
    int32_t* dst = nullptr;
    cudaError_t e = cudaMalloc(&d1, tree_b * c1);
    if (e == cudaSuccess) e = cudaMalloc(&d2, tree_b * c2);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dout, sizeof(double) * chunk);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dst, sizeof(int32_t) * chunk);
    if (e == cudaSuccess && wsb) e = cudaMalloc(&dws, wsb);
    for (int64_t p0 = 0; p0 < P && e == cudaSuccess; p0 += chunk) {
        const int64_t nb = (P - p0 < chunk) ? (P - p0) : chunk;
        const int64_t n1 = B1 == 1 ? 1 : nb, n2 = B2 == 1 ? 1 : nb;
        if (k > 0) {
            e = cudaMemcpy(d1, (const char*)v1 + (B1 == 1 ? 0 : (size_t)p0 * tree_b), tree_b * n1, cudaMemcpyHostToDevice);
            if (e == cudaSuccess)
                e = cudaMemcpy(d2, (const char*)v2 + (B2 == 1 ? 0 : (size_t)p0 * tree_b), tree_b * n2, cudaMemcpyHostToDevice);
        }
        if (e == cudaSuccess) e = launch_robinson_foulds(d1, d2, idx_bytes, n1, n2, k, normalize, dout, dst, dws, wsb, nullptr);
        if (e == cudaSuccess) e = cudaMemcpy(out + p0, dout, sizeof(double) * nb, cudaMemcpyDeviceToHost);
        if (e == cudaSuccess && status) e = cudaMemcpy(status + p0, dst, sizeof(int32_t) * nb, cudaMemcpyDeviceToHost);
    }
    cudaFree(d1); cudaFree(d2); cudaFree(dout); cudaFree(dst); cudaFree(dws);
    return from_cuda(e, "robinson_foulds host entry");
}

// ---- to_newick --------------------------------------------------------------------------
size_t p2v_newick_max_bytes(int64_t k) { return newick_max_bytes(k); }
size_t p2v_newick_workspace_bytes(int64_t B, int64_t k) { return newick_workspace_bytes(B, k); }
static int newick_check(const void* v, const char* out, int idx_bytes, int64_t B, int64_t k, int64_t out_stride) {
    if (idx_bytes != 4 && idx_bytes != 8) return fail(P2V_ERR_INVALID_ARG, "idx_bytes must be 4 or 8");
    if (B < 0 || k < 0) return fail(P2V_ERR_INVALID_ARG, "negative batch or length");
    if (k > LARGE_K_MAX) return fail(P2V_ERR_UNSUPPORTED, "to_newick: k above 2^19");
    if (B > 0 && (!out || (k > 0 && !v))) return fail(P2V_ERR_INVALID_ARG, "null buffer");
    if (out_stride < (int64_t)newick_max_bytes(k)) return fail(P2V_ERR_INVALID_ARG, "out_stride below p2v_newick_max_bytes(k)");
    return P2V_OK;
}
__global__ void newick_k0_kernel(char* out, int64_t out_stride, int64_t B, int64_t* lengths, int32_t* status) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < B; t += (int64_t)gridDim.x * blockDim.x) {
        char* o = out + t * out_stride;          // a single leaf: "0;" (convert.rs:361-381 with no pairs)
        o[0] = '0'; o[1] = ';'; o[2] = 0;
        if (lengths) lengths[t] = 2;
        if (status) status[t] = 0;
    }
}
static cudaError_t newick_launch_any(const void* v, int idx_bytes, int64_t B, int64_t k, char* out, int64_t out_stride,
                                     int64_t* lengths, int32_t* status, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    if (k == 0) {
        newick_k0_kernel<<<(int)((B + 255) / 256 < 1024 ? (B + 255) / 256 : 1024), 256, 0, st>>>(out, out_stride, B, lengths, status);
        count_launch();
        return cudaGetLastError();
    }
    return launch_to_newick(v, idx_bytes, B, k, out, out_stride, lengths, status, ws, ws_bytes, st);
}
int p2v_to_newick_batch(const void* v, int idx_bytes, int64_t B, int64_t k, char* out, int64_t out_stride,
                        int64_t* lengths, int32_t* status, void* ws, size_t ws_bytes, p2v_stream_t stream) {
    int rc = newick_check(v, out, idx_bytes, B, k, out_stride);
    if (rc) return rc;
    if (ws_bytes < newick_workspace_bytes(B, k) || (!ws && newick_workspace_bytes(B, k) > 0))
        return fail(P2V_ERR_WORKSPACE, "to_newick workspace too small");
    return from_cuda(newick_launch_any(v, idx_bytes, B, k, out, out_stride, lengths, status, ws, ws_bytes,
                                       (cudaStream_t)stream), "to_newick launch");
}
int p2v_to_newick_host(const void* v, int idx_bytes, int64_t B, int64_t k, char* out, int64_t out_stride,
                       int64_t* lengths, int32_t* status) {
    HostTrim trim_on_return;
    int rc = newick_check(v, out, idx_bytes, B, k, out_stride);
    if (rc) return rc;
    if (B == 0) return P2V_OK;
    int64_t done = 0;
    int64_t* dlen[HOST_STREAMS] = {};
    int64_t cap_chunk = 0;
    int flip = 0;
    int r = host_chunked(
        v, (size_t)k * idx_bytes, out, (size_t)out_stride, status, B,
        [&](int64_t nb) { cap_chunk = nb; return newick_workspace_bytes(nb, k); },
        [&](void* din, void* dout, int32_t* dst, int64_t nb, void* ws, size_t wsb, cudaStream_t st) {
            cudaError_t e = cudaSuccess;
            const int s = flip; flip = (flip + 1) % HOST_STREAMS;
            if (!dlen[s]) e = cudaMalloc((void**)&dlen[s], sizeof(int64_t) * (size_t)cap_chunk);
            if (e != cudaSuccess) return e;
            e = newick_launch_any(din, idx_bytes, nb, k, (char*)dout, out_stride, dlen[s], dst, ws, wsb, st);
            if (e == cudaSuccess && lengths)
                e = cudaMemcpyAsync(lengths + done, dlen[s], sizeof(int64_t) * nb, cudaMemcpyDeviceToHost, st);
            done += nb;
            return e;
        });
    cudaDeviceSynchronize();
    for (int i = 0; i < HOST_STREAMS; ++i) if (dlen[i]) cudaFree(dlen[i]);
    return r;
}

// ---- from_newick (vectors, Newick with parent labels) ------------------------------------------
size_t p2v_from_newick_workspace_bytes(int64_t B, int64_t k, int idx_bytes) { return from_newick_workspace_bytes(B, k, idx_bytes); }
static int from_newick_check(const char* newick, int64_t in_stride, const void* v, int idx_bytes, int64_t B, int64_t k) {
    if (idx_bytes != 4 && idx_bytes != 8) return fail(P2V_ERR_INVALID_ARG, "idx_bytes must be 4 or 8");
    if (B < 0 || k < 0) return fail(P2V_ERR_INVALID_ARG, "negative batch or length");
    if (k > LARGE_K_MAX) return fail(P2V_ERR_UNSUPPORTED, "from_newick: k above 2^19");
    if (B > 0 && (!newick || (k > 0 && !v))) return fail(P2V_ERR_INVALID_ARG, "null buffer");
    if (in_stride < 1) return fail(P2V_ERR_INVALID_ARG, "in_stride must be positive");
    return P2V_OK;
}
int p2v_from_newick_batch(const char* newick, int64_t in_stride, const int64_t* lengths, int64_t B, int64_t k, void* v,
                          int idx_bytes, int32_t* status, void* ws, size_t ws_bytes, p2v_stream_t stream) {
    int rc = from_newick_check(newick, in_stride, v, idx_bytes, B, k);
    if (rc) return rc;
    if (B == 0) return P2V_OK;
    if (k == 0) {                                 // "0;": the empty vector
        if (status) return from_cuda(cudaMemsetAsync(status, 0, sizeof(int32_t) * B, (cudaStream_t)stream), "from_newick status");
        return P2V_OK;
    }
    if (ws_bytes < from_newick_workspace_bytes(B, k, idx_bytes) || !ws)
        return fail(P2V_ERR_WORKSPACE, "from_newick workspace too small");
    return from_cuda(launch_from_newick(newick, in_stride, lengths, B, k, v, idx_bytes, status, ws, ws_bytes,
                                        (cudaStream_t)stream), "from_newick launch");
}
int p2v_from_newick_host(const char* newick, int64_t in_stride, int64_t B, int64_t k, void* v, int idx_bytes,
                         int32_t* status) {
    HostTrim trim_on_return;
    int rc = from_newick_check(newick, in_stride, v, idx_bytes, B, k);
    if (rc) return rc;
    if (k == 0) { if (status && B > 0) memset(status, 0, sizeof(int32_t) * B); return P2V_OK; }
    return host_chunked(
        newick, (size_t)in_stride, v, (size_t)k * idx_bytes, status, B,
        [&](int64_t nb) { return from_newick_workspace_bytes(nb, k, idx_bytes); },
        [&](void* din, void* dout, int32_t* dst, int64_t nb, void* ws, size_t wsb, cudaStream_t st) {
            return launch_from_newick((const char*)din, in_stride, nullptr, nb, k, dout, idx_bytes, dst, ws, wsb, st);
        });
}

// ---- matrices <-> Newick with branch lengths --------------------------------------------------
size_t p2v_newick_matrix_max_bytes(int64_t k) { return newick_matrix_max_bytes(k); }
size_t p2v_newick_matrix_workspace_bytes(int64_t B, int64_t k) { return newick_matrix_workspace_bytes(B, k); }
size_t p2v_from_newick_matrix_workspace_bytes(int64_t B, int64_t k) { return from_newick_matrix_workspace_bytes(B, k); }
static int nm_to_check(const double* m, const char* out, int64_t B, int64_t k, int64_t out_stride) {
    if (B < 0 || k < 0) return fail(P2V_ERR_INVALID_ARG, "negative batch or length");
    if (k == 0 && B > 0) return fail(P2V_ERR_INVALID_ARG, "Matrix should not be empty");
    if (k > LARGE_K_MAX) return fail(P2V_ERR_UNSUPPORTED, "to_newick: k above 2^19");
    if (B > 0 && (!out || !m)) return fail(P2V_ERR_INVALID_ARG, "null buffer");
    if (out_stride < (int64_t)newick_matrix_max_bytes(k)) return fail(P2V_ERR_INVALID_ARG, "out_stride below p2v_newick_matrix_max_bytes(k)");
    return P2V_OK;
}
int p2v_to_newick_matrix_batch(const double* matrix, int64_t B, int64_t k, char* out, int64_t out_stride, int64_t* lengths,
                               int32_t* status, void* ws, size_t ws_bytes, p2v_stream_t stream) {
    int rc = nm_to_check(matrix, out, B, k, out_stride);
    if (rc) return rc;
    if (B == 0) return P2V_OK;
    if (!ws || ws_bytes < newick_matrix_workspace_bytes(B, k)) return fail(P2V_ERR_WORKSPACE, "to_newick workspace too small");
    return from_cuda(launch_to_newick_matrix(matrix, B, k, out, out_stride, lengths, status, ws, ws_bytes,
                                             (cudaStream_t)stream), "to_newick (matrix) launch");
}
int p2v_to_newick_matrix_host(const double* matrix, int64_t B, int64_t k, char* out, int64_t out_stride, int64_t* lengths,
                              int32_t* status) {
    HostTrim trim_on_return;
    int rc = nm_to_check(matrix, out, B, k, out_stride);
    if (rc) return rc;
    if (B == 0) return P2V_OK;
    int64_t done = 0;
    int64_t* dlen[HOST_STREAMS] = {};
    int64_t cap_chunk = 0;
    int flip = 0;
    int r = host_chunked(
        matrix, (size_t)k * 3 * sizeof(double), out, (size_t)out_stride, status, B,
        [&](int64_t nb) { if (nb > cap_chunk) cap_chunk = nb; return newick_matrix_workspace_bytes(nb, k); },
        [&](void* din, void* dout, int32_t* dst, int64_t nb, void* ws, size_t wsb, cudaStream_t st) {
            cudaError_t e = cudaSuccess;
            const int s = flip; flip = (flip + 1) % HOST_STREAMS;
            if (!dlen[s]) e = cudaMalloc((void**)&dlen[s], sizeof(int64_t) * (size_t)cap_chunk);
            if (e != cudaSuccess) return e;
            e = launch_to_newick_matrix((const double*)din, nb, k, (char*)dout, out_stride, dlen[s], dst, ws, wsb, st);
            if (e == cudaSuccess && lengths)
                e = cudaMemcpyAsync(lengths + done, dlen[s], sizeof(int64_t) * nb, cudaMemcpyDeviceToHost, st);
            done += nb;
            return e;
        });
    cudaDeviceSynchronize();
    for (int i = 0; i < HOST_STREAMS; ++i) if (dlen[i]) cudaFree(dlen[i]);
    return r;
}
static int nm_from_check(const char* newick, int64_t in_stride, const double* m, int64_t B, int64_t k) {
    if (B < 0 || k < 0) return fail(P2V_ERR_INVALID_ARG, "negative batch or length");
    if (k > LARGE_K_MAX) return fail(P2V_ERR_UNSUPPORTED, "from_newick: k above 2^19");
    if (B > 0 && (!newick || (k > 0 && !m))) return fail(P2V_ERR_INVALID_ARG, "null buffer");
    if (in_stride < 1) return fail(P2V_ERR_INVALID_ARG, "in_stride must be positive");
    return P2V_OK;
}
int p2v_from_newick_matrix_batch(const char* newick, int64_t in_stride, const int64_t* lengths, int64_t B, int64_t k,
                                 double* matrix, int32_t* status, void* ws, size_t ws_bytes, p2v_stream_t stream) {
    int rc = nm_from_check(newick, in_stride, matrix, B, k);
    if (rc) return rc;
    if (B == 0) return P2V_OK;
    if (k == 0) {
        if (status) return from_cuda(cudaMemsetAsync(status, 0, sizeof(int32_t) * B, (cudaStream_t)stream), "from_newick status");
        return P2V_OK;
    }
    if (!ws || ws_bytes < from_newick_matrix_workspace_bytes(B, k)) return fail(P2V_ERR_WORKSPACE, "from_newick workspace too small");
    return from_cuda(launch_from_newick_matrix(newick, in_stride, lengths, B, k, matrix, status, ws, ws_bytes,
                                               (cudaStream_t)stream), "from_newick (matrix) launch");
}
int p2v_from_newick_matrix_host(const char* newick, int64_t in_stride, int64_t B, int64_t k, double* matrix, int32_t* status) {
    HostTrim trim_on_return;
    int rc = nm_from_check(newick, in_stride, matrix, B, k);
    if (rc) return rc;
    if (k == 0) { if (status && B > 0) memset(status, 0, sizeof(int32_t) * B); return P2V_OK; }
    return host_chunked(
        newick, (size_t)in_stride, matrix, (size_t)k * 3 * sizeof(double), status, B,
        [&](int64_t nb) { return from_newick_matrix_workspace_bytes(nb, k); },
        [&](void* din, void* dout, int32_t* dst, int64_t nb, void* ws, size_t wsb, cudaStream_t st) {
            return launch_from_newick_matrix((const char*)din, in_stride, nullptr, nb, k, (double*)dout, dst, ws, wsb, st);
        });
}

// ---- CSV text -> numbers (the load side of io/reader.py) ----------------------------------------
size_t p2v_csv_workspace_bytes(int64_t nbytes) { return csv_workspace_bytes(nbytes); }
int p2v_csv_parse_batch(const char* text, int64_t nbytes, int delimiter, int kind, void* out, int idx_bytes,
                        int64_t capacity, int64_t* counts, void* ws, size_t ws_bytes, p2v_stream_t stream) {
    if (nbytes < 0 || capacity < 0) return fail(P2V_ERR_INVALID_ARG, "negative size");
    if (kind != 0 && kind != 1) return fail(P2V_ERR_INVALID_ARG, "kind must be 0 (integers) or 1 (float64)");
    if (kind == 0 && idx_bytes != 4 && idx_bytes != 8) return fail(P2V_ERR_INVALID_ARG, "idx_bytes must be 4 or 8");
    if (!counts || (nbytes > 0 && !text) || (capacity > 0 && !out)) return fail(P2V_ERR_INVALID_ARG, "null buffer");
    if (!ws || ws_bytes < csv_workspace_bytes(nbytes)) return fail(P2V_ERR_WORKSPACE, "csv workspace too small");
    return from_cuda(launch_csv_parse(text, nbytes, (char)delimiter, kind, out, idx_bytes, capacity,
                                      reinterpret_cast<long long*>(counts), ws, ws_bytes, (cudaStream_t)stream), "csv parse");
}

int p2v_to_pairs_host(const void* v, int idx_bytes, int64_t B, int64_t k, void* out, int32_t* status) {
    HostTrim trim_on_return;
    return decode_host(MODE_PAIRS, 2, v, idx_bytes, B, k, out, status);
}
int p2v_to_ancestry_host(const void* v, int idx_bytes, int64_t B, int64_t k, void* out, int32_t* status) {
    HostTrim trim_on_return;
    // big int64 batches: two int32 columns over PCIe, widened by host threads (see ancestry_host_compact)
    if (idx_bytes == 8 && k > 0 && k <= 0x3FFFFFFF / 2 && B * k >= ((int64_t)1 << 22) && host_compact_enabled()) {
        int rc = check_common(v, out, idx_bytes, B, k);
        if (rc) return rc;
        return ancestry_host_compact(v, B, k, out, status);
    }
    return decode_host(MODE_ANCESTRY, 3, v, idx_bytes, B, k, out, status);
}
int p2v_to_edges_host(const void* v, int idx_bytes, int64_t B, int64_t k, void* out, int32_t* status) {
    HostTrim trim_on_return;
    return decode_host(MODE_EDGES, 4, v, idx_bytes, B, k, out, status);
}
int p2v_from_pairs_host(const void* in, int idx_bytes, int64_t B, int64_t k, void* v, int32_t* status) {
    HostTrim trim_on_return;
    return encode_host(FROM_PAIRS, 2, in, idx_bytes, B, k, v, status);
}
int p2v_from_ancestry_host(const void* in, int idx_bytes, int64_t B, int64_t k, void* v, int32_t* status) {
    HostTrim trim_on_return;
    return encode_host(FROM_ANCESTRY, 3, in, idx_bytes, B, k, v, status);
}
int p2v_from_edges_host(const void* in, int idx_bytes, int64_t B, int64_t k, void* v, int32_t* status) {
    HostTrim trim_on_return;
    return encode_host(FROM_EDGES, 4, in, idx_bytes, B, k, v, status);
}
int p2v_check_v_host(const void* v, int idx_bytes, int64_t B, int64_t k, int32_t* status) {
    HostTrim trim_on_return;
    if (!status) return fail(P2V_ERR_INVALID_ARG, "status is required");
    int rc = check_common(v, status, idx_bytes, B, k);
    if (rc) return rc;
    if (k == 0 || B == 0) { if (B > 0) memset(status, 0, sizeof(int32_t) * B); return P2V_OK; }
    return host_chunked(
        v, (size_t)k * idx_bytes, nullptr, 0, status, B, [&](int64_t) { return (size_t)0; },
        [&](void* din, void*, int32_t* dst, int64_t nb, void*, size_t, cudaStream_t st) {
            return launch_check_v(din, idx_bytes, nb, k, dst, st);
        });
}
int p2v_cophenetic_host(const void* v, int idx_bytes, const double* bls, int64_t B, int64_t k, int unrooted,
                        int64_t row_begin, int64_t row_end, double* out, int32_t* status) {
    HostTrim trim_on_return;
    if (idx_bytes != 4 && idx_bytes != 8) return fail(P2V_ERR_INVALID_ARG, "idx_bytes must be 4 or 8");
    return coph_host(v, idx_bytes, bls, B, k, unrooted, row_begin, row_end, out, status);
}
int p2v_cophenetic_matrix_host(const double* m, int64_t B, int64_t k, int unrooted, int64_t row_begin,
                               int64_t row_end, double* out, int32_t* status) {
    HostTrim trim_on_return;
    return coph_host(m, 0, nullptr, B, k, unrooted, row_begin, row_end, out, status);
}

void p2v_host_release(void) { g_host.release(); g_compact.release(); g_stage.release(); }
int p2v_host_bind_thread(void) {
    cpu_set_t set;
    if (!gpu_local_cpus(&set)) return 0;
    if (sched_setaffinity(0, sizeof(set), &set) != 0) return 0;
    return CPU_COUNT(&set);
}
int64_t p2v_host_cache_bytes(void) { return (int64_t)(g_host.bytes() + g_stage.bytes() + g_compact.bytes()); }
void p2v_host_cache_limit(int64_t bytes) { g_host_cache_limit.store(bytes, std::memory_order_relaxed); }

// ---- check_m (matrix/base.rs:76-95) ---------------------------------------------------------
static int check_m_args(const double* m, const int32_t* status, int64_t B, int64_t k) {
    if (!status) return fail(P2V_ERR_INVALID_ARG, "status is required");
    if (B < 0 || k < 0) return fail(P2V_ERR_INVALID_ARG, "negative batch or length");
    if (B > 0 && k > 0 && !m) return fail(P2V_ERR_INVALID_ARG, "null buffer");
    if (k > LARGE_K_MAX) return fail(P2V_ERR_UNSUPPORTED, "k above the supported maximum (2^19)");
    return P2V_OK;
}
int p2v_check_m_batch(const double* m, int64_t B, int64_t k, int32_t* status, p2v_stream_t stream) {
    int rc = check_m_args(m, status, B, k);
    if (rc) return rc;
    if (B == 0) return P2V_OK;
    if (k == 0) return from_cuda(cudaMemsetAsync(status, 0, sizeof(int32_t) * B, (cudaStream_t)stream), "memset");
    return from_cuda(launch_check_m(m, m + 1, 3, 3 * k, B, k, status, 0, (cudaStream_t)stream), "check_m launch");
}
int p2v_check_m_host(const double* m, int64_t B, int64_t k, int32_t* status) {
    HostTrim trim_on_return;
    int rc = check_m_args(m, status, B, k);
    if (rc) return rc;
    if (k == 0 || B == 0) { if (B > 0) memset(status, 0, sizeof(int32_t) * B); return P2V_OK; }
    return host_chunked(
        m, (size_t)k * 3 * sizeof(double), nullptr, 0, status, B, [&](int64_t) { return (size_t)0; },
        [&](void* din, void*, int32_t* dst, int64_t nb, void*, size_t, cudaStream_t st) {
            const double* dm = static_cast<const double*>(din);
            return launch_check_m(dm, dm + 1, 3, 3 * k, nb, k, dst, 0, st);
        });
}

// profiling aid: which = 0 decode_large_kernel, 1 tables_kernel (whole-grid forms); out[64] u64
int p2v_debug_phase_ns(int which, uint64_t* out) {
    cudaDeviceSynchronize();
    return from_cuda(which ? debug_phase_ns_coph((unsigned long long*)out) : debug_phase_ns_decode((unsigned long long*)out),
                     "phase stamps");
}

int p2v_fill_f64(double* out, int64_t count, double value, p2v_stream_t stream) {
    if (count > 0 && !out) return fail(P2V_ERR_INVALID_ARG, "null buffer");
    return from_cuda(launch_fill_f64(out, count, value, (cudaStream_t)stream), "fill launch");
}

}  // extern "C"
