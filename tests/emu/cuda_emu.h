// TEST INFRASTRUCTURE ONLY — a minimal CUDA-on-host-threads shim.
//
// `make -C tests/emu` compiles whale.jl_b200/csrc/whalecuda.cu as plain C++ with -DWHALE_EMU against this
// header into tests/emu/libwhalecuda_emu.so.  Each CTA runs as blockDim.x cooperative fibers on one host
// thread (a barrier yields to a round-robin scheduler); a few host threads run independent CTAs concurrently.  It exists so the packer and the kernel logic
// can be debugged (and regression-tested by `pytest -m "not gpu"`) on the GPU-less build container.
// It is NOT a fallback: the package only ever loads whale.jl_b200/libwhalecuda.so, which requires a GPU.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <ucontext.h>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __restrict__
#define __shared__ static thread_local

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
inline thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

struct uint4 { unsigned x, y, z, w; };
struct double2 { double x, y; };
struct int4 { int x, y, z, w; };
struct int2 { int x, y; };
inline int2 make_int2(int x, int y) { return int2{x, y}; }
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
// explicit round-to-nearest ops (the emulation build is compiled with -ffp-contract=off)
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double2 make_double2(double x, double y) { return double2{x, y}; }
template <class T> inline T __ldg(const T* p) { return *p; }
inline double __hiloint2double(int hi, int lo) {
    uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double d; memcpy(&d, &u, 8); return d;
}
inline int __double2hiint(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(uint32_t)(u >> 32); }
inline int __double2loint(double d) { uint64_t u; memcpy(&u, &d, 8); return (int)(uint32_t)u; }
inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
using std::min;
using std::max;
using std::isfinite;

namespace emu {
// One CTA = blockDim.x cooperative fibers (ucontext) on ONE host thread; a barrier is "arrive, then yield to the
// round-robin scheduler until the generation changes".  CTAs are independent, so a small pool of host threads
// runs them concurrently; everything a CTA shares (barriers, shared memory, __shared__ statics) is thread_local.
struct Barrier { int n = 0, count = 0, gen = 0; };
struct Cta {
    ucontext_t sched;
    std::vector<ucontext_t> ctx;
    std::vector<unsigned char> done;
    std::vector<int> or_phase;
    int cur = 0;
};
inline thread_local Cta* g_cta = nullptr;
inline thread_local Barrier* g_bar = nullptr;
inline thread_local std::vector<Barrier>* g_wbar = nullptr;
inline thread_local unsigned char* g_smem = nullptr;
inline thread_local size_t g_smem_bytes = 0;
inline thread_local const std::function<void()>* g_body = nullptr;
inline void yield() {
    Cta* c = g_cta;
    const int me = c->cur;
    swapcontext(&c->ctx[me], &c->sched);
}
inline void bar_wait(Barrier& b) {
    const int g = b.gen;
    if (++b.count == b.n) { b.count = 0; b.gen++; return; }
    while (b.gen == g) yield();
}
inline void fiber_main() {
    Cta* c = g_cta;
    const int me = c->cur;
    (*g_body)();
    c->done[me] = 1;
    swapcontext(&c->ctx[me], &c->sched);
}
constexpr size_t kFiberStack = 256 << 10;
inline void run_cta(int b, int grid, int block, size_t smem, const std::function<void()>& body,
                    std::vector<unsigned char>& stacks) {
    std::vector<unsigned char> sm(smem + 64);
    Barrier bar; bar.n = block;
    std::vector<Barrier> wb((block + 31) / 32);
    for (size_t w = 0; w < wb.size(); w++) wb[w].n = std::min(32, block - (int)w * 32);
    Cta cta;
    cta.ctx.resize(block); cta.done.assign(block, 0); cta.or_phase.assign(block, 0);
    g_cta = &cta; g_bar = &bar; g_wbar = &wb; g_body = &body;
    g_smem = (unsigned char*)(((uintptr_t)sm.data() + 15) & ~(uintptr_t)15);
    g_smem_bytes = smem;
    blockIdx.x = b; blockDim.x = block; gridDim.x = grid;
    for (int t = 0; t < block; t++) {
        getcontext(&cta.ctx[t]);
        cta.ctx[t].uc_stack.ss_sp = stacks.data() + (size_t)t * kFiberStack;
        cta.ctx[t].uc_stack.ss_size = kFiberStack;
        cta.ctx[t].uc_link = &cta.sched;
        makecontext(&cta.ctx[t], (void (*)())fiber_main, 0);
    }
    // WHALE_EMU_SCHED=reverse|random changes the order in which the fibers run between two barriers: a result that
    // depends on it means a missing barrier (or a warp-lockstep assumption) in the kernel
    static const int sched = [] {
        const char* e = getenv("WHALE_EMU_SCHED");
        return !e ? 0 : !strcmp(e, "reverse") ? 1 : !strcmp(e, "random") ? 2 : 0;
    }();
    std::vector<int> ord(block);
    for (int t = 0; t < block; t++) ord[t] = sched == 1 ? block - 1 - t : t;
    uint64_t rng = 0x9E3779B97F4A7C15ull * (uint64_t)(b + 1);
    for (int live = block; live > 0;) {
        if (sched == 2)
            for (int i = block - 1; i > 0; i--) {
                rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                std::swap(ord[i], ord[(int)((rng >> 33) % (uint64_t)(i + 1))]);
            }
        for (int k = 0; k < block; k++) {
            const int t = ord[k];
            if (cta.done[t] == 2) continue;
            cta.cur = t; threadIdx.x = t;
            swapcontext(&cta.sched, &cta.ctx[t]);
            if (cta.done[t] == 1) { cta.done[t] = 2; live--; }
        }
    }
    g_cta = nullptr;
}
inline void launch(int grid, int block, size_t smem, const std::function<void()>& body) {
    int nw = (int)std::thread::hardware_concurrency();
    if (const char* e = getenv("WHALE_EMU_THREADS")) nw = atoi(e);
    nw = std::max(1, std::min(nw, grid));
    std::atomic<int> next{0};
    auto worker = [&] {
        std::vector<unsigned char> stacks((size_t)block * kFiberStack);
        for (int b; (b = next.fetch_add(1)) < grid;) run_cta(b, grid, block, smem, body, stacks);
    };
    if (nw == 1) { worker(); return; }
    std::vector<std::thread> th;
    for (int i = 0; i < nw; i++) th.emplace_back(worker);
    for (auto& x : th) x.join();
}
}  // namespace emu
namespace emu {
inline thread_local double g_shfl[64][32];  // per-warp exchange buffer
// warp shuffle (all 32 lanes of the warp must call it, as on the device with a full mask)
inline double shfl_down(double v, int delta) {
    const int w = threadIdx.x / 32, l = threadIdx.x % 32;
    g_shfl[w][l] = v;
    bar_wait((*g_wbar)[w]);
    const double r = (l + delta < 32) ? g_shfl[w][l + delta] : v;
    bar_wait((*g_wbar)[w]);
    return r;
}
inline bool warp_any(bool p) {
    const int w = threadIdx.x / 32, l = threadIdx.x % 32;
    g_shfl[w][l] = p ? 1.0 : 0.0;
    bar_wait((*g_wbar)[w]);
    bool r = false;
    for (int i = 0; i < (*g_wbar)[w].n; i++) r = r || g_shfl[w][i] != 0.0;
    bar_wait((*g_wbar)[w]);
    return r;
}
inline double shfl_idx(double v, int src) {
    const int w = threadIdx.x / 32, l = threadIdx.x % 32;
    g_shfl[w][l] = v;
    bar_wait((*g_wbar)[w]);
    const double r = g_shfl[w][src & 31];
    bar_wait((*g_wbar)[w]);
    return r;
}
}  // namespace emu
inline void __syncthreads() { emu::bar_wait(*emu::g_bar); }
namespace emu { inline thread_local int g_or_flag[2] = {0, 0}; }
inline int __syncthreads_or(int p) {  // barrier + OR of the predicate over the CTA (double-buffered flag)
    int& phase = emu::g_cta->or_phase[threadIdx.x];
    const int ph = phase; phase ^= 1;
    if (p) emu::g_or_flag[ph] = 1;
    emu::bar_wait(*emu::g_bar);
    const int r = emu::g_or_flag[ph];
    emu::bar_wait(*emu::g_bar);
    if (threadIdx.x == 0) emu::g_or_flag[ph] = 0;
    emu::bar_wait(*emu::g_bar);
    return r;
}
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline void __syncwarp() { emu::bar_wait((*emu::g_wbar)[threadIdx.x / 32]); }
#define EXTERN_SHARED(name) unsigned char* name = emu::g_smem

// ---- runtime API subset ----
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef struct { std::chrono::steady_clock::time_point t; }* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int multiProcessorCount = 2; };
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { *p = cudaDeviceProp(); return 0; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(1, n ? n : 1); return 0; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = calloc(1, n ? n : 1); return 0; }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return 0; }
inline cudaError_t cudaMemGetInfo(size_t* fre, size_t* tot) { *fre = (size_t)8 << 30; *tot = (size_t)16 << 30; return 0; }  // (pretend: 8 GiB free)
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
enum { cudaEventDisableTiming = 2 };
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return 0; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new std::remove_pointer<cudaEvent_t>::type(); return 0; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new std::remove_pointer<cudaEvent_t>::type(); return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return 0;
}
