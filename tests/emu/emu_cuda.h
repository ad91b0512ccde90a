// Host emulation of the CUDA execution model -- TEST INFRASTRUCTURE ONLY.
//
// The container that builds this repository has no GPU.  To exercise the *same kernel
// source* on the CPU (index maps, FFT staging, reductions, the whole C ABI), the .cu files
// are also compiled as plain C++ with -DPCB_EMU -include emu_cuda.h into
// tests/emu/_build/libpcb200_emu.so.  Every CUDA thread of a block runs as a ucontext
// fiber; __syncthreads() and the warp shuffles are barriers between fibers; blocks run one
// after another.  It is slow (small N only), reports pcb_backend() == "host-emu", and is
// never loaded by the product package (which loads libpcb200.so and requires a CUDA device);
// only tests/ load it, explicitly by path.
#pragma once
#ifndef PCB_EMU
#error "emu_cuda.h is only for the PCB_EMU host-emulation build"
#endif

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct double2 { double x, y; } __attribute__((aligned(16)));
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

typedef int cudaError_t;
typedef void* cudaStream_t;
struct pcbemu_event { std::chrono::steady_clock::time_point t; };
typedef pcbemu_event* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaStreamNonBlocking = 1 };

namespace pcbemu {

// Minimal x86-64 cooperative context switch (callee-saved registers + stack pointer); ucontext's swapcontext makes a
// signal-mask system call per switch, which dominated the emulation time.
#if !defined(__x86_64__)
#error "tests/emu needs x86-64"
#endif
__attribute__((naked, noinline)) static void ctx_switch(void** /*save_sp: rdi*/, void* /*load_sp: rsi*/) {
    asm volatile(
        "pushq %rbp\n pushq %rbx\n pushq %r12\n pushq %r13\n pushq %r14\n pushq %r15\n"
        "movq %rsp, (%rdi)\n"
        "movq %rsi, %rsp\n"
        "popq %r15\n popq %r14\n popq %r13\n popq %r12\n popq %rbx\n popq %rbp\n"
        "ret\n");
}

struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    uint3 tid;
    int lin = 0;            // linear thread index in block
    int state = 0;          // 0 runnable, 1 wait block barrier, 2 wait warp barrier, 3 done
};

struct Block {
    std::vector<Fiber> fibers;
    void* sched_sp = nullptr;
    std::function<void()> body;
    uint3 bid;
    dim3 bdim, gdim;
    std::vector<unsigned char> smem;
    int cur = -1;
    unsigned char xchg[64][32][16];   // per warp, per lane shuffle exchange
};

inline Block*& cur_block() { static Block* b = nullptr; return b; }
inline Fiber& cur_fiber() { Block* b = cur_block(); return b->fibers[b->cur]; }
inline void* dyn_smem() { return cur_block()->smem.data(); }

inline void yield_state(int st) {
    Block* b = cur_block();
    Fiber& f = b->fibers[b->cur];
    f.state = st;
    ctx_switch(&f.sp, b->sched_sp);
}

inline void fiber_entry() {
    Block* b = cur_block();
    b->body();
    Fiber& f = b->fibers[b->cur];
    f.state = 3;
    ctx_switch(&f.sp, b->sched_sp);
    abort();   // a finished fiber is never resumed
}

static const size_t kStack = 512 * 1024;

inline void run_block(Block& b) {
    cur_block() = &b;
    const int nt = (int)b.fibers.size();
    for (int i = 0; i < nt; ++i) {
        Fiber& f = b.fibers[i];
        // initial frame: six zeroed callee-saved slots, then the entry point as the "return address" (16-byte aligned
        // slot, so that the stack is misaligned by 8 at function entry exactly as after a call)
        uintptr_t top = ((uintptr_t)f.stack + kStack) & ~(uintptr_t)15;
        void** frame = (void**)(top - 16);       // [0] = return address slot, [1] = padding
        frame[0] = (void*)&fiber_entry;
        frame[1] = nullptr;
        void** sp = frame - 6;
        for (int q = 0; q < 6; ++q) sp[q] = nullptr;
        f.sp = sp;
        f.state = 0;
    }
    for (;;) {
        bool progressed = false;
        int done = 0;
        for (int i = 0; i < nt; ++i) {
            Fiber& f = b.fibers[i];
            if (f.state == 0) {
                b.cur = i;
                ctx_switch(&b.sched_sp, f.sp);
                progressed = true;
            }
            if (f.state == 3) ++done;
        }
        if (done == nt) break;
        // release barriers
        bool all_block = true;
        for (int i = 0; i < nt; ++i)
            if (b.fibers[i].state != 1 && b.fibers[i].state != 3) { all_block = false; break; }
        if (all_block) {
            for (int i = 0; i < nt; ++i) if (b.fibers[i].state == 1) b.fibers[i].state = 0;
            continue;
        }
        bool released = false;
        for (int w = 0; w * 32 < nt; ++w) {
            bool all_warp = true, any = false;
            for (int i = w * 32; i < nt && i < w * 32 + 32; ++i) {
                int s = b.fibers[i].state;
                if (s == 2) any = true;
                else if (s != 3) { all_warp = false; break; }
            }
            if (all_warp && any) {
                for (int i = w * 32; i < nt && i < w * 32 + 32; ++i)
                    if (b.fibers[i].state == 2) b.fibers[i].state = 0;
                released = true;
            }
        }
        if (!released && !progressed) {
            fprintf(stderr, "pcbemu: deadlock (divergent barrier) in emulated kernel\n");
            abort();
        }
    }
    cur_block() = nullptr;
}

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem, F&& body) {
    static std::vector<char*> stacks;
    Block b;
    const int nt = (int)(block.x * block.y * block.z);
    while ((int)stacks.size() < nt) stacks.push_back((char*)malloc(kStack));
    b.fibers.resize(nt);
    b.body = body;
    b.bdim = block;
    b.gdim = grid;
    b.smem.assign(smem + 16, 0);
    for (int i = 0; i < nt; ++i) {
        b.fibers[i].stack = stacks[i];
        b.fibers[i].lin = i;
        b.fibers[i].tid.x = i % block.x;
        b.fibers[i].tid.y = (i / block.x) % block.y;
        b.fibers[i].tid.z = i / (block.x * block.y);
    }
    for (unsigned z = 0; z < grid.z; ++z)
        for (unsigned y = 0; y < grid.y; ++y)
            for (unsigned x = 0; x < grid.x; ++x) {
                b.bid.x = x; b.bid.y = y; b.bid.z = z;
                run_block(b);
            }
}

template <class T>
inline T shfl(T v, int src_lane) {
    Block* b = cur_block();
    Fiber& f = cur_fiber();
    int w = f.lin / 32, l = f.lin % 32;
    static_assert(sizeof(T) <= 16, "shuffle payload");
    memcpy(b->xchg[w][l], &v, sizeof(T));
    yield_state(2);
    T r;
    int nt = (int)b->fibers.size();
    if (src_lane < 0 || src_lane > 31 || w * 32 + src_lane >= nt) src_lane = l;
    memcpy(&r, b->xchg[w][src_lane], sizeof(T));
    yield_state(2);
    return r;
}

// warp vote: true if the predicate holds on any lane of the calling fiber's warp (all lanes of the warp must call it)
inline bool vote_any(bool pred) {
    Block* b = cur_block();
    Fiber& f = cur_fiber();
    const int w = f.lin / 32, l = f.lin % 32;
    const int nt = (int)b->fibers.size();
    int v = pred ? 1 : 0;
    memcpy(b->xchg[w][l], &v, sizeof(int));
    yield_state(2);
    bool any = false;
    for (int k = 0; k < 32 && w * 32 + k < nt; ++k) {
        int o;
        memcpy(&o, b->xchg[w][k], sizeof(int));
        any = any || (o != 0);
    }
    yield_state(2);
    return any;
}

// mma.sync.aligned.m8n8k4.row.col.f64: D(8x8) += A(8x4) * B(4x8); lane holds A[lane>>2][lane&3], B[lane&3][lane>>2],
// C[lane>>2][2*(lane&3) + {0,1}]  (PTX ISA fragment layout).
inline void mma884(double& c0, double& c1, double a, double b) {
    Block* blk = cur_block();
    Fiber& f = cur_fiber();
    const int w = f.lin / 32, l = f.lin % 32;
    memcpy(blk->xchg[w][l], &a, 8);
    memcpy(blk->xchg[w][l] + 8, &b, 8);
    yield_state(2);
    const int row = l >> 2, tig = l & 3;
    for (int k = 0; k < 4; ++k) {
        double av, b0, b1;
        memcpy(&av, blk->xchg[w][row * 4 + k], 8);
        memcpy(&b0, blk->xchg[w][(2 * tig) * 4 + k] + 8, 8);
        memcpy(&b1, blk->xchg[w][(2 * tig + 1) * 4 + k] + 8, 8);
        c0 += av * b0;
        c1 += av * b1;
    }
    yield_state(2);
}

}  // namespace pcbemu

#define threadIdx (pcbemu::cur_fiber().tid)
#define blockIdx (pcbemu::cur_block()->bid)
#define blockDim (pcbemu::cur_block()->bdim)
#define gridDim (pcbemu::cur_block()->gdim)
static inline void __syncthreads() { pcbemu::yield_state(1); }
static inline void __syncwarp(unsigned = 0xffffffffu) { pcbemu::yield_state(2); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
    return pcbemu::shfl(v, (pcbemu::cur_fiber().lin % 32) ^ m);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d, int = 32) {
    return pcbemu::shfl(v, (pcbemu::cur_fiber().lin % 32) + d);
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return pcbemu::shfl(v, src); }
static inline bool __any_sync(unsigned, bool pred) { return pcbemu::vote_any(pred); }
static inline int __ffs(int v) { return v == 0 ? 0 : __builtin_ffs(v); }
static inline int __ffsll(long long v) { return v == 0 ? 0 : __builtin_ffsll(v); }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline double atomicAdd(double* p, double v) { double o = *p; *p += v; return o; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p += v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p += v; return o; }
static inline unsigned atomicOr(unsigned* p, unsigned v) { unsigned o = *p; *p |= v; return o; }
static inline void __threadfence() {}

// ---- minimal CUDA runtime ------------------------------------------------------------
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = 0) {
    for (size_t i = 0; i < h; ++i) memmove((char*)d + i * dp, (const char*)s + i * sp, w);
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "pcbemu error"; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = *t = (size_t)8 << 30; return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new pcbemu_event; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new pcbemu_event; return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess;
}
template <class K> static inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
