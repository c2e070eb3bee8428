// cuda_emu.h — a minimal CPU emulation of the CUDA constructs used by
// cranesched_b200/csrc/*.cu.  TEST HARNESS ONLY.
//
// Purpose: this container has no GPU. Compiling the *unmodified* kernel source
// with g++ -DCRANE_EMU against this header lets the `-m "not gpu"` tests run
// the real kernel logic (every CUDA thread is a host thread, every
// __syncthreads/__shfl/__ballot is a real barrier/exchange), so logic errors
// and barrier-divergence deadlocks show up here instead of costing a GPU round
// trip. It plays the role the reference's sanitizer builds play
// (CMakeLists.txt:81-83). The product library (libcrane_sched.so, built by
// nvcc) never includes this file and has no CPU path.
//
// Supported: 1-D grids/blocks (blockDim.x multiple of 32), static and dynamic
// shared memory, full-mask warp collectives, global/shared atomics, the small
// part of the CUDA runtime API the C-ABI layer uses.
#pragma once

#include <atomic>
#include <sched.h>
#include <barrier>
#include <condition_variable>
#include <mutex>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __constant__ static
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct uint3_emu { unsigned x = 0, y = 0, z = 0; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

struct alignas(16) uint4 { unsigned x, y, z, w; };

namespace emu {

struct WarpCtx {
  std::barrier<> bar{32};
  uint64_t slot[32];
  explicit WarpCtx(int n) : bar(n) {}
};
struct NamedBar {  // CUDA named barrier: completes when `n` threads have arrived (bar.sync or bar.arrive)
  std::mutex m;
  std::condition_variable cv;
  unsigned count = 0;
  unsigned long gen = 0;
};
struct BlockCtx {
  NamedBar named[16];
  std::barrier<> bar;
  std::vector<std::unique_ptr<WarpCtx>> warps;
  std::vector<unsigned char> dyn_smem;
  explicit BlockCtx(int nthreads) : bar(nthreads) {
    for (int w = 0; w < (nthreads + 31) / 32; ++w)
      warps.emplace_back(new WarpCtx(std::min(32, nthreads - 32 * w)));
  }
};
struct ThreadCtx {
  uint3_emu tid, bid;
  dim3 bdim, gdim;
  WarpCtx* warp = nullptr;
  BlockCtx* block = nullptr;
  int lane = 0;
};
inline thread_local ThreadCtx tctx;

// Runs `body` once per (block, thread). Blocks run one after another; the
// threads of a block run concurrently as host threads.
inline void launch(dim3 grid, dim3 block, size_t smem_bytes,
                   const std::function<void()>& body) {
  const int nt = static_cast<int>(block.x);
  if (nt % 32 != 0 && nt > 32) {
    fprintf(stderr, "cuda_emu: blockDim.x must be a multiple of 32\n");
    abort();
  }
  BlockCtx bc(nt);
  bc.dyn_smem.assign(smem_bytes + 16, (unsigned char)(getenv("CRANE_EMU_SMEM_FILL") ? atoi(getenv("CRANE_EMU_SMEM_FILL")) : 0));
  auto worker = [&](int t) {
    tctx.bdim = block;
    tctx.gdim = grid;
    tctx.block = &bc;
    tctx.warp = bc.warps[t / 32].get();
    tctx.lane = t % 32;
    tctx.tid.x = t;
    for (unsigned b = 0; b < grid.x; ++b) {
      tctx.bid.x = b;
      body();
      bc.bar.arrive_and_wait();  // statics ("__shared__") are reused by the next block
    }
  };
  if (nt == 1) {
    worker(0);
    return;
  }
  std::vector<std::thread> th;
  th.reserve(nt);
  for (int t = 0; t < nt; ++t) th.emplace_back(worker, t);
  for (auto& x : th) x.join();
}

template <class T>
inline uint64_t to_bits(T v) {
  static_assert(sizeof(T) <= 8);
  uint64_t b = 0;
  memcpy(&b, &v, sizeof(T));
  return b;
}
template <class T>
inline T from_bits(uint64_t b) {
  T v;
  memcpy(&v, &b, sizeof(T));
  return v;
}
template <class T>
inline T exchange(T v, int src_lane) {
  WarpCtx* w = tctx.warp;
  w->slot[tctx.lane] = to_bits(v);
  w->bar.arrive_and_wait();
  uint64_t got = w->slot[src_lane & 31];
  w->bar.arrive_and_wait();
  return from_bits<T>(got);
}
}  // namespace emu

#define threadIdx (emu::tctx.tid)
#define blockIdx (emu::tctx.bid)
#define blockDim (emu::tctx.bdim)
#define gridDim (emu::tctx.gdim)
#define warpSize 32

// dynamic shared memory: `extern __shared__ T name[];` is spelled
// CRANE_DYN_SMEM(T, name) in the kernels.
#define CRANE_DYN_SMEM(T, name) \
  T* name = reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(emu::tctx.block->dyn_smem.data()) + 15) & ~uintptr_t(15))

inline void __syncthreads() { emu::tctx.block->bar.arrive_and_wait(); }
inline void emu_named_bar(int id, int n, bool wait) {
  emu::NamedBar& b = emu::tctx.block->named[id];
  std::unique_lock<std::mutex> lk(b.m);
  const unsigned long g = b.gen;
  if (++b.count == (unsigned)n) {
    b.count = 0;
    ++b.gen;
    b.cv.notify_all();
  } else if (wait) {
    b.cv.wait(lk, [&] { return b.gen != g; });
  }
}
inline void __syncwarp(unsigned = 0xffffffffu) { emu::tctx.warp->bar.arrive_and_wait(); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline unsigned __activemask() { return 0xffffffffu; }

template <class T>
inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu::exchange(v, src); }
template <class T>
inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
  int l = emu::tctx.lane;
  T got = emu::exchange(v, l >= (int)d ? l - (int)d : l);
  return l >= (int)d ? got : v;
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
  int l = emu::tctx.lane;
  T got = emu::exchange(v, l + (int)d < 32 ? l + (int)d : l);
  return l + (int)d < 32 ? got : v;
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
  return emu::exchange(v, emu::tctx.lane ^ m);
}
inline unsigned __ballot_sync(unsigned, int pred) {
  emu::WarpCtx* w = emu::tctx.warp;
  w->slot[emu::tctx.lane] = pred ? 1 : 0;
  w->bar.arrive_and_wait();
  unsigned m = 0;
  for (int i = 0; i < 32; ++i) m |= (w->slot[i] ? 1u : 0u) << i;
  w->bar.arrive_and_wait();
  return m;
}
inline unsigned __reduce_and_sync(unsigned, unsigned v) {
  emu::WarpCtx* w = emu::tctx.warp;
  w->slot[emu::tctx.lane] = v;
  w->bar.arrive_and_wait();
  unsigned r = 0xffffffffu;
  for (int i = 0; i < 32; ++i) r &= (unsigned)w->slot[i];
  w->bar.arrive_and_wait();
  return r;
}
inline unsigned __reduce_max_sync(unsigned, unsigned v) {
  emu::WarpCtx* w = emu::tctx.warp;
  w->slot[emu::tctx.lane] = v;
  w->bar.arrive_and_wait();
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r = (unsigned)w->slot[i] > r ? (unsigned)w->slot[i] : r;
  w->bar.arrive_and_wait();
  return r;
}
inline unsigned __vmaxu4(unsigned a, unsigned b) {
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) {
    unsigned x = (a >> (8 * i)) & 0xff, y = (b >> (8 * i)) & 0xff;
    r |= (x > y ? x : y) << (8 * i);
  }
  return r;
}
inline long long clock64() { return (long long)std::chrono::steady_clock::now().time_since_epoch().count(); }
inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, p) == 0xffffffffu; }
inline int __syncthreads_or(int p) {
  static std::atomic<int> acc{0};
  if (p) acc.store(1);
  __syncthreads();
  int r = acc.load();
  __syncthreads();
  if (threadIdx.x == 0) acc.store(0);
  __syncthreads();
  return r;
}

inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
inline unsigned __brev(unsigned v) {
  unsigned r = 0;
  for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i);
  return r;
}
template <class T>
inline T __ldg(const T* p) { return *p; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __ll2double_rn(long long v) { return static_cast<double>(v); }
inline double __ull2double_rn(unsigned long long v) { return static_cast<double>(v); }
inline double __uint2double_rn(unsigned v) { return static_cast<double>(v); }
inline long long __double_as_longlong(double d) { return emu::from_bits<long long>(emu::to_bits(d)); }
inline double __longlong_as_double(long long v) { return emu::from_bits<double>(emu::to_bits(v)); }

#define EMU_ATOMIC_RMW(name, builtin)                                             \
  template <class T>                                                              \
  inline T name(T* p, T v) { return builtin(p, v, __ATOMIC_SEQ_CST); }
EMU_ATOMIC_RMW(atomicAdd, __atomic_fetch_add)
EMU_ATOMIC_RMW(atomicOr, __atomic_fetch_or)
EMU_ATOMIC_RMW(atomicAnd, __atomic_fetch_and)
EMU_ATOMIC_RMW(atomicExch, __atomic_exchange_n)
template <class T>
inline T atomicMax(T* p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
template <class T>
inline T atomicMin(T* p, T v) {
  T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
template <class T>
inline T atomicCAS(T* p, T cmp, T v) {
  __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return cmp;
}

// ---------------------------------------------------------------------------
// CUDA runtime API subset
// ---------------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
struct EmuEvent { std::chrono::steady_clock::time_point t; };
typedef EmuEvent* cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0 };
struct cudaDeviceProp { int multiProcessorCount = 148; int major = 10, minor = 0; char name[64] = "cuda_emu"; size_t sharedMemPerBlockOptin = 232448; };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaPeekAtLastError() { return 0; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { *p = cudaDeviceProp(); return 0; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? 0 : 2; }
template <class T>
inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n); }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) memset(d, v, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { if (n) memset(d, v, n); return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new EmuEvent(); return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return 0;
}
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }
#define cudaMemcpyToSymbolAsync(sym, src, n, off, kind, stream) (memcpy(reinterpret_cast<char*>(&(sym)) + (off), (src), (n)), 0)
#define cudaMemcpyToSymbol(sym, src, n) (memcpy(&(sym), (src), (n)), 0)

// kernel launch: CRANE_LAUNCH(kernel, grid, block, smem, stream, args...)
#define CRANE_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu::launch(dim3(grid), dim3(block), (smem), [&]() { kernel(__VA_ARGS__); })
