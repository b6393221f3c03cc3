// Shared device/host helpers for the sm_100a decoder kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <vector>

#include "../../include/dic.h"

namespace dic {

typedef __nv_bfloat16 bf16;

// ---- error plumbing (thread-local message returned by dic_last_error) ----------------
extern thread_local char g_err[512];

#define DIC_FAIL(code, ...)                          \
  do {                                               \
    snprintf(dic::g_err, sizeof(dic::g_err), __VA_ARGS__); \
    return (code);                                   \
  } while (0)

#define DIC_CUDA(expr)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      DIC_FAIL(-2, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define DIC_LAUNCH_CHECK()                                                                   \
  do {                                                                                       \
    dic::g_launches.fetch_add(1, std::memory_order_relaxed);                                 \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess)                                                                   \
      DIC_FAIL(-3, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define DIC_TRY(expr)       \
  do {                      \
    int _r = (expr);        \
    if (_r != 0) return _r; \
  } while (0)

// ---- launch counter and optional per-kernel-class CUDA-event profiling ------------------------
// (bench.py reads these through dic_launch_count / dic_profile_*; off by default, no cost)
inline std::atomic<long long> g_launches{0};

enum ProfClass {
  P_ATTN_FWD = 0, P_ATTN_BWD, P_DATT1, P_GEMM_TC, P_GEMM_FMA, P_LSTM, P_FUSE, P_COLSUM,
  P_ATTN_ALPHA, P_ATTN_BWD_SMALL, P_DFEAT, P_BEAM_SELECT, P_DECODE_MISC, P_LOSS, P_GEMM_ATT1, P_GEMM_LOGITS, P_N
};
inline const char* prof_class_name(int c) {
  static const char* names[P_N] = {"attn_context_fwd", "attn_stream_bwd", "datt1", "gemm_tcgen05",
                                   "gemm_fma", "lstm_pointwise", "fuse_feats", "colsum",
                                   "attn_alpha_fwd", "attn_small_bwd", "dfeat_accumulate",
                                   "beam_select", "decode_misc", "caption_loss", "gemm_att1", "gemm_logits"};
  return (c >= 0 && c < P_N) ? names[c] : "?";
}
struct ProfState {
  bool on = false;
  std::vector<cudaEvent_t> ev[P_N];   // start/end pairs
  size_t used[P_N] = {0};
  double bytes[P_N] = {0};            // algorithmic bytes attributed by the launch site
};
inline ProfState g_prof;

struct ProfScope {
  int cls;
  cudaStream_t st;
  bool active;
  ProfScope(int c, cudaStream_t s, double algo_bytes = 0.0) : cls(c), st(s), active(g_prof.on) {
    if (!active) return;
    auto& v = g_prof.ev[cls];
    size_t& u = g_prof.used[cls];
    while (v.size() < u + 2) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      v.push_back(e);
    }
    cudaEventRecord(v[u], st);
    g_prof.bytes[cls] += algo_bytes;
  }
  ~ProfScope() {
    if (!active) return;
    size_t& u = g_prof.used[cls];
    cudaEventRecord(g_prof.ev[cls][u + 1], st);
    u += 2;
  }
};

// One-time-per-DEVICE guard (cudaFuncSetAttribute is a per-device setting, and two host threads -- the
// caller's and PyTorch's autograd worker -- may reach a launch site at once: the flags are an atomic bit
// mask; the guarded setup is idempotent, so two threads running it concurrently is harmless).
struct DeviceOnce {
  std::atomic<unsigned long long> done{0};
  bool need(int* dev) {
    cudaGetDevice(dev);
    return ((done.load(std::memory_order_acquire) >> (*dev & 63)) & 1ull) == 0;
  }
  void mark(int dev) { done.fetch_or(1ull << (dev & 63), std::memory_order_release); }
};

// ---- dtype-erased scalar access --------------------------------------------------------
__device__ __forceinline__ float ld_as_float(const void* p, size_t i, int is_bf16) {
  return is_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(p)[i])
                 : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void st_from_float(void* p, size_t i, float v, int is_bf16) {
  if (is_bf16)
    reinterpret_cast<bf16*>(p)[i] = __float2bfloat16_rn(v);
  else
    reinterpret_cast<float*>(p)[i] = v;
}

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 8 consecutive elements -> fp32 registers (16 B for bf16, 2 x 16 B for fp32); p must be
// 16-byte aligned.  Streaming variant bypasses L1 allocation (annotations are read once per CTA).
template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p));
  float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<bf16>(const bf16* p, float (&v)[8]) {
  uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}

template <typename T>
__device__ __forceinline__ void load8_stream(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8_stream<float>(const float* p, float (&v)[8]) {
  float4 a, b;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p));
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8_stream<bf16>(const bf16* p, float (&v)[8]) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}

// 4 consecutive elements (8 B for bf16, 16 B for fp32); p must be aligned to that size
template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
template <>
__device__ __forceinline__ void load4<bf16>(const bf16* p, float (&v)[4]) {
  const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
  const float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
  v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
}
template <typename T>
__device__ __forceinline__ void store4(T* p, const float (&v)[4]);
template <>
__device__ __forceinline__ void store4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void store4<bf16>(bf16* p, const float (&v)[4]) {
  uint2 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
  h[0] = __floats2bfloat162_rn(v[0], v[1]);
  h[1] = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = r;
}

template <typename T, int N>
__device__ __forceinline__ void vloadN(const T* p, float (&v)[N]) {
  if constexpr (N == 8) load8<T>(p, v); else load4<T>(p, v);
}

// Raw (unconverted) 8-element vectors: streaming kernels keep several of these in flight per
// thread; holding bf16 data packed (4 registers instead of 8) is what lets 8+ CTAs fit per SM.
template <typename T>
struct Raw8;
template <>
struct Raw8<bf16> {
  uint4 r;
  __device__ __forceinline__ void load_stream(const bf16* p) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  }
  __device__ __forceinline__ void zero() { r = make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
};
template <>
struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load_stream(const float* p) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p));
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + 4));
  }
  __device__ __forceinline__ void zero() { a = make_float4(0.f, 0.f, 0.f, 0.f); b = a; }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

// N-element (8 or 4) raw vectors with streaming loads: RawV<T, 8> = Raw8<T>; RawV<T, 4> is half of it
template <typename T, int N>
struct RawV;
template <typename T>
struct RawV<T, 8> : Raw8<T> {};
template <>
struct RawV<bf16, 4> {
  uint2 r;
  __device__ __forceinline__ void load_stream(const bf16* p) {
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  }
  __device__ __forceinline__ void zero() { r = make_uint2(0u, 0u); }
  __device__ __forceinline__ void unpack(float (&v)[4]) const {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
    const float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
    v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
  }
};
template <>
struct RawV<float, 4> {
  float4 a;
  __device__ __forceinline__ void load_stream(const float* p) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p));
  }
  __device__ __forceinline__ void zero() { a = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void unpack(float (&v)[4]) const { v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; }
};

template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&v)[8]);
template <>
__device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store8<bf16>(bf16* p, const float (&v)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = r;
}

template <typename T, int N>
__device__ __forceinline__ void vstoreN(T* p, const float (&v)[N]) {
  if constexpr (N == 8) store8<T>(p, v); else store4<T>(p, v);
}

// ---- reductions ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum / max over blockDim.x threads (multiple of 32, <= 1024).  `scratch` holds
// >= 33 floats of shared memory.  All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();  // scratch may still be read from a previous reduction
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nw ? scratch[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nw ? scratch[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }
// bf16 mode: ex2.approx + rcp.approx (relative error ~1e-6, far below the bf16 storage of what it gates)
// (four instructions: __expf carries range fix-ups -- three multiplies and a compare around the ex2 -- and __frcp_rn is
// an IEEE-rounded reciprocal; the gate this feeds is stored / multiplied in bf16)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoidf_fast(float x) { return rcp_approx(1.f + ex2_approx(x * -1.4426950408889634f)); }

// per-step valid batch sizes, passed to kernels by value
struct StepSizes {
  int32_t n[DIC_MAX_STEPS];
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------
// The decoder time loop is a chain of 5-10 short dependent kernels per step.  Launched with the
// programmatic-stream-serialization attribute, kernel N+1 is scheduled while kernel N is still
// running: its prologue (shared-memory carve-up, barrier init, TMEM allocation, tensor-map prefetch)
// overlaps N's tail, and it blocks in pdl_wait() -- before its first access to global memory that N
// may write or still read -- until N has completed and flushed.  Every kernel in the chain calls
// pdl_wait() unconditionally, so the ordering guarantees are those of plain stream order.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }

inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DIC_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  // event profiling puts event records between the kernels; keep those launches plain
  return v == 1 && !g_prof.on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- in-kernel timeline trace (debug; off unless dic_trace_start was called) ----------------------
// Thread 0 of every CTA records %globaltimer at entry, after its programmatic-dependency wait and
// at exit.  This is the only way to see the real timeline of the PDL-chained step kernels (ncu
// serialises launches and flushes caches; CUDA events between kernels disable the overlap).
struct TraceRec {
  unsigned long long t0, t1, t2;
  int kid, blk;
};
enum TraceKid {
  TK_GEMM_TC = 0 /* + GemmArgs.tag * 100 */, TK_ALPHA = 1, TK_CTX, TK_LSTM_FWD, TK_LSTM_BWD, TK_BWD_STREAM,
  TK_BWD_SMALL, TK_ARGMAX, TK_BEAM_TOPK, TK_BEAM_MERGE, TK_BEAM_REORDER, TK_GEMM_FMA, TK_MISC
};
inline TraceRec* g_trace_host = nullptr;      // host mirror: launch sites pass it to the kernels by value
__device__ unsigned int g_trace_cap = 0;
__device__ unsigned int g_trace_cnt = 0;
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
struct Trace {
  TraceRec* rec = nullptr;     // this CTA's record (thread 0 only, when tracing is on): the only live state
  // `b` comes from the kernel's argument block (constant bank): zero cost when tracing is off
  __device__ __forceinline__ explicit Trace(TraceRec* b) {
#ifndef DIC_NO_TRACE
    if (b != nullptr && threadIdx.x == 0) {
      const unsigned int i = atomicAdd(&g_trace_cnt, 1u);
      if (i < g_trace_cap) {
        rec = b + i;
        rec->t1 = 0;
        rec->t0 = gtimer();
      }
    }
#endif
  }
  __device__ __forceinline__ void mark() {
#ifndef DIC_NO_TRACE
    if (rec) rec->t1 = gtimer();
#endif
  }
  __device__ __forceinline__ void end(int kid) {
#ifndef DIC_NO_TRACE
    if (rec) {
      rec->kid = kid;
      rec->blk = (int)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z));
      rec->t2 = gtimer();
    }
#endif
  }
};

// ---- sub-batch streams -----------------------------------------------------------------------------
// A decoder timestep is a chain of short dependent kernels around one HBM-bound pass over the
// annotations.  Images are independent, so the time loop runs over S contiguous sub-batches on S
// streams: while one sub-batch streams its annotations the others sit in their latency-bound
// kernels, and the chain latency is hidden instead of serialised.  Sub-batch 0 stays on the caller's
// stream; the others use library-owned non-blocking streams forked / joined with events (no host
// synchronisation; the caller still sees plain stream order).
constexpr int kMaxSub = 8;
struct SubStreams {
  cudaStream_t s[kMaxSub] = {nullptr};
  cudaEvent_t fork_ev = nullptr;
  cudaEvent_t join_ev[kMaxSub] = {nullptr};
  bool ready = false;
};
inline thread_local SubStreams g_sub;
inline std::atomic<int> g_sub_override{0};   // dic_set_substreams: 0 = heuristic

inline int sub_fork(cudaStream_t main_st, int S, cudaStream_t* out) {
  out[0] = main_st;
  if (S <= 1) return 0;
  SubStreams& g = g_sub;
  if (!g.ready) {
    for (int i = 1; i < kMaxSub; ++i) {
      DIC_CUDA(cudaStreamCreateWithFlags(&g.s[i], cudaStreamNonBlocking));
      DIC_CUDA(cudaEventCreateWithFlags(&g.join_ev[i], cudaEventDisableTiming));
    }
    DIC_CUDA(cudaEventCreateWithFlags(&g.fork_ev, cudaEventDisableTiming));
    g.ready = true;
  }
  DIC_CUDA(cudaEventRecord(g.fork_ev, main_st));
  for (int i = 1; i < S; ++i) {
    out[i] = g.s[i];
    DIC_CUDA(cudaStreamWaitEvent(g.s[i], g.fork_ev, 0));
  }
  return 0;
}
inline int sub_join(cudaStream_t main_st, int S) {
  SubStreams& g = g_sub;
  for (int i = 1; i < S; ++i) {
    DIC_CUDA(cudaEventRecord(g.join_ev[i], g.s[i]));
    DIC_CUDA(cudaStreamWaitEvent(main_st, g.join_ev[i], 0));
  }
  return 0;
}
// sub-batch boundaries: multiples of `quantum` rows, S <= kMaxSub
inline int sub_bounds(int B, int S, int quantum, int* r0) {
  if (S > kMaxSub) S = kMaxSub;
  if (S < 1) S = 1;
  int per = cdiv(cdiv(B, S), quantum) * quantum;
  int n = 0;
  for (int r = 0; r < B; r += per) r0[n++] = r;
  r0[n] = B;
  return n;
}

}  // namespace dic
