// Data-parallel gradient all-reduce over NVLink peer memory (the one exchange step of the path).
//
// Every rank holds the same flat fp32 gradient buffer in peer-visible ("symmetric") memory; bufs[r] is rank r's
// buffer as seen from THIS rank.  Two-shot, in place, one kernel:
//
//   start barrier  block b of every rank has started (=> every rank's backward kernels have completed)
//   reduce + push  rank r owns slice r of the buffer: it reads that slice from every rank (16-byte loads over
//                  NVLink, summed in rank order 0..n-1, so every rank ends up with bit-identical values), scales,
//                  and writes the result into slice r of EVERY rank's buffer
//   end barrier    block b of every rank has finished pushing
//
// Element i of slice r is read and then overwritten on all ranks by one thread of rank r only: no other rank ever
// touches it, so the exchange needs no staging copy.  The barriers are per block (block b of rank r waits for
// block b of the other ranks) through flag words in the peers' memory: st.release.sys / ld.acquire.sys, epoch
// counted by the host, two flag sets so that a fast rank entering the next call cannot overrun a slow one.
// With a multicast mapping (NVLS) the reduce and the push are one multimem.ld_reduce / multimem.st each and the
// NVSwitch does the arithmetic.
//
// Why not NCCL here: profiles/r02_dp_timeline_n2.txt -- three bucketed ncclAllReduce kernels (RING_LL, 75-100 us
// each for 1-13 MB) ran next to the backward kernels and stretched them (dW_enc 43 -> 73 us, dL/dF 94 -> 141 us);
// 0.115 ms of the 2.83 ms step stayed exposed whatever the protocol (NCCL_PROTO / NCCL_ALGO / NCCL_MAX_CTAS sweep
// in profiles/r02_dp_nccl_sweep.txt).
#pragma once
#include "common.cuh"

namespace dic {

constexpr int kDpMaxRanks = 8;
constexpr int kDpMaxBlocks = 148;
constexpr int kDpThreads = 512;
// flag block of one rank: [2 sets][kDpMaxBlocks][kDpMaxRanks] uint32
constexpr size_t kDpFlagWords = (size_t)2 * kDpMaxBlocks * kDpMaxRanks;

struct DpArgs {
  float* bufs[kDpMaxRanks];
  uint32_t* flags[kDpMaxRanks];
  float* mc;               // multicast (NVLS) mapping of the buffer or null
  int rank, world;
  long long n4;            // float4 elements in the buffer (a multiple of world)
  float scale;
  uint32_t epoch;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Data moves with weak 16-byte accesses that bypass L1 (the flag protocol orders them: a block reads only after its
// acquire of the peers' start flags and releases its end flag only after a system-scope fence behind its stores).
// .relaxed.sys vector accesses measured 53 us per 19.3 MB all-reduce on two GPUs.
__device__ __forceinline__ float4 ld_peer16(const float4* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer16(float4* p, float4 v) {
  asm volatile("st.global.cg.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__device__ __forceinline__ void dp_barrier(const DpArgs& p, int set) {
  const int tid = threadIdx.x, b = blockIdx.x;
  __syncthreads();
  if (tid < p.world) {
    __threadfence_system();
    const size_t slot = ((size_t)set * kDpMaxBlocks + b) * kDpMaxRanks;
    st_release_sys(p.flags[tid] + slot + p.rank, p.epoch);          // "block b of rank `rank` is here" on peer tid
    const uint32_t* mine = p.flags[p.rank] + slot + tid;
    unsigned long long spins = 0;
    while ((int)(ld_acquire_sys(mine) - p.epoch) < 0) {
      if (++spins > (1ull << 31)) __trap();                         // a peer never arrived: fail loudly, do not hang
    }
  }
  __syncthreads();
}

// W = compiled rank count (2, 4 or 8; the world size is rounded up to it): 16 / W vectors per thread and pass, so a
// thread always has 16 sixteen-byte loads in flight whatever the world size.
template <bool MULTIMEM, int W>
__global__ void __launch_bounds__(kDpThreads, 1) dp_allreduce_kernel(const DpArgs p) {
  constexpr int kDpUnroll = MULTIMEM ? 8 : 16 / W;
  pdl_wait();
  const int tid = threadIdx.x, b = blockIdx.x, nb = gridDim.x;
  dp_barrier(p, 0);
  const long long per = p.n4 / p.world;
  const long long lo = per * p.rank, hi = lo + per;
  const long long stride = (long long)nb * kDpThreads;
  for (long long i0 = lo + (long long)b * kDpThreads + tid; i0 < hi; i0 += stride * kDpUnroll) {
    if constexpr (MULTIMEM) {
      float4 acc[kDpUnroll];
#pragma unroll
      for (int u = 0; u < kDpUnroll; ++u) {
        const long long i = i0 + u * stride;
        if (i < hi)
          asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                       : "=f"(acc[u].x), "=f"(acc[u].y), "=f"(acc[u].z), "=f"(acc[u].w)
                       : "l"(reinterpret_cast<const float4*>(p.mc) + i) : "memory");
      }
#pragma unroll
      for (int u = 0; u < kDpUnroll; ++u) {
        const long long i = i0 + u * stride;
        if (i < hi) {
          const float4 v = make_float4(acc[u].x * p.scale, acc[u].y * p.scale, acc[u].z * p.scale, acc[u].w * p.scale);
          asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(reinterpret_cast<float4*>(p.mc) + i),
                       "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
      }
    } else {
      float4 v[kDpUnroll][W];
#pragma unroll
      for (int u = 0; u < kDpUnroll; ++u) {
        const long long i = i0 + u * stride;
#pragma unroll
        for (int r = 0; r < W; ++r)
          if (r < p.world && i < hi) v[u][r] = ld_peer16(reinterpret_cast<const float4*>(p.bufs[r]) + i);
      }
#pragma unroll
      for (int u = 0; u < kDpUnroll; ++u) {
        const long long i = i0 + u * stride;
        if (i >= hi) continue;
        float4 s = v[u][0];
#pragma unroll
        for (int r = 1; r < W; ++r)
          if (r < p.world) { s.x += v[u][r].x; s.y += v[u][r].y; s.z += v[u][r].z; s.w += v[u][r].w; }
        s.x *= p.scale; s.y *= p.scale; s.z *= p.scale; s.w *= p.scale;
#pragma unroll
        for (int r = 0; r < W; ++r)
          if (r < p.world) st_peer16(reinterpret_cast<float4*>(p.bufs[r]) + i, s);
      }
    }
  }
  dp_barrier(p, 1);
}

inline int launch_dp_allreduce(const DpArgs& p, int blocks, cudaStream_t st) {
  if (p.world < 1 || p.world > kDpMaxRanks) DIC_FAIL(-4, "dp_allreduce: world size %d not in 1..%d", p.world, kDpMaxRanks);
  if (p.rank < 0 || p.rank >= p.world) DIC_FAIL(-4, "dp_allreduce: rank %d outside world %d", p.rank, p.world);
  if (p.n4 <= 0 || p.n4 % p.world) DIC_FAIL(-4, "dp_allreduce: %lld float4 elements do not split over %d ranks", p.n4, p.world);
  const long long per = p.n4 / p.world;
  int nb = (int)((per + (long long)kDpThreads * 2 - 1) / ((long long)kDpThreads * 2));
  if (blocks <= 0) blocks = 64;
  if (blocks > kDpMaxBlocks) blocks = kDpMaxBlocks;
  if (nb > blocks) nb = blocks;
  if (nb < 1) nb = 1;
  // every rank launches the same grid (per is the same everywhere): the per-block barriers pair up
  if (p.mc) DIC_CUDA(launch_pdl(dp_allreduce_kernel<true, 2>, dim3(nb), dim3(kDpThreads), 0, st, p));
  else if (p.world <= 2) DIC_CUDA(launch_pdl(dp_allreduce_kernel<false, 2>, dim3(nb), dim3(kDpThreads), 0, st, p));
  else if (p.world <= 4) DIC_CUDA(launch_pdl(dp_allreduce_kernel<false, 4>, dim3(nb), dim3(kDpThreads), 0, st, p));
  else DIC_CUDA(launch_pdl(dp_allreduce_kernel<false, 8>, dim3(nb), dim3(kDpThreads), 0, st, p));
  DIC_LAUNCH_CHECK();
  return 0;
}

}  // namespace dic
