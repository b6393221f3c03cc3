// Fused "head" of a decoder timestep (bf16 storage, reference shape A = H = 128):
//
//   att2  = h . W_dec^T  + b_dec                       attention.py:85
//   beta  = sigmoid(h . W_beta^T + b_beta)             depth_models.py:189
//   e[l]  = relu(att1[l,:] + att2) . w_full + b_full   attention.py:86-87
//   alpha = softmax_L(e) | gumbel variants             attention.py:90, :12-48
//
// in ONE launch instead of the h-projection GEMM followed by the alpha kernel.  Both were pure latency
// chains (5.3 us + 6.0 us + a launch gap per timestep at 256 captions, in-kernel timeline of round 1)
// around a few MFLOP: the GEMM CTA waits for its operands to come through TMA, commits one 128-deep MMA
// and drains a 2 MB fp32 output; the alpha kernel then waits for that output to re-read 512 bytes of it.
//
// Work split: a group of 16 rows (one m16 tile of warp-level MMA) is served by 8 CTAs.  CTA `part` of a
// group computes beta[16 rows, D/8 columns] on the tensor cores (mma.sync.m16n8k16, the whole weight slice
// of the CTA is register resident: it is static, so it is requested BEFORE the programmatic-dependency
// wait, like the two att1 slabs that go to shared memory through cp.async), att2 for all 16 rows (16 more
// n-tiles, 8x redundant across the parts: 0.5 MFLOP) and then energies + normalisation for rows
// 2*part and 2*part+1 of the group.
//
// The K index of the MMA is permuted so that operands load as 16-byte vectors straight from their
// row-major global layout: a thread's fragment registers for the virtual k-tile (c, half) hold elements
// 32c + 8*tq + 4*half + {0,1} and + {2,3} of a row; A and B use the same permutation, the sum is the same.
#pragma once
#include "attention.cuh"
#include "attention_mma.cuh"

namespace dic {

constexpr int kHeadThreads = 512;
constexpr int kHeadParts = 8;                                // CTAs per group
constexpr int kHeadDim = 128;                                // A = H = 128
// KB = rows (beams) per image.  KB == 1 (training, greedy decode): a group is 16 rows (one m16 tile), a CTA
// owns attention rows 2*part, 2*part+1 (two att1 slabs).  KB > 1 (beam search): a group is 8 images = 8*KB rows
// (ceil(8*KB/16) m16 tiles against the same register-resident weight fragments) and CTA `part` owns the KB
// rows of image 8*group + part, which share ONE att1 slab.
__host__ __device__ constexpr int head_group_rows(int KB) { return KB == 1 ? 16 : 8 * KB; }
__host__ __device__ constexpr int head_cta_rows(int KB) { return KB == 1 ? 2 : KB; }
__host__ __device__ constexpr int head_slabs(int KB) { return KB == 1 ? 2 : 1; }

struct HeadArgs {
  const bf16* h;          // [rows, H], row r at h + r*h_ld
  long long h_ld;
  const bf16* Wdb;        // [A + D, H]: rows W_dec | W_beta  (Pack::Wdb)
  const float* bias_db;   // [A + D] fp32: b_dec | b_beta
  float* HP;              // [rows, A + D] fp32 out: att2 | beta
  AttnFwdArgs a;          // att1, w_full, b_full, u, alpha_out / alpha16_out, mode, inv_temp, L, D, A
  int rows;
};

inline size_t attn_head_smem_bytes(int L, int KB) {
  const size_t Lp = (size_t)(L + 3) & ~(size_t)3;
  return (size_t)head_slabs(KB) * L * kHeadDim * 2 +
         sizeof(float) * ((size_t)head_cta_rows(KB) * kHeadDim + kHeadDim + (size_t)head_cta_rows(KB) * Lp);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ uint4 ldg_nc16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
// h was written by the preceding kernel of the chain: read through L2 (no stale L1 line of an older step)
__device__ __forceinline__ uint4 ldg_cg16(const void* p) {
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// NT = beta n-tiles (8 columns) per warp: D = 8 parts * 16 warps * NT * 8
template <int NT, int KB>
__global__ void __launch_bounds__(kHeadThreads, 1) attn_head_kernel(const HeadArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int A = kHeadDim, H = kHeadDim;
  constexpr int GROUP = head_group_rows(KB), CROWS = head_cta_rows(KB), SLABS = head_slabs(KB);
  constexpr int MT = (GROUP + 15) / 16;
  Trace trace(p.a.trace);
  const int L = p.a.L, D = p.a.D;
  const int Lp = (L + 3) & ~3;
  bf16* att1_s = reinterpret_cast<bf16*>(smem_raw);                         // [SLABS][L][A]
  float* att2_s = reinterpret_cast<float*>(att1_s + (size_t)SLABS * L * A);   // [CROWS][A]
  float* w_s = att2_s + CROWS * A;                                          // [A]
  float* e_s = w_s + A;                                                     // [CROWS][Lp]

  const int part = blockIdx.x, grp = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tq = lane & 3;
  const int row_base = grp * GROUP;
  const int arow0 = row_base + part * CROWS;       // first attention row of this CTA

  // ---- static operands, requested before the dependency wait ------------------------------------------
  // (1) the att1 slab(s) of this CTA's attention rows -> shared memory
#pragma unroll
  for (int r = 0; r < SLABS; ++r) {
    const int row = arow0 + r;
    if (row < p.rows) {
      const int img = row / KB;
      const bf16* src = reinterpret_cast<const bf16*>(p.a.att1) + (size_t)img * L * A;
      bf16* dst = att1_s + (size_t)r * L * A;
      for (int i = tid; i < L * A / 8; i += kHeadThreads) cp_async16(dst + i * 8, src + i * 8);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  // (2) weight fragments: NT beta tiles + 1 att2 tile per warp, 4 x 16 bytes per tile (the whole K = 128)
  const int ncol0 = part * (D / kHeadParts) + warp * NT * 8;   // first beta column of this warp
  uint4 bw[NT + 1][4];
#pragma unroll
  for (int jj = 0; jj <= NT; ++jj) {
    const int wrow = jj < NT ? A + ncol0 + jj * 8 + gid : warp * 8 + gid;
#pragma unroll
    for (int c = 0; c < 4; ++c) bw[jj][c] = ldg_nc16(p.Wdb + (size_t)wrow * H + 32 * c + 8 * tq);
  }
  // (3) biases of this thread's accumulator columns, the scoring vector
  float2 bz[NT + 1];
#pragma unroll
  for (int jj = 0; jj <= NT; ++jj) {
    const int col = jj < NT ? A + ncol0 + jj * 8 + 2 * tq : warp * 8 + 2 * tq;
    bz[jj] = *reinterpret_cast<const float2*>(p.bias_db + col);
  }
  if (tid < A) w_s[tid] = p.a.w_full[tid];
  const float b_full = p.a.b_full[0];

  pdl_wait();
  pdl_trigger();
  trace.mark();

#pragma unroll 1
  for (int mt = 0; mt < MT; ++mt) {
    // ---- h rows of this m16 tile (A fragments) -----------------------------------------------------------
    const int g_lo = mt * 16 + gid, g_hi = g_lo + 8;               // row index inside the group
    const int r_lo = row_base + g_lo, r_hi = row_base + g_hi;
    const bool ok_lo = g_lo < GROUP && r_lo < p.rows, ok_hi = g_hi < GROUP && r_hi < p.rows;
    uint4 ha[4], hb[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      ha[c] = ok_lo ? ldg_cg16(p.h + (size_t)r_lo * p.h_ld + 32 * c + 8 * tq) : make_uint4(0u, 0u, 0u, 0u);
      hb[c] = ok_hi ? ldg_cg16(p.h + (size_t)r_hi * p.h_ld + 32 * c + 8 * tq) : make_uint4(0u, 0u, 0u, 0u);
    }
    float acc[NT + 1][4];
#pragma unroll
    for (int jj = 0; jj <= NT; ++jj)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[jj][q] = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int jj = 0; jj <= NT; ++jj) {
        mma_bf16_16816(acc[jj], ha[c].x, hb[c].x, ha[c].y, hb[c].y, bw[jj][c].x, bw[jj][c].y);
        mma_bf16_16816(acc[jj], ha[c].z, hb[c].z, ha[c].w, hb[c].w, bw[jj][c].z, bw[jj][c].w);
      }
    }
    // ---- att2 (all parts compute it; each keeps its own rows and writes them out) ------------------------
    {
      const float v0 = acc[NT][0] + bz[NT].x, v1 = acc[NT][1] + bz[NT].y;     // row g_lo
      const float v2 = acc[NT][2] + bz[NT].x, v3 = acc[NT][3] + bz[NT].y;     // row g_hi
      const int a0 = warp * 8 + 2 * tq;
      const int lr_lo = g_lo - part * CROWS, lr_hi = g_hi - part * CROWS;
      if (lr_lo >= 0 && lr_lo < CROWS) {
        att2_s[lr_lo * A + a0] = v0; att2_s[lr_lo * A + a0 + 1] = v1;
        if (ok_lo) *reinterpret_cast<float2*>(p.HP + (size_t)r_lo * (A + D) + a0) = make_float2(v0, v1);
      }
      if (lr_hi >= 0 && lr_hi < CROWS) {
        att2_s[lr_hi * A + a0] = v2; att2_s[lr_hi * A + a0 + 1] = v3;
        if (ok_hi) *reinterpret_cast<float2*>(p.HP + (size_t)r_hi * (A + D) + a0) = make_float2(v2, v3);
      }
    }
    // ---- beta ----------------------------------------------------------------------------------------------
#pragma unroll
    for (int jj = 0; jj < NT; ++jj) {
      const int col = A + ncol0 + jj * 8 + 2 * tq;
      if (ok_lo)
        *reinterpret_cast<float2*>(p.HP + (size_t)r_lo * (A + D) + col) =
            make_float2(sigmoidf_fast(acc[jj][0] + bz[jj].x), sigmoidf_fast(acc[jj][1] + bz[jj].y));
      if (ok_hi)
        *reinterpret_cast<float2*>(p.HP + (size_t)r_hi * (A + D) + col) =
            make_float2(sigmoidf_fast(acc[jj][2] + bz[jj].x), sigmoidf_fast(acc[jj][3] + bz[jj].y));
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();

  // ---- energies: half-warp per annotation row, 16 bytes (8 columns) per lane -------------------------------
  {
    const int half = lane >> 4, hl = lane & 15;
    float w8[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) w8[q] = w_s[hl * 8 + q];
    constexpr int RPW = 2 * (kHeadThreads / 32);     // 32 rows per CTA pass
    if constexpr (KB == 1) {
#pragma unroll
      for (int r = 0; r < CROWS; ++r) {
        if (arow0 + r >= p.rows) continue;             // CTA-uniform
        float a2[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) a2[q] = att2_s[r * A + hl * 8 + q];
        const bf16* slab = att1_s + (size_t)r * L * A;
        for (int lb = 0; lb < L; lb += RPW) {      // warp-uniform trip count: every lane runs the shuffles
          const int l = lb + warp * 2 + half;
          float s = 0.f;
          if (l < L) {
            const uint4 raw = *reinterpret_cast<const uint4*>(slab + (size_t)l * A + hl * 8);
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = __bfloat1622float2(h2[i]);
              s = fmaf(fmaxf(f.x + a2[2 * i], 0.f), w8[2 * i], s);
              s = fmaf(fmaxf(f.y + a2[2 * i + 1], 0.f), w8[2 * i + 1], s);
            }
          }
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          if (hl == 0 && l < L) e_s[r * Lp + l] = s + b_full;
        }
      }
    } else if (arow0 < p.rows) {
      // the KB rows of one image against one read of its att1 slab
      for (int lb = 0; lb < L; lb += RPW) {
        const int l = lb + warp * 2 + half;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = 0.f;
        if (l < L) {
          const uint4 raw = *reinterpret_cast<const uint4*>(att1_s + (size_t)l * A + hl * 8);
          const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(h2[i]);
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
          }
        }
#pragma unroll
        for (int j = 0; j < KB; ++j) {
          float s = 0.f;
#pragma unroll
          for (int q = 0; q < 8; ++q) s = fmaf(fmaxf(v[q] + att2_s[j * A + hl * 8 + q], 0.f), w8[q], s);
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          if (hl == 0 && l < L) e_s[j * Lp + l] = s + b_full;
        }
      }
    }
  }
  __syncthreads();
  if (warp < CROWS && arow0 + warp < p.rows) attn_normalise_row(p.a, e_s + warp * Lp, arow0 + warp, lane);
  trace.end(TK_ALPHA);
}

inline bool head_env_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DIC_FUSED_HEAD"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

// The fused kernel covers the reference shape in bf16 storage (1, 3 or 5 rows per image); everything else takes
// the h-projection GEMM + alpha kernel.
inline bool attn_head_eligible(int A, int H, int D, int L, int rows_per_image) {
  if (!head_env_enabled()) return false;
  if (A != kHeadDim || H != kHeadDim) return false;
  if (D != kHeadParts * (kHeadThreads / 32) * 8 * 2) return false;     // NT = 2: D = 2048
  if (rows_per_image != 1 && rows_per_image != 3 && rows_per_image != 5) return false;
  return attn_head_smem_bytes(L, rows_per_image) <= 200 * 1024;
}

template <int KB>
inline int launch_attn_head_kb(const HeadArgs& p, cudaStream_t st) {
  static DeviceOnce attr_set;
  if (int dev_ = 0; attr_set.need(&dev_)) {
    DIC_CUDA(cudaFuncSetAttribute(attn_head_kernel<2, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set.mark(dev_);
  }
  ProfScope prof(P_ATTN_ALPHA, st, (double)(p.rows / KB) * p.a.L * p.a.A * 2);
  DIC_CUDA(launch_pdl(attn_head_kernel<2, KB>, dim3(kHeadParts, cdiv(p.rows, head_group_rows(KB))), dim3(kHeadThreads),
                      attn_head_smem_bytes(p.a.L, KB), st, p));
  DIC_LAUNCH_CHECK();
  return 0;
}

inline int launch_attn_head(const HeadArgs& p_in, int KB, cudaStream_t st) {
  if (p_in.rows <= 0) return 0;
  HeadArgs p = p_in;
  p.a.trace = g_trace_host;
  switch (KB) {
    case 1: return launch_attn_head_kb<1>(p, st);
    case 3: return launch_attn_head_kb<3>(p, st);
    case 5: return launch_attn_head_kb<5>(p, st);
    default: DIC_FAIL(-4, "attn_head: %d rows per image not instantiated", KB);
  }
}

}  // namespace dic
