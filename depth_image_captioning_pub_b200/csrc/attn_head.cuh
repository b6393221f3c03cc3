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

constexpr int kHeadSlabLd = kHeadDim + 8;                    // bf16 per staged att1 row: 272 bytes, so that the 16-byte
                                                             // reads of 8 consecutive rows fall into 8 different bank groups
__host__ __device__ constexpr int head_halves(int KB) { return KB == 1 ? 1 : 2; }   // column halves of the energy pass
inline size_t attn_head_smem_bytes(int L, int KB) {
  const size_t Lp = (size_t)(L + 3) & ~(size_t)3;
  return (size_t)head_slabs(KB) * L * kHeadSlabLd * 2 +
         sizeof(float) * ((size_t)head_cta_rows(KB) * kHeadDim + kHeadDim +
                          (size_t)head_halves(KB) * head_cta_rows(KB) * Lp) +
         (size_t)((head_group_rows(KB) + 15) / 16) * 16 * (kHeadDim + 32) * 2;
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
  constexpr int SA = kHeadSlabLd, NH = head_halves(KB);
  bf16* att1_s = reinterpret_cast<bf16*>(smem_raw);                         // [SLABS][L][SA]
  float* att2_s = reinterpret_cast<float*>(att1_s + (size_t)SLABS * L * SA);  // [CROWS][A]
  float* w_s = att2_s + CROWS * A;                                          // [A]
  float* e_s = w_s + A;                                                     // [NH][CROWS][Lp]
  constexpr int HS = kHeadDim + 32;
  bf16* h_s = reinterpret_cast<bf16*>(e_s + NH * CROWS * Lp);                // [MT * 16][HS]

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
      bf16* dst = att1_s + (size_t)r * L * SA;
      for (int i = tid; i < L * A / 8; i += kHeadThreads) cp_async16(dst + (i >> 4) * SA + (i & 15) * 8, src + i * 8);
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
  // (kept in a register across the wait: a shared-memory store here would make the thread sit out the load's
  // round trip BEFORE it reaches the dependency wait)
  const float w_mine = tid < A ? p.a.w_full[tid] : 0.f;
  const float b_full = p.a.b_full[0];

  pdl_wait();
  pdl_trigger();
  trace.mark();
  if (tid < A) w_s[tid] = w_mine;

  // h rows of the group -> shared memory, once per CTA (every warp needs all of them as A fragments; the first
  // version had each warp fetch them from L2 itself, one dependent round trip per m16 tile: 6.4 us for the three
  // tiles of a 5-beam group).  Rows are 320 bytes apart: the fragment reads of a quarter-warp (2 rows x 4 chunks)
  // then fall into 8 different bank groups.
  {
    constexpr int CH = MT * 16 * (H / 8);
#pragma unroll
    for (int i0 = 0; i0 < CH; i0 += kHeadThreads) {
      const int i = i0 + tid;
      if (i < CH) {
        const int g = i >> 4, ck = i & 15, r = row_base + g;
        const uint4 v = (g < GROUP && r < p.rows) ? ldg_cg16(p.h + (size_t)r * p.h_ld + ck * 8) : make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(h_s + g * HS + ck * 8) = v;
      }
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int mt = 0; mt < MT; ++mt) {
    const int g_lo = mt * 16 + gid, g_hi = g_lo + 8;               // row index inside the group
    const int r_lo = row_base + g_lo, r_hi = row_base + g_hi;
    const bool ok_lo = g_lo < GROUP && r_lo < p.rows, ok_hi = g_hi < GROUP && r_hi < p.rows;
    uint4 ha[4], hb[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      ha[c] = *reinterpret_cast<const uint4*>(h_s + g_lo * HS + 32 * c + 8 * tq);
      hb[c] = *reinterpret_cast<const uint4*>(h_s + g_hi * HS + 32 * c + 8 * tq);
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

  // ---- energies: one thread per annotation row (and column half), no cross-lane reduction -------------------
  // e[j][l] = sum_a w[a] relu(att1[l][a] + att2[j][a]).  A warp owns 32 consecutive annotation rows of one slab
  // (KB == 1: two slabs, all 128 columns per thread) or of one column half (KB > 1: 64 columns per thread, all KB
  // rows of the image against one read of the slab); att2 and w come from shared memory as warp-wide broadcasts.
  // (The first version gave a half-warp to each annotation row: 4 shuffle stages per energy and a re-read of att2
  // per pass -- 7.5 us of the kernel's 15 us at 5 beams, profiles/r02_beam_phase_times.txt.)
  {
    constexpr int COMBOS = KB == 1 ? SLABS : NH;
    constexpr int CW = A / NH;                       // columns per thread
    constexpr int NR = KB == 1 ? 1 : KB;             // att2 rows per thread
    const int Lw = (L + 31) >> 5;
    const bool cta_live = arow0 < p.rows;
    for (int it = warp; it < COMBOS * Lw; it += kHeadThreads / 32) {
      const int combo = it / Lw, l = (it - combo * Lw) * 32 + lane;
      const int slab = KB == 1 ? combo : 0, half = KB == 1 ? 0 : combo;
      if (!cta_live || (KB == 1 && arow0 + slab >= p.rows)) continue;      // warp-uniform
      const bf16* rowp = att1_s + ((size_t)slab * L + (l < L ? l : L - 1)) * SA + half * CW;
      const float* a2p = att2_s + (KB == 1 ? slab * A : 0) + half * CW;
      const float* wp = w_s + half * CW;
      float sacc[NR];
#pragma unroll
      for (int j = 0; j < NR; ++j) sacc[j] = 0.f;
#pragma unroll 4
      for (int c = 0; c < CW / 8; ++c) {
        const uint4 raw = *reinterpret_cast<const uint4*>(rowp + c * 8);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
        float v[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(h2[i]);
          v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
        const float4 w0 = *reinterpret_cast<const float4*>(wp + c * 8), w1 = *reinterpret_cast<const float4*>(wp + c * 8 + 4);
        const float w8[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int j = 0; j < NR; ++j) {
          const float4 a0 = *reinterpret_cast<const float4*>(a2p + j * A + c * 8);
          const float4 a1 = *reinterpret_cast<const float4*>(a2p + j * A + c * 8 + 4);
          const float a8[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
          for (int q = 0; q < 8; ++q) sacc[j] = fmaf(fmaxf(v[q] + a8[q], 0.f), w8[q], sacc[j]);
        }
      }
      if (l < L) {
        if constexpr (KB == 1) {
          e_s[slab * Lp + l] = sacc[0] + b_full;
        } else {
#pragma unroll
          for (int j = 0; j < NR; ++j) e_s[(half * CROWS + j) * Lp + l] = sacc[j] + (half == 0 ? b_full : 0.f);
        }
      }
    }
  }
  __syncthreads();
  if (warp < CROWS && arow0 + warp < p.rows) {
    if constexpr (NH == 2) {
      for (int l = lane; l < L; l += 32) e_s[warp * Lp + l] += e_s[(CROWS + warp) * Lp + l];
      __syncwarp();
    }
    attn_normalise_row<true>(p.a, e_s + warp * Lp, arow0 + warp, lane);
  }
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
