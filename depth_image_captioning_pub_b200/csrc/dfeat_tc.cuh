// dL/dF as ONE tcgen05 GEMM (bf16 mode).
//
//   dF[b] (L x D) = datt1[b] (L x A) . W_enc (A x D)  +  alpha[b]^T (L x T) . dz[b] (T x D)  +  dmeanF[b] / L
//
// i.e. per image a contraction over K = A + T of the concatenated operands [datt1 | alpha^T] and
// [W_enc ; dz].  The first version ran the datt1.W_enc GEMM into an fp32 [B,L,D] accumulator (411 MB
// written) and a second kernel re-read it to add the T rank-1 updates and cast to bf16 (ncu: 108 us +
// 265 us); here the accumulator never leaves TMEM and dF is written once, in bf16.
//
// Tiling: per image, ceil(L/128) M tiles of 128 rows (rows past L are computed on whatever the boxes
// pick up and masked at the store), N over D (128 columns).  Tiles do not straddle images: the
// alpha box would then start at l0 = m0 - b*L, and TMA needs the innermost box coordinate 16-byte
// aligned (an unaligned start raises an illegal-instruction fault; found the hard way).
// Per tile the K loop is
//   segment 1: ceil(A/64) blocks   A-operand datt1 (K-major),        B-operand W_enc (MN-major)
//   segment 2: ceil(T/64) blocks   A-operand alpha16[b] as [t, l] (MN-major; columns past Lp zero-filled),
//              B-operand dz viewed as [T, B*D] (image b at columns b*D..): rows t >= T are zero-filled,
//              which also cancels whatever the alpha box picked up from the next image's rows.
// Both segments accumulate into the same TMEM tile (the instruction descriptor's A-major bit differs).
// Warp roles and the smem ring are those of gemm_tc.cuh.
#pragma once
#include "gemm_tc.cuh"

namespace dic {

struct DfeatArgs {
  bf16* dF;              // [B*L, D] bf16
  const float* dmeanF;   // [B, D] fp32 (d mean_l F; enters dF as dmeanF / L)
  float inv_l;
  int B, L, D, A, T;
  int tiles_m, tiles_n;
  TraceRec* trace;
};

// ring as in gemm_tc.cuh + per-warp epilogue staging tiles of [32][68] fp32
constexpr size_t dfeat_smem_bytes() {
  return 1024 + (size_t)kTcStages * (kTcBM * kTcBK * 2 + 128 * kTcBK * 2) + 256 + (size_t)kTcEpiWarps * 32 * 68 * sizeof(float);
}

__global__ void __launch_bounds__(kTcThreads, 1)
dfeat_gemm_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                  const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                  const DfeatArgs p) {
  constexpr int BN = 128;
  extern __shared__ uint8_t smem_raw[];
  Trace trace(p.trace);
  constexpr uint32_t A_BYTES = kTcBM * kTcBK * 2;
  constexpr uint32_t B_BYTES = BN * kTcBK * 2;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + kTcStages * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kTcStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kTcStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kTcStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kTcStages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* stage_base = reinterpret_cast<float*>(smem_raw + (bar_base + 256 - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KA = (p.A + kTcBK - 1) / kTcBK, KT = (p.T + kTcBK - 1) / kTcBK;
  const int total = p.tiles_m * p.tiles_n;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB2) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kTcEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();
  pdl_trigger();
  trace.mark();

  // n fastest: the CTAs working at the same time share the tile's A operands (datt1, alpha) in L2
  const int tpi = (p.L + kTcBM - 1) / kTcBM;      // M tiles per image
  auto decode = [&](int t, int& m0, int& n0, int& b, int& l0) {
    const int mt = t / p.tiles_n;
    n0 = (t % p.tiles_n) * BN;
    b = mt / tpi;
    l0 = (mt - b * tpi) * kTcBM;
    m0 = b * p.L + l0;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        int m0, n0, b, l0;
        decode(t, m0, n0, b, l0);
        for (int kb = 0; kb < KA; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          tma_load_2d(sa, &tmA1, full_bar(stage), kb * kTcBK, m0);
#pragma unroll
          for (int h = 0; h < BN / 64; ++h)
            tma_load_2d(sb + h * (kTcBK * 128), &tmB1, full_bar(stage), n0 + 64 * h, kb * kTcBK);
          if (++stage == kTcStages) { stage = 0; phase ^= 1; }
        }
        {
          for (int kb = 0; kb < KT; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
            mbar_expect_tx(full_bar(stage), STAGE_BYTES);
#pragma unroll
            for (int h = 0; h < kTcBM / 64; ++h)
              tma_load_2d(sa + h * (kTcBK * 128), &tmA2, full_bar(stage), l0 + 64 * h, b * p.T + kb * kTcBK);
#pragma unroll
            for (int h = 0; h < BN / 64; ++h)
              tma_load_2d(sb + h * (kTcBK * 128), &tmB2, full_bar(stage), b * p.D + n0 + 64 * h, kb * kTcBK);
            if (++stage == kTcStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(kTcBM, BN, false, true);
      constexpr uint32_t idesc2 = umma_idesc_bf16(kTcBM, BN, true, true);
      constexpr uint32_t k_kmajor = 32 >> 4, k_mnmajor = (16 * 128) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        int m0, n0, b, l0;
        decode(t, m0, n0, b, l0);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        const int nkb = KA + KT;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          const bool seg1 = kb < KA;
          const uint64_t adesc = seg1 ? umma_desc_kmajor_sw128(sa) : umma_desc_mnmajor_sw128(sa);
          const uint64_t bdesc = umma_desc_mnmajor_sw128(sb);
          const uint32_t ak = seg1 ? k_kmajor : k_mnmajor;
          const uint32_t idesc = seg1 ? idesc1 : idesc2;
#pragma unroll
          for (int k = 0; k < kTcBK / 16; ++k)
            umma_bf16(tmem_d, adesc + ak * k, bdesc + k_mnmajor * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(empty_bar(stage));
          if (++stage == kTcStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int q = warp & 3;
    constexpr int COLS = BN / 2;
    const int grp = ew >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      int m0, n0, b, l0;
      decode(t, m0, n0, b, l0);
      mbar_wait(tfull_bar(acc), acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[COLS / 32][32];
#pragma unroll
      for (int c = 0; c < COLS / 32; ++c)
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + grp * COLS + c * 32), r[c]);
      tmem_ld_wait();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }

      // Per-warp staging tile [32 rows][64 + 4] fp32 holds the warp's whole 64-column half, so a store instruction
      // writes 16 bytes per lane = full 128-byte row segments (8-byte stores of 64-byte segments left this
      // kernel bound by its epilogue at ~2 TB/s of bf16 output).
      float* stg = stage_base + ew * (32 * 68);
      const int mrow0 = m0 + q * 32;
      const int rows_valid = min(32, p.L - (l0 + q * 32));      // rows of THIS image in the warp's 32-row slab
      const int nb = n0 + grp * COLS;
      if (nb < p.D && rows_valid > 0) {
#pragma unroll
        for (int c = 0; c < COLS / 32; ++c)
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(stg + lane * 68 + c * 32 + j) = make_uint4(r[c][j], r[c][j + 1], r[c][j + 2], r[c][j + 3]);
        __syncwarp();
        const int rrow = lane >> 3, col8 = (lane & 7) * 8;
        const int n = nb + col8;
        if (n < p.D) {           // D % 8 == 0: a lane's 8 columns are in or out together
          const float4 mb0 = __ldg(reinterpret_cast<const float4*>(p.dmeanF + (size_t)b * p.D + n));
          const float4 mb1 = __ldg(reinterpret_cast<const float4*>(p.dmeanF + (size_t)b * p.D + n + 4));
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + rrow;
            if (rr < rows_valid) {
              const float4 v0 = *reinterpret_cast<const float4*>(stg + rr * 68 + col8);
              const float4 v1 = *reinterpret_cast<const float4*>(stg + rr * 68 + col8 + 4);
              uint4 pk;
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
              h2[0] = __floats2bfloat162_rn(fmaf(mb0.x, p.inv_l, v0.x), fmaf(mb0.y, p.inv_l, v0.y));
              h2[1] = __floats2bfloat162_rn(fmaf(mb0.z, p.inv_l, v0.z), fmaf(mb0.w, p.inv_l, v0.w));
              h2[2] = __floats2bfloat162_rn(fmaf(mb1.x, p.inv_l, v1.x), fmaf(mb1.y, p.inv_l, v1.y));
              h2[3] = __floats2bfloat162_rn(fmaf(mb1.z, p.inv_l, v1.z), fmaf(mb1.w, p.inv_l, v1.w));
              *reinterpret_cast<uint4*>(p.dF + (size_t)(mrow0 + rr) * p.D + n) = pk;
            }
          }
        }
        __syncwarp();
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
  trace.end(TK_MISC);
}

// (An A-stationary variant -- the image's A tiles resident in shared memory, every B tile feeding both row tiles:
// 2.4x less operand traffic through TMA -- was built and measured in rounds 1-2: 107 us against 94-100 us for the
// kernel above at the benchmark shape; removed.  With TMA loads AND the per-thread bf16 stores both near 1.7 us per
// 128 x 128 tile the pass needs both fixed at once: DESIGN.md section 4, "bf16 GEMM outputs through TMA stores".)

inline bool dfeat_tc_eligible(int A, int D, int Lp) {
  return tc_enabled() && A % 8 == 0 && D % 8 == 0 && Lp % 8 == 0;
}

// datt1 [B*L, A] bf16; Wenc [A, D] bf16; alpha16 [B*T, Lp] bf16 (columns >= L and inactive rows zero);
// dz [T, B, D] bf16 (inactive rows zero); dmeanF [B, D] fp32; dF [B*L, D] bf16
inline int launch_dfeat_tc(const bf16* datt1, const bf16* Wenc, const bf16* alpha16, int Lp, const bf16* dz,
                           const float* dmeanF, bf16* dF, int B, int L, int D, int A, int T,
                           cudaStream_t st, int reserve_sms = 0) {
  // reserve_sms: SMs left free for a gradient all-reduce running next to this (persistent, one CTA per SM) kernel
  const int max_ctas = tc_num_sms() - reserve_sms > 8 ? tc_num_sms() - reserve_sms : 8;
  CUtensorMap tmA1, tmB1, tmA2, tmB2;
  DIC_TRY(make_tmap_bf16(&tmA1, datt1, (long long)B * L, A, A, kTcBM));          // K-major: [rows, A], box 128 x 64
  DIC_TRY(make_tmap_bf16(&tmB1, Wenc, A, D, D, kTcBK));                          // MN-major: [A rows, D], box 64 x 64
  DIC_TRY(make_tmap_bf16(&tmA2, alpha16, (long long)B * T, Lp, Lp, kTcBK));      // MN-major: [(b,t) rows, Lp], box 64 x 64
  // dz [T][B][D] seen as a row-major [T, B*D] matrix: image b's columns start at b*D, rows t >= T are
  // out of bounds (zero-filled); columns past D belong to the next image and are masked at the store
  DIC_TRY(make_tmap_bf16(&tmB2, dz, T, (long long)B * D, (long long)B * D, kTcBK));
  DfeatArgs p;
  p.dF = dF; p.dmeanF = dmeanF; p.inv_l = 1.f / (float)L; p.B = B; p.L = L; p.D = D; p.A = A; p.T = T;
  p.tiles_m = B * cdiv(L, kTcBM);
  p.tiles_n = cdiv(D, 128);
  p.trace = g_trace_host;
  const long long total = (long long)p.tiles_m * p.tiles_n;
  const int grid = (int)(total < max_ctas ? total : max_ctas);
  static DeviceOnce attr_set;
  if (int dev_ = 0; attr_set.need(&dev_)) {
    DIC_CUDA(cudaFuncSetAttribute(dfeat_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)dfeat_smem_bytes()));
    attr_set.mark(dev_);
  }
  ProfScope prof(P_DFEAT, st, (double)B * L * D * 2);
  DIC_CUDA(launch_pdl(dfeat_gemm_kernel, dim3(grid), dim3(kTcThreads), dfeat_smem_bytes(), st, tmA1, tmB1, tmA2,
                      tmB2, p));
  DIC_LAUNCH_CHECK();
  return 0;
}

}  // namespace dic
