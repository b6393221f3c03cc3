// tcgen05 + TMA GEMM engine for sm_100a (bf16 operands, fp32 accumulation in TMEM).
//
//   C[M,N] = act( A . B^T (+ bias[N]) ),  A(m,k), B(n,k) each K-major or MN-major in global memory
//
// Persistent, warp-specialised kernel: one CTA per SM walks the (split, m, n) tile list.
//   warp 0    : TMA producer  (cp.async.bulk.tensor.2d, 128B swizzle, 4-stage mbarrier ring that keeps
//               running across tiles)
//   warp 1    : MMA issuer    (one elected thread, tcgen05.mma.cta_group::1.kind::f16, UMMA 128 x BN x 16,
//               two TMEM accumulator stages so tile i+1 is computed while tile i drains)
//   warp 2    : TMEM allocator (2*BN columns)
//   warps 4-11: epilogue      (tcgen05.ld 32x32b -> bias / sigmoid / cast -> global; the TMEM stage is
//               released as soon as the accumulator is in registers, before the global stores)
// Out-of-range rows/columns/K are zero-filled by TMA, so there are no tail cases in the main loop.
// Split-K: atomic fp32 accumulation or per-split partial buffers (GemmArgs.split_mode).
//
// Operand layouts (DESIGN.md "GEMM engine"):
//   K-major  : row-major [MN, K]; TMA box = 64 K-elements (128 B) x rows; UMMA desc SBO = 1024 B
//   MN-major : row-major [K, MN]; TMA box = 64 MN-elements (128 B) x 64 K rows per 64-wide MN block;
//              UMMA desc SBO = 1024 B (next 8 K rows), LBO = 8192 B (next 64-wide MN block)
// Descriptor field layouts follow the PTX ISA tcgen05 tables (cf. cute/arch/mma_sm100_desc.hpp).
#pragma once
#include <cuda.h>

#include <type_traits>

#include "gemm_generic.cuh"

namespace dic {

constexpr int kTcBM = 128;
constexpr int kTcBK = 64;       // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int kTcStages = 4;
constexpr int kTcEpiWarps = 8;
constexpr int kTcThreads = 128 + 32 * kTcEpiWarps;   // 384
constexpr uint32_t kSpinLimit = 1u << 27;            // turn a would-be hang into a trap

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > kSpinLimit) __trap();
  }
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// store of one [rows, 64 bf16] box (128B-swizzled in shared memory) to global memory; rows / columns outside the
// tensor are clipped by the TMA unit
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tm), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row atoms 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address        bits [0,14)
  d |= (uint64_t)0 << 16;                               // leading byte offset  bits [16,30) (unused: one atom along K)
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset   bits [32,46)
  d |= (uint64_t)1 << 46;                               // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
  return d;
}

// MN-major, 128B-swizzled operand tile as TMA lays it down from a row-major [K, MN] matrix:
// box = 64 MN-elements (128 bytes) x BK rows; atom = 64 MN x 8 K = 1024 bytes; the next 8 K rows
// are 1024 bytes further (SBO), the next block of 64 MN elements is a separate box BK*128 bytes
// further (LBO).
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((kTcBK * 128) >> 4) << 16;            // leading byte offset: next 64-wide MN block
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset: next 8 K rows
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// kind::f16 instruction descriptor: D=f32, A=B=bf16, M x N; a_mn / b_mn select MN-major operands.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, bool a_mn = false, bool b_mn = false) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

struct TcArgs {
  void* C;
  const float* bias;
  int M, N, K;
  long long ldc;
  int c_bf16;
  int splits, split_mode;
  long long split_stride;
  float alpha;
  int sig_lo, sig_hi;
  int tiles_m, tiles_n;
  int tag;
  int fast_act;
  int b_static;
  int c_tma;             // bf16 C leaves through shared memory + cp.async.bulk.tensor stores (BN = 128 only)
  int dbg;               // DIC_GEMM_DEBUG=1: CTA 0 prints a globaltimer breakdown of its first tile
  TraceRec* trace;
};

template <int BN>
constexpr size_t tc_smem_bytes() {
  return 1024 /*align slack*/ + (size_t)kTcStages * (kTcBM * kTcBK * 2 + BN * kTcBK * 2) + 256 /*barriers*/ +
         (size_t)kTcEpiWarps * 32 * 36 * sizeof(float) /*epilogue staging*/;
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const TcArgs p) {
  static_assert(BN == 32 || BN == 64 || BN == 128, "BN");
  static_assert(!B_MN || BN >= 64, "MN-major B needs whole 64-wide blocks");
  extern __shared__ uint8_t smem_raw[];
  Trace trace(p.trace);
  __shared__ unsigned long long dbg_t[12];
#ifdef DIC_GEMM_DEBUG_BUILD
#ifdef DIC_GEMM_DEBUG_BUILD
  const bool dbg = p.dbg == 1 && blockIdx.x == 0;
#else
  constexpr bool dbg = false;      // build with -DDIC_GEMM_DEBUG_BUILD for the DIC_GEMM_DEBUG=1 latency print
#endif
#else
  constexpr bool dbg = false;      // build with -DDIC_GEMM_DEBUG_BUILD for the DIC_GEMM_DEBUG=1 latency print
#endif
  if (dbg && threadIdx.x == 0) dbg_t[0] = gtimer();
  constexpr uint32_t A_BYTES = kTcBM * kTcBK * 2;
  constexpr uint32_t B_BYTES = BN * kTcBK * 2;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN;              // two accumulator stages (64/128/256: powers of 2)
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + kTcStages * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kTcStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kTcStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kTcStages + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kTcStages + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  // per-epilogue-warp staging tiles [32][36] fp32 (store transpose), after the barrier block
  float* stage_base = reinterpret_cast<float*>(smem_raw + (bar_base + 256 - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (p.K + kTcBK - 1) / kTcBK;
  const int tiles_mn = p.tiles_m * p.tiles_n;
  const int total = tiles_mn * p.splits;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
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
  // tile index -> (split, m block, n block); n fastest so that concurrent CTAs share the A tile in L2
  auto decode = [&](int t, int& split, int& m0, int& n0, int& kb0, int& kb1) {
    split = t / tiles_mn;
    const int r = t - split * tiles_mn;
    m0 = (r / p.tiles_n) * kTcBM;
    n0 = (r % p.tiles_n) * BN;
    kb0 = (int)(((long long)num_kb * split) / p.splits);
    kb1 = (int)(((long long)num_kb * (split + 1)) / p.splits);
  };
  // Static B operand (a packed weight): the B halves of the first tile's first stages are requested
  // BEFORE the dependency wait, so their latency overlaps the predecessor's tail; the A halves follow
  // after the wait and complete the same mbarrier transaction counts.
  int pre_b = 0;
  if (p.b_static && threadIdx.x == 0 && (int)blockIdx.x < total) {
    int split, m0, n0, kb0, kb1;
    decode(blockIdx.x, split, m0, n0, kb0, kb1);
    for (int kb = kb0; kb < kb1 && pre_b < kTcStages; ++kb, ++pre_b) {
      const uint32_t sb = base + pre_b * STAGE_BYTES + A_BYTES;
      mbar_expect_tx(full_bar(pre_b), STAGE_BYTES);
      if (B_MN) {
#pragma unroll
        for (int h = 0; h < BN / 64; ++h)
          tma_load_2d(sb + h * (kTcBK * 128), &tmB, full_bar(pre_b), n0 + 64 * h, kb * kTcBK);
      } else {
        tma_load_2d(sb, &tmB, full_bar(pre_b), kb * kTcBK, n0);
      }
    }
  }
  if (dbg && threadIdx.x == 0) dbg_t[1] = gtimer();
  // Everything above (barrier init, TMEM allocation, tensor-map prefetch) overlapped the previous
  // kernel's tail; operands and the output buffer may only be touched after it has completed.
  pdl_wait();
  pdl_trigger();
  trace.mark();
  if (dbg && threadIdx.x == 0) dbg_t[2] = gtimer();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        int split, m0, n0, kb0, kb1;
        decode(t, split, m0, n0, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb) {
          const bool b_done = pre_b > 0;       // this stage's B half (and expect_tx) went out before the wait
          if (b_done) --pre_b;
          const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          if (!b_done) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          }
          if (A_MN) {
#pragma unroll
            for (int h = 0; h < kTcBM / 64; ++h)
              tma_load_2d(sa + h * (kTcBK * 128), &tmA, full_bar(stage), m0 + 64 * h, kb * kTcBK);
          } else {
            tma_load_2d(sa, &tmA, full_bar(stage), kb * kTcBK, m0);
          }
          if (b_done) {
          } else if (B_MN) {
#pragma unroll
            for (int h = 0; h < BN / 64; ++h)
              tma_load_2d(sb + h * (kTcBK * 128), &tmB, full_bar(stage), n0 + 64 * h, kb * kTcBK);
          } else {
            tma_load_2d(sb, &tmB, full_bar(stage), kb * kTcBK, n0);
          }
          if (++stage == kTcStages) { stage = 0; phase ^= 1; }
        }
        if (dbg && t == (int)blockIdx.x) dbg_t[3] = gtimer();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTcBM, BN, A_MN, B_MN);
      // K advance of one UMMA (16 elements): 32 bytes inside the atom row for K-major operands,
      // 16 rows of 128 bytes for MN-major operands (address field is in 16-byte units)
      constexpr uint32_t a_kstep = A_MN ? (16 * 128) >> 4 : 32 >> 4;
      constexpr uint32_t b_kstep = B_MN ? (16 * 128) >> 4 : 32 >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        int split, m0, n0, kb0, kb1;
        decode(t, split, m0, n0, kb0, kb1);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);     // epilogue has drained this accumulator stage
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          if (dbg && t == (int)blockIdx.x && kb == kb0) dbg_t[4] = gtimer();
          if (dbg && t == (int)blockIdx.x && kb == kb1 - 1) dbg_t[5] = gtimer();
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          const uint64_t adesc = A_MN ? umma_desc_mnmajor_sw128(sa) : umma_desc_kmajor_sw128(sa);
          const uint64_t bdesc = B_MN ? umma_desc_mnmajor_sw128(sb) : umma_desc_kmajor_sw128(sb);
#pragma unroll
          for (int k = 0; k < kTcBK / 16; ++k) {
            umma_bf16(tmem_d, adesc + a_kstep * k, bdesc + b_kstep * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));   // frees the smem slot when these MMAs retire
          if (++stage == kTcStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));       // accumulator complete
        if (dbg && t == (int)blockIdx.x) dbg_t[6] = gtimer();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> global =====
    const int ew = warp - 4;
    const int q = warp & 3;                           // TMEM lane quadrant this warp may read
    constexpr int GROUPS = BN >= 64 ? 2 : 1;          // column halves handled by different warps
    constexpr int COLS = BN / GROUPS;
    const int grp = ew >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      int split, m0, n0, kb0, kb1;
      decode(t, split, m0, n0, kb0, kb1);
      // the bias of this thread's store columns is fetched while the accumulator is still being computed
      float bzc[COLS / 32][4];
      {
        const bool pre_bias = p.bias != nullptr && split == 0 && grp < GROUPS;
#pragma unroll
        for (int c = 0; c < COLS / 32; ++c) {
          const int n = n0 + grp * COLS + c * 32 + (lane & 7) * 4;
#pragma unroll
          for (int e = 0; e < 4; ++e) bzc[c][e] = (pre_bias && n + e < p.N) ? __ldg(p.bias + n + e) : 0.f;
        }
      }
      // TMA-store path: lane l keeps the bias of columns l and 32 + l of the warp's 64 (broadcast by shuffle later)
      float bl[2] = {0.f, 0.f};
      if (BN == 128 && p.c_tma && p.bias != nullptr) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int n = n0 + grp * COLS + c * 32 + lane;
          bl[c] = n < p.N ? __ldg(p.bias + n) : 0.f;
        }
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      if (dbg && t == (int)blockIdx.x && warp == 4 && lane == 0) dbg_t[7] = gtimer();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[COLS / 32][32];
      if (grp < GROUPS) {
#pragma unroll
        for (int c = 0; c < COLS / 32; ++c)
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + grp * COLS + c * 32), r[c]);
        tmem_ld_wait();
      }
      // accumulator is in registers: hand the TMEM stage back before touching global memory
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (dbg && t == (int)blockIdx.x && warp == 4 && lane == 0) dbg_t[8] = gtimer();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }

      // Stores go through a per-warp shared-memory transpose: after tcgen05.ld a thread owns one ROW
      // of the tile (32 consecutive columns), which would make every store instruction touch 32
      // different 128-byte lines.  Staged through smem ([32][36] fp32, 16-byte accesses both ways,
      // conflict free), one warp instruction writes 4 full 128-byte row segments.  The epilogue was
      // instruction bound before this (ncu: 15k warp instructions per 64 KB tile), so everything
      // here is 128-bit wide and the bias / activation math is skipped when not requested.
      if constexpr (BN == 128) {
        if (p.c_tma) {
          // bf16 C through shared memory and one bulk tensor store per warp and tile: the warp's [32 rows x 64
          // columns] block is laid down as 32 rows of 128 bytes in the 128B-swizzle pattern (16-byte chunk k of
          // row r at chunk k ^ (r & 7): the 8 lanes of a quarter-warp hit 8 different chunks, conflict free) and
          // leaves as ONE cp.async.bulk.tensor instruction instead of 16 8-byte store instructions per thread,
          // each of which touched four half-written 128-byte lines (logits GEMM: 48 us for 102 MB before).
          const uint32_t stg = bar_base + 1024 + (uint32_t)ew * 4096u;
          if (lane == 0) tma_store_wait_read();        // the previous tile's store has drained this buffer
          __syncwarp();
          const bool has_bias = p.bias != nullptr;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                v[e] = __uint_as_float(r[c][k * 8 + e]);
                if (has_bias) v[e] += __shfl_sync(0xffffffffu, bl[c], k * 8 + e);
              }
              uint4 pk;
              __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
              for (int e = 0; e < 4; ++e) h2[e] = __floats2bfloat162_rn(v[2 * e], v[2 * e + 1]);
              const uint32_t chunk = (uint32_t)((c * 4 + k) ^ (lane & 7));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (uint32_t)lane * 128u + chunk * 16u),
                           "r"(pk.x), "r"(pk.y), "r"(pk.z), "r"(pk.w) : "memory");
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          const int nb = n0 + grp * COLS, mrow0 = m0 + q * 32;
          if (lane == 0 && nb < p.N && mrow0 < p.M) {
            tma_store_2d(&tmC, stg, nb, mrow0);
            tma_store_commit();
          }
          if (dbg && t == (int)blockIdx.x && warp == 4 && lane == 0) dbg_t[9] = gtimer();
          continue;
        }
      }
      if (grp < GROUPS) {
        float* stg = stage_base + ew * (32 * 36);
        const int mrow0 = m0 + q * 32;
        size_t coff = 0;
        if (p.splits > 1 && p.split_mode == 1) coff = (size_t)split * p.split_stride;
        const bool atomic = p.splits > 1 && p.split_mode == 0;
        const bool add_bias = p.bias != nullptr && split == 0;
        const bool plain = !add_bias && p.sig_hi <= p.sig_lo && p.alpha == 1.f;
        const int rows_valid = min(32, p.M - mrow0);
        const int esz = p.c_bf16 ? 2 : 4;
        const bool vec_ok = ((p.ldc * esz) % 16 == 0) && ((reinterpret_cast<uintptr_t>(p.C) + coff * esz) % 16 == 0);
        const int rrow = lane >> 3, col4 = (lane & 7) * 4;
#pragma unroll
        for (int c = 0; c < COLS / 32; ++c) {
          const int nb = n0 + grp * COLS + c * 32;
          if (nb >= p.N || rows_valid <= 0) continue;      // warp-uniform
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(stg + lane * 36 + j) = make_uint4(r[c][j], r[c][j + 1], r[c][j + 2], r[c][j + 3]);
          __syncwarp();
          const int n = nb + col4;
          char* crow = reinterpret_cast<char*>(p.C) + (coff + (size_t)(mrow0 + rrow) * p.ldc + n) * esz;
          const size_t rstep = (size_t)4 * p.ldc * esz;
          const bool full = vec_ok && (n + 3 < p.N);
          // bias / activation are per COLUMN, and after the transpose a thread owns 4 fixed columns.
          // The loop body is specialised on warp-uniform (MODE, OUT) outside the 8 iterations: with the
          // choices tested per element the epilogue more than doubled its instruction count and the
          // logits GEMM (bias) ran at 137 us against 58 us without a bias (ncu launch list, profiles/).
          float bz[4] = {0.f, 0.f, 0.f, 0.f};
          bool sg[4] = {false, false, false, false};
          bool any_sig = false;
          if (!plain) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              bz[e] = bzc[c][e];
              sg[e] = (n + e >= p.sig_lo) && (n + e < p.sig_hi);
            }
            any_sig = (nb + 32 > p.sig_lo) && (nb < p.sig_hi);          // warp-uniform
          }
          // Few specialisations on purpose: MODE 0 = scale + bias (alpha = 1, bias = 0 when not requested:
          // four FMAs per store are free next to the store itself), MODE 1 adds the sigmoid columns;
          // OUT 0/1/2 = 16-byte fp32 / 8-byte bf16 / vector red.add, OUT 3 = scalar tail (any mode).
          // With all 12 (mode, output) pairs unrolled the BN = 128 kernel was 120 KB of code and its
          // epilogue ran 30% slower from instruction-cache misses alone.
          const float al = p.alpha;
          const int outk = (!full || (any_sig && (atomic || p.c_bf16))) ? 3 : (atomic ? 2 : (p.c_bf16 ? 1 : 0));
          auto run = [&](auto MODE, auto OUT) {
#pragma unroll(OUT.value == 3 ? 1 : 8)
            for (int it = 0; it < 8; ++it) {
              const int rr = it * 4 + rrow;
              float4 v = *reinterpret_cast<const float4*>(stg + rr * 36 + col4);
              v.x = fmaf(v.x, al, bz[0]); v.y = fmaf(v.y, al, bz[1]);
              v.z = fmaf(v.z, al, bz[2]); v.w = fmaf(v.w, al, bz[3]);
              if constexpr (MODE.value == 1) {
                // this engine only ever runs in bf16 mode (tc_gemm_eligible): ex2.approx / rcp.approx sigmoid
                if (sg[0]) v.x = sigmoidf_fast(v.x);
                if (sg[1]) v.y = sigmoidf_fast(v.y);
                if (sg[2]) v.z = sigmoidf_fast(v.z);
                if (sg[3]) v.w = sigmoidf_fast(v.w);
              }
              if (rr < rows_valid && n < p.N) {
                if constexpr (OUT.value == 0) {
                  *reinterpret_cast<float4*>(crow) = v;
                } else if constexpr (OUT.value == 1) {
                  uint2 pk;
                  __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
                  h2[0] = __floats2bfloat162_rn(v.x, v.y);
                  h2[1] = __floats2bfloat162_rn(v.z, v.w);
                  *reinterpret_cast<uint2*>(crow) = pk;
                } else if constexpr (OUT.value == 2) {
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(crow), "f"(v.x), "f"(v.y),
                               "f"(v.z), "f"(v.w) : "memory");
                } else {
                  const float ve[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    if (n + e < p.N) {
                      if (atomic) atomicAdd(reinterpret_cast<float*>(crow) + e, ve[e]);
                      else if (p.c_bf16) reinterpret_cast<bf16*>(crow)[e] = __float2bfloat16_rn(ve[e]);
                      else reinterpret_cast<float*>(crow)[e] = ve[e];
                    }
                  }
                }
              }
              crow += rstep;
            }
          };
          using I0 = std::integral_constant<int, 0>;
          using I1 = std::integral_constant<int, 1>;
          using I2 = std::integral_constant<int, 2>;
          using I3 = std::integral_constant<int, 3>;
          if (outk == 0) {
            if (any_sig) run(I1{}, I0{}); else run(I0{}, I0{});
          } else if (outk == 1) {
            run(I0{}, I1{});
          } else if (outk == 2) {
            run(I0{}, I2{});
          } else {
            run(I1{}, I3{});
          }
          __syncwarp();
        }
      }
      if (dbg && t == (int)blockIdx.x && warp == 4 && lane == 0) dbg_t[9] = gtimer();
    }
    if (BN == 128 && p.c_tma && lane == 0) tma_store_wait_all();     // bulk stores complete before the CTA retires
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
  if (dbg && threadIdx.x == 0) {
    const unsigned long long t0 = dbg_t[0];
    printf("[gemm dbg] M=%d N=%d K=%d BN=%d grid=%d | setup %llu wait %llu tma_issued %llu full0 %llu fullN %llu commit %llu tfull %llu tmem_ld %llu stores %llu exit %llu (ns from entry)\n",
           p.M, p.N, p.K, BN, (int)gridDim.x, dbg_t[1] - t0, dbg_t[2] - t0, dbg_t[3] - t0, dbg_t[4] - t0, dbg_t[5] - t0,
           dbg_t[6] - t0, dbg_t[7] - t0, dbg_t[8] - t0, dbg_t[9] - t0, gtimer() - t0);
  }
  trace.end(TK_GEMM_TC + 100 * p.tag);
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled tmap_encoder() {
  static PFN_tmapEncodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  return fn;
}

// 2-D bf16 tensor map over a row-major [rows, cols] matrix with row stride ld (elements);
// box = [box_rows, 64 cols] (64 bf16 = 128 bytes), 128B swizzle, zero fill out of bounds.
inline int make_tmap_bf16(CUtensorMap* tm, const void* ptr, long long rows, long long cols, long long ld,
                          int box_rows, CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B) {
  PFN_tmapEncodeTiled enc = tmap_encoder();
  if (!enc) DIC_FAIL(-6, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) DIC_FAIL(-6, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return 0;
}

inline bool tc_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DIC_DISABLE_TC");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

// DIC_TMA_STORE=0: bf16 outputs take the per-thread store epilogue (A/B switch of the measurement)
inline bool tc_tma_store_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DIC_TMA_STORE"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

inline int tc_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

// operand layouts the engine takes: K-major (k stride 1) or MN-major (m/n stride 1); the other
// stride must keep TMA's 16-byte global stride rule
inline bool tc_operand_ok(const void* p, long long s_mn, long long s_k) {
  if (reinterpret_cast<uintptr_t>(p) & 15) return false;
  if (s_k == 1) return s_mn % 8 == 0 && s_mn > 0;
  if (s_mn == 1) return s_k % 8 == 0 && s_k > 0;
  return false;
}

inline bool tc_gemm_eligible(const GemmArgs& g) {
  if (!tc_enabled()) return false;
  if (!g.a_bf16 || !g.b_bf16) return false;
  if (!tc_operand_ok(g.A, g.a_m, g.a_k) || !tc_operand_ok(g.B, g.b_n, g.b_k)) return false;
  if (g.batch != 1 || g.accumulate) return false;
  if (g.K < 16 || g.M < 1 || g.N < 8) return false;
  if ((long long)g.M * g.N < 128LL * 128LL) return false;   // tiny outputs: the FMA engine is as fast
  if (g.splits > 1 && g.split_mode == 0 && (g.c_bf16 || g.sig_hi > g.sig_lo)) return false;
  return true;
}

template <int BN, bool A_MN, bool B_MN>
inline int tc_gemm_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const TcArgs& p,
                          int grid, cudaStream_t st) {
  static DeviceOnce attr_set;
  if (int dev_ = 0; attr_set.need(&dev_)) {
    DIC_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)tc_smem_bytes<BN>()));
    attr_set.mark(dev_);
  }
  DIC_CUDA(launch_pdl(tc_gemm_kernel<BN, A_MN, B_MN>, dim3(grid), dim3(kTcThreads), tc_smem_bytes<BN>(), st, tmA, tmB,
                      tmC, p));
  DIC_LAUNCH_CHECK();
  return 0;
}

template <int BN>
inline int tc_gemm_bn(const GemmArgs& g, cudaStream_t st) {
  const bool a_mn = g.a_k != 1, b_mn = g.b_k != 1;
  CUtensorMap tmA, tmB;
  // K-major: matrix [MN rows, K cols], box MN x 64(K).  MN-major: matrix [K rows, MN cols], box 64(K) x 64(MN).
  if (a_mn) DIC_TRY(make_tmap_bf16(&tmA, g.A, g.K, g.M, g.a_k, kTcBK));
  else      DIC_TRY(make_tmap_bf16(&tmA, g.A, g.M, g.K, g.a_m, kTcBM));
  if (b_mn) DIC_TRY(make_tmap_bf16(&tmB, g.B, g.K, g.N, g.b_k, kTcBK));
  else      DIC_TRY(make_tmap_bf16(&tmB, g.B, g.N, g.K, g.b_n, BN));
  TcArgs p;
  p.C = g.C; p.bias = g.bias; p.M = g.M; p.N = g.N; p.K = g.K; p.ldc = g.ldc; p.c_bf16 = g.c_bf16;
  const int num_kb = cdiv(g.K, kTcBK);
  p.splits = g.splits < num_kb ? g.splits : num_kb;
  if (p.splits < 1) p.splits = 1;
  if (g.splits > 1 && g.split_mode == 1 && p.splits != g.splits)
    DIC_FAIL(-4, "tc_gemm: partial-buffer split-K needs splits <= K/64");
  p.split_mode = g.split_mode; p.split_stride = g.split_stride;
  p.alpha = g.alpha; p.sig_lo = g.sig_lo; p.sig_hi = g.sig_hi;
  p.tiles_m = cdiv(g.M, kTcBM);
  p.tiles_n = cdiv(g.N, BN);
  p.tag = g.tag;
  p.fast_act = g.fast_act;
  p.b_static = g.b_static;
  p.trace = g_trace_host;
  // bf16 output of a plain (optionally biased) product: bulk tensor stores (see the epilogue)
  CUtensorMap tmC = tmA;
  p.c_tma = 0;
  if (BN == 128 && g.c_bf16 && p.splits == 1 && g.alpha == 1.f && g.sig_hi <= g.sig_lo && g.ldc % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(g.C) & 15) == 0 && tc_tma_store_enabled()) {
    DIC_TRY(make_tmap_bf16(&tmC, g.C, g.M, g.N, g.ldc, 32));
    p.c_tma = 1;
  }
  {
    static int dbg_env = -1;
    if (dbg_env < 0) { const char* e = getenv("DIC_GEMM_DEBUG"); dbg_env = (e && e[0] >= '1' && e[0] <= '9') ? e[0] - '0' : 0; }
    p.dbg = dbg_env;
  }
  const long long total = (long long)p.tiles_m * p.tiles_n * p.splits;
  const int grid = (int)(total < tc_num_sms() ? total : tc_num_sms());
  ProfScope prof(g.prof ? g.prof : (int)P_GEMM_TC, st, g.prof_bytes);
  if constexpr (BN >= 64) {
    if (!a_mn && b_mn) return tc_gemm_launch<BN, false, true>(tmA, tmB, tmC, p, grid, st);
    if (a_mn && b_mn) return tc_gemm_launch<BN, true, true>(tmA, tmB, tmC, p, grid, st);
  }
  if (a_mn) return tc_gemm_launch<BN, true, false>(tmA, tmB, tmC, p, grid, st);
  return tc_gemm_launch<BN, false, false>(tmA, tmB, tmC, p, grid, st);
}

// Tile width: 128 columns unless that leaves most SMs idle (few tiles, no split-K), then 64 / 32.
inline int tc_gemm(const GemmArgs& g, cudaStream_t st) {
  const bool b_mn = g.b_k != 1;
  const int num_kb = cdiv(g.K, kTcBK);
  const int splits = g.splits < 1 ? 1 : (g.splits < num_kb ? g.splits : num_kb);
  const long long tm = cdiv(g.M, kTcBM);
  const int half = tc_num_sms() / 2;
  int bn = 128;
  if (tm * cdiv(g.N, 128) * splits < half && g.N > 64) bn = 64;
  if (bn == 64 && tm * cdiv(g.N, 64) * splits < half && g.N > 32 && !b_mn) bn = 32;
  if (g.bn == 128 || g.bn == 64 || (g.bn == 32 && !b_mn)) bn = g.bn;
  if (bn == 128) return tc_gemm_bn<128>(g, st);
  if (bn == 64) return tc_gemm_bn<64>(g, st);
  return tc_gemm_bn<32>(g, st);
}

}  // namespace dic
