// Beam search, bf16 mode: vocabulary projection + per-slice softmax statistics in ONE tcgen05 kernel, then
// log-sum-exp + two-level top-K + candidate merge + beam reorder in one per-image kernel.  Replaces, per step,
//   logits GEMM (9.0 us) -> beam_row_topk (15.4 us: stages all 10000 logits of a row, K block-wide rounds) ->
//   beam_merge (2.4 us) -> beam_reorder (1.2 us)   [+ three launch gaps; profiles/r01_timeline_beam_v17.txt]
// Specification: oracle/decoder_oracle.py beam_search / beam_select.  The fp32 parity mode
// and every call that asks for the logits / lse traces keep the unfused kernels of decode.cuh; this path takes
// the same fp32 logits, the same candidate arithmetic (score + (x - lse), plain fp32 add/sub) and the same order
// (value descending, flat index ascending), so it differs from them only through the summation order of the
// row's log-sum-exp.
//
// (1) beam_logits_stats_kernel, grid (vocabulary slices of 80 columns, chunks of <= 640 rows):
//     the CTA's weight slice W_out[n0:n0+80, :] (20 KB, static: requested before the dependency wait) and ALL
//     its row tiles (h' of 128 rows x 128, 32 KB each) go to shared memory by TMA; one thread issues
//     8 x tcgen05.mma (128 x 80 x 16) per tile into its own TMEM accumulator (5 x 96 columns); eight epilogue
//     warps read an accumulator as one row per thread (tcgen05.ld 32x32b), add the bias, store the fp32 logits
//     and the slice's (max, sum exp(x - max)) pair:  stats[slice][row] = { m, s }.
//     A first version also took the slice's top-K in this epilogue (K arg-max rounds over the thread's 80
//     registers): 27.9 us per launch, instruction bound (80 x K x ~8 instructions per row and slice).
// (2) beam_select_reorder_kernel<K>, one CTA per image, warp j = beam row j.  Two-level selection: the K-th largest
//     slice maximum is a lower bound of the row's K-th largest logit, so only slices whose maximum reaches it
//     (K of them, more only on exact ties; at most 8 are taken) can hold a top-K element -- the warp reads those
//     <= 8 x 80 logits instead of 10000.  Then: candidates score + (x - lse) (finished rows: <end> at cost 0),
//     K warp arg-best rounds, warp 0 merges the K x K row candidates into the new scores / backpointers / tokens /
//     finished flags, and the whole CTA gathers (h, c) by backpointer and embeds the chosen tokens.
#pragma once
#include "decode.cuh"
#include "gemm_tc.cuh"

namespace dic {

constexpr int kBfNB = 80;             // vocabulary columns per CTA (UMMA N; 10000 = 125 x 80)
constexpr int kBfTileCols = 96;       // TMEM columns reserved per accumulator
constexpr int kBfMaxTiles = 5;        // 128-row tiles per CTA (5 x 96 <= 512 TMEM columns, 5 x 32 KB of shared memory)
constexpr int kBfEpiWarps = 8;
constexpr int kBfThreads = 128 + 32 * kBfEpiWarps;
constexpr int kBfH = 128;             // decoder width this kernel is built for (K of the GEMM)
constexpr int kBfMaxSlicesPerLane = 4;

constexpr int kBfStageLd = 84;         // floats per staged row (80 + 4: conflict-free 16-byte accesses both ways)
constexpr int kBfSel = 8;             // slices a row may draw candidates from (K + exact ties of slice maxima)

constexpr size_t kBfBBytes = (size_t)2 * kBfNB * 128;                  // two 64-wide K blocks of the weight slice
constexpr size_t kBfABytes = (size_t)2 * 128 * 128;                    // one row tile
inline size_t beam_logits_smem_bytes(int tiles) {
  return 1024 + kBfBBytes + (size_t)tiles * kBfABytes + 256 + sizeof(float) * kBfNB;
}

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

struct BeamLogitsArgs {
  const float* bias;     // [V]
  float* part;           // [slices][rows] (max, sum-exp) pairs
  float* logits;         // [rows, V] fp32
  int rows, V, slices;
  TraceRec* trace;
};

__global__ void __launch_bounds__(kBfThreads, 1)
beam_logits_stats_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const BeamLogitsArgs p) {
  extern __shared__ uint8_t smem_raw[];
  Trace trace(p.trace);
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int slice = blockIdx.x, chunk = blockIdx.y;
  const int n0 = slice * kBfNB;
  const int row_base = chunk * (kBfMaxTiles * 128);
  const int tiles = min(kBfMaxTiles, (p.rows - row_base + 127) / 128);
  const uint32_t sB = base, sA = base + (uint32_t)kBfBBytes;
  const uint32_t bar_base = sA + (uint32_t)(kBfMaxTiles * kBfABytes);
  auto bfull = [&]() { return bar_base; };
  auto afull = [&](int t) { return bar_base + 8u * (1 + t); };
  auto tfull = [&](int t) { return bar_base + 8u * (1 + kBfMaxTiles + t); };
  const uint32_t tmem_slot = bar_base + 8u * (1 + 2 * kBfMaxTiles);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  float* bias_s = reinterpret_cast<float*>(smem_raw + (bar_base + 256 - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    mbar_init(bfull(), 1);
    for (int t = 0; t < kBfMaxTiles; ++t) { mbar_init(afull(t), 1); mbar_init(tfull(t), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp == 3) {
    for (int i = lane; i < kBfNB; i += 32) bias_s[i] = (n0 + i < p.V) ? p.bias[n0 + i] : 0.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot_ptr;

  // the weight slice is static: in flight before the dependency wait (rows past V are zero-filled by TMA)
  if (threadIdx.x == 0) {
    mbar_expect_tx(bfull(), (uint32_t)kBfBBytes);
    tma_load_2d(sB, &tmB, bfull(), 0, n0);
    tma_load_2d(sB + kBfNB * 128, &tmB, bfull(), 64, n0);
  }
  pdl_wait();          // h' comes from the LSTM kernel of this step
  pdl_trigger();
  trace.mark();

  if (warp == 0) {
    if (lane == 0) {
      for (int t = 0; t < tiles; ++t) {
        const uint32_t sa = sA + (uint32_t)(t * kBfABytes);
        mbar_expect_tx(afull(t), (uint32_t)kBfABytes);
        tma_load_2d(sa, &tmA, afull(t), 0, row_base + t * 128);
        tma_load_2d(sa + 128 * 128, &tmA, afull(t), 64, row_base + t * 128);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, kBfNB);
      mbar_wait(bfull(), 0);
      for (int t = 0; t < tiles; ++t) {
        mbar_wait(afull(t), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = sA + (uint32_t)(t * kBfABytes);
        const uint32_t tmem_d = tmem_base + (uint32_t)(t * kBfTileCols);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t adesc = umma_desc_kmajor_sw128(sa + kb * (128 * 128));
          const uint64_t bdesc = umma_desc_kmajor_sw128(sB + kb * (kBfNB * 128));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(tfull(t));
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4, q = warp & 3;
    bool staged_ok = false;
    const bool full_slice = n0 + kBfNB <= p.V;          // CTA-uniform
    static_assert(kBfNB / 4 == 20, "store pattern below: 5 x 32 lanes = 8 rows of 20 float4");
    int rb[5], cb[5], so[5];
#pragma unroll
    for (int b = 0; b < 5; ++b) {
      const int idx = b * 32 + lane;
      rb[b] = idx / 20;
      cb[b] = idx - rb[b] * 20;
      so[b] = rb[b] * kBfStageLd + cb[b] * 4;
    }
    for (int t = ew >> 2; t < tiles; t += 2) {
      mbar_wait(tfull(t), 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float v[kBfNB];
      {
        uint32_t vr[kBfNB];          // all five loads in flight, one wait
#pragma unroll
        for (int c = 0; c < kBfNB / 16; ++c)
          tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * kBfTileCols + c * 16),
                             *reinterpret_cast<uint32_t(*)[16]>(&vr[c * 16]));
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < kBfNB; ++j) v[j] = __uint_as_float(vr[j]);
      }
      const int row = row_base + t * 128 + q * 32 + lane;
      // This epilogue is instruction-issue bound (ncu: 33 k warp instructions per CTA at 35 % issue activity = the
      // kernel's 12 us): the common case -- a slice that lies inside the vocabulary, a full 128-row tile -- runs
      // without per-element masks, reads the bias as float4, and takes exp(x - m) as ONE fused multiply-add into
      // ex2.approx (__expf costs three multiplies, a compare and the ex2).
      float m = -INFINITY;
      if (full_slice) {
#pragma unroll
        for (int j = 0; j < kBfNB; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + j);
          v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
          m = fmaxf(m, fmaxf(fmaxf(v[j], v[j + 1]), fmaxf(v[j + 2], v[j + 3])));
        }
      } else {
#pragma unroll
        for (int j = 0; j < kBfNB; ++j) {
          v[j] = (n0 + j < p.V) ? v[j] + bias_s[j] : -INFINITY;
          m = fmaxf(m, v[j]);
        }
      }
      constexpr float kLog2e = 1.4426950408889634f;
      const float m2 = m * kLog2e;
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int j = 0; j < kBfNB; j += 4) {
        s0 += ex2_approx(fmaf(v[j], kLog2e, -m2)); s1 += ex2_approx(fmaf(v[j + 1], kLog2e, -m2));
        s2 += ex2_approx(fmaf(v[j + 2], kLog2e, -m2)); s3 += ex2_approx(fmaf(v[j + 3], kLog2e, -m2));
      }
      if (row < p.rows)
        reinterpret_cast<float2*>(p.part)[(size_t)slice * p.rows + row] = make_float2(m, (s0 + s1) + (s2 + s3));
      // Logits out through a per-warp shared-memory transpose (a thread owns a ROW: its own 16-byte stores hit 32
      // different lines per instruction -- ncu: 1.6 M store sectors, the LSU was the bottleneck).  The staging
      // area re-uses the row tiles' shared memory, so it waits for the LAST tile's MMAs (issued in order).
      if (!staged_ok) { mbar_wait(tfull(tiles - 1), 0); staged_ok = true; }
      float* stg = reinterpret_cast<float*>(smem_raw + (sA - smem_u32(smem_raw))) + ew * (32 * kBfStageLd);
#pragma unroll
      for (int j = 0; j < kBfNB; j += 4)
        *reinterpret_cast<float4*>(stg + lane * kBfStageLd + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      __syncwarp();
      const int wrow0 = row_base + t * 128 + q * 32;
      const bool all_in = full_slice && wrow0 + 32 <= p.rows;        // warp-uniform: no bounds checks in the store loop
      float* gdst = p.logits + (size_t)wrow0 * p.V + n0;
      // 640 float4 per warp tile, 32 lanes: the (row, chunk) pattern of a lane repeats every 5 instructions with the
      // row advanced by 8, so the five shared-memory offsets are per-kernel constants and a store is one add away
      // from the five per-tile pointers (the linear index / 20 of the first version was a third of the loop)
      if (all_in) {
        float* g5[5];
#pragma unroll
        for (int b = 0; b < 5; ++b) g5[b] = gdst + (size_t)rb[b] * p.V + cb[b] * 4;
        const size_t step8 = (size_t)8 * p.V;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 5; ++b)
            *reinterpret_cast<float4*>(g5[b] + a * step8) =
                *reinterpret_cast<const float4*>(stg + so[b] + a * (8 * kBfStageLd));
      } else {
#pragma unroll 4
        for (int it = 0; it < kBfNB / 4; ++it) {
          const int idx = it * 32 + lane;
          const int rr = idx / (kBfNB / 4), c4 = idx - rr * (kBfNB / 4);
          const float4 o = *reinterpret_cast<const float4*>(stg + rr * kBfStageLd + c4 * 4);
          if (wrow0 + rr < p.rows && n0 + c4 * 4 + 3 < p.V)
            *reinterpret_cast<float4*>(gdst + (size_t)rr * p.V + c4 * 4) = o;
        }
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  trace.end(TK_GEMM_TC + 500);
}

struct BeamMergeArgs {
  const float* part;           // [slices][rows] (max, sum-exp) pairs
  const float* logits;         // [rows, V]
  const float* scores;         // [rows]
  const uint8_t* finished;     // [rows]
  float* new_scores;           // [rows]
  int32_t* back;               // [rows]
  int32_t* tok;                // [rows]
  uint8_t* new_finished;       // [rows]
  const void* h_tmp;           // ST [rows, H]
  const float* c_tmp;          // [rows, H]
  const void* emb;             // ST [V, E]
  void* Xnext;                 // ST rows of the next step's input: [emb | . | h]
  long long x_row;
  int col_h;
  float* c;                    // [rows, H]
  int rows, V, slices, end_id, E, H;
  TraceRec* trace;
};

template <typename ST, int K>
__global__ void __launch_bounds__(256, 3) beam_select_reorder_kernel(const BeamMergeArgs p) {      // <= 80 registers: 7 context CTAs fit next to it (decode_impl look-ahead)
  constexpr int SPL = kBfMaxSlicesPerLane;
  constexpr int EPL = (kBfNB + 31) / 32;          // logits of one slice per lane
  constexpr int NC = kBfSel * EPL;                // candidates per lane
  __shared__ float cand_v[K * K];
  __shared__ int cand_i[K * K];
  __shared__ int s_back[K], s_tok[K];
  Trace trace(p.trace);
  pdl_wait();
  pdl_trigger();
  trace.mark();
  const int img = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp < K) {
    const int j = warp, r = img * K + j;
    const float sc = p.scores[r];
    const bool fin = p.finished[r] != 0;
    // slice statistics of this row: all loads first (one L2 round trip)
    float2 st[SPL];
    const float2* stats = reinterpret_cast<const float2*>(p.part);
#pragma unroll
    for (int u = 0; u < SPL; ++u) {
      const int s = lane + 32 * u;
      st[u] = stats[(size_t)(s < p.slices ? s : p.slices - 1) * p.rows + r];
    }
    float pm[SPL];
    float m = -INFINITY;
#pragma unroll
    for (int u = 0; u < SPL; ++u) {
      pm[u] = (lane + 32 * u < p.slices) ? st[u].x : -INFINITY;
      m = fmaxf(m, pm[u]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float ssum = 0.f;
#pragma unroll
    for (int u = 0; u < SPL; ++u) ssum += (pm[u] == -INFINITY) ? 0.f : st[u].y * __expf(pm[u] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
    const float ls = m + logf(ssum);
    // The K slices with the largest maxima (maximum descending, slice ascending), known to every lane; slices
    // beyond them whose maximum TIES the K-th one are appended (exact ties only: up to kBfSel in total).
    int sel_s[kBfSel];
#pragma unroll
    for (int k = 0; k < kBfSel; ++k) sel_s[k] = 0x7fffffff;
    float tau = -INFINITY;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      float bm = -INFINITY;
      int bs = 0x7fffffff;
#pragma unroll
      for (int u = 0; u < SPL; ++u) {
        const int s = lane + 32 * u;
        if (s < p.slices && pm[u] != -INFINITY && (bs == 0x7fffffff || pm[u] > bm)) { bm = pm[u]; bs = s; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, bm, o);
        const int os = __shfl_xor_sync(0xffffffffu, bs, o);
        if (os != 0x7fffffff && (bs == 0x7fffffff || om > bm || (om == bm && os < bs))) { bm = om; bs = os; }
      }
      sel_s[k] = bs;
      tau = bm;                            // after the last round: lower bound of the row's K-th largest logit
#pragma unroll
      for (int u = 0; u < SPL; ++u)
        if (lane + 32 * u == bs) pm[u] = -INFINITY;
    }
    if (!fin) {
#pragma unroll
      for (int k = K; k < kBfSel; ++k) {   // warp-uniform early exit: no slice left at tau (the usual case)
        int ts = 0x7fffffff;
#pragma unroll
        for (int u = 0; u < SPL; ++u)
          if (lane + 32 * u < p.slices && pm[u] == tau && pm[u] != -INFINITY && lane + 32 * u < ts) ts = lane + 32 * u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ts = min(ts, __shfl_xor_sync(0xffffffffu, ts, o));
        if (ts == 0x7fffffff) break;
        sel_s[k] = ts;
#pragma unroll
        for (int u = 0; u < SPL; ++u)
          if (lane + 32 * u == ts) pm[u] = -INFINITY;
      }
    }
    // candidates: the logits of the selected slices (flat index j*V + token); a lane keeps its best and its
    // runner-up so that most arg-best rounds need no rescan of the lane's list
    float cv[NC];
    int ci[NC];
    const float* lrow = p.logits + (size_t)r * p.V;
#pragma unroll
    for (int k = 0; k < kBfSel; ++k) {
      const bool use = !fin && sel_s[k] != 0x7fffffff;       // warp-uniform
#pragma unroll
      for (int e = 0; e < EPL; ++e) {
        const int col = sel_s[k] * kBfNB + lane + 32 * e;
        const bool ok = use && (lane + 32 * e < kBfNB) && col < p.V;
        cv[k * EPL + e] = ok ? __ldcg(lrow + col) : -INFINITY;
        ci[k * EPL + e] = ok ? j * p.V + col : 0x7fffffff;
      }
    }
    if (fin && lane == 0) {
      // a finished row keeps only <end> at cost 0; the fillers are the lowest other tokens at -inf
      cv[0] = sc; ci[0] = j * p.V + p.end_id;
#pragma unroll
      for (int k = 1; k < K; ++k) { cv[k] = -INFINITY; ci[k] = j * p.V + ((k - 1) < p.end_id ? (k - 1) : k); }
    }
    float bv = -INFINITY, b2v = -INFINITY;
    int bi = 0x7fffffff, b2i = 0x7fffffff;
#pragma unroll
    for (int x = 0; x < NC; ++x) {
      if (ci[x] == 0x7fffffff) continue;
      if (!fin) cv[x] = __fadd_rn(sc, __fsub_rn(cv[x], ls));
      if (bi == 0x7fffffff || cand_better(cv[x], ci[x], bv, bi)) { b2v = bv; b2i = bi; bv = cv[x]; bi = ci[x]; }
      else if (b2i == 0x7fffffff || cand_better(cv[x], ci[x], b2v, b2i)) { b2v = cv[x]; b2i = ci[x]; }
    }
    float tv = INFINITY;       // last candidate taken from this lane's list
    int ti = -1;
#pragma unroll 1
    for (int rd = 0; rd < K; ++rd) {
      float wv = bv;
      int wi = bi;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
        if (oi != 0x7fffffff && (wi == 0x7fffffff || cand_better(ov, oi, wv, wi))) { wv = ov; wi = oi; }
      }
      if (lane == 0) { cand_v[j * K + rd] = wv; cand_i[j * K + rd] = wi; }
      if (bi == wi && bi != 0x7fffffff) {       // this lane's candidate was taken: promote the runner-up or rescan
        tv = bv; ti = bi;
        if (b2i != 0x7fffffff) { bv = b2v; bi = b2i; b2i = 0x7fffffff; }
        else {
          bv = -INFINITY; bi = 0x7fffffff;
#pragma unroll
          for (int x = 0; x < NC; ++x) {
            const bool elig = ci[x] != 0x7fffffff && (cv[x] < tv || (cv[x] == tv && ci[x] > ti));
            if (elig && (bi == 0x7fffffff || cand_better(cv[x], ci[x], bv, bi))) { bv = cv[x]; bi = ci[x]; }
          }
        }
      }
    }
  }
  __syncthreads();
  if (warp == 0) {
    constexpr int N = K * K;
    float v0 = -INFINITY, v1 = -INFINITY;
    int i0 = 0x7fffffff, i1 = 0x7fffffff;
    if (lane < N) { v0 = cand_v[lane]; i0 = cand_i[lane]; }
    if (lane + 32 < N) { v1 = cand_v[lane + 32]; i1 = cand_i[lane + 32]; }
    const int fin_l = (lane < K) ? (int)p.finished[img * K + lane] : 0;      // lane l holds row l's flag
    for (int rd = 0; rd < K; ++rd) {
      float bv = v0;
      int bi = i0;
      if (cand_better(v1, i1, bv, bi)) { bv = v1; bi = i1; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
      }
      const int jb = bi / p.V, tk = bi - jb * p.V;
      const int was_fin = __shfl_sync(0xffffffffu, fin_l, jb & 31);
      if (lane == 0) {
        p.new_scores[img * K + rd] = bv;
        p.back[img * K + rd] = jb;
        p.tok[img * K + rd] = tk;
        p.new_finished[img * K + rd] = (uint8_t)((was_fin != 0) || (tk == p.end_id));
        s_back[rd] = jb;
        s_tok[rd] = tk;
      }
      if (i0 == bi && bi != 0x7fffffff) { v0 = -INFINITY; i0 = 0x7fffffff; }
      if (i1 == bi && bi != 0x7fffffff) { v1 = -INFINITY; i1 = 0x7fffffff; }
    }
  }
  __syncthreads();
  // reorder: next-step rows of this image (decode.cuh beam_reorder_kernel); not in the look-ahead order, where the
  // next LSTM kernel follows the backpointers itself (Xnext == null)
  if (p.Xnext == nullptr) { trace.end(TK_BEAM_MERGE); return; }
  const int Wd = p.E + p.H;
  const ST* h_tmp = reinterpret_cast<const ST*>(p.h_tmp);
  const ST* emb = reinterpret_cast<const ST*>(p.emb);
  ST* Xn = reinterpret_cast<ST*>(p.Xnext);
  for (int i = tid; i < K * Wd; i += 256) {
    const int rr = i / Wd, qd = i - rr * Wd;
    const int r = img * K + rr;
    if (qd < p.E) {
      Xn[(size_t)r * p.x_row + qd] = emb[(size_t)s_tok[rr] * p.E + qd];
    } else {
      const int jj = qd - p.E;
      const int src = img * K + s_back[rr];
      Xn[(size_t)r * p.x_row + p.col_h + jj] = h_tmp[(size_t)src * p.H + jj];
      p.c[(size_t)r * p.H + jj] = p.c_tmp[(size_t)src * p.H + jj];
    }
  }
  trace.end(TK_BEAM_MERGE);
}

inline bool beam_fused_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DIC_BEAM_FUSED"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

// bf16 storage, H = 128, K in {3, 5} instantiated
inline bool beam_fused_eligible(int H, int V, int K, int rows) {
  if (!beam_fused_enabled() || !tc_enabled()) return false;
  if (H != kBfH || (K != 3 && K != 5)) return false;
  const int slices = cdiv(V, kBfNB);
  if (slices > 32 * kBfMaxSlicesPerLane) return false;
  if (V % 4) return false;                                           // 16-byte logits stores
  return rows > 0;
}

// the two launches of a fused beam step; the look-ahead order of decode_impl puts the next step's head kernel between them
inline int launch_beam_logits_stats(const void* h_tmp, long long h_ld, const void* w_out, const float* b_out, float* logits,
                                    float* part, int rows, int V, cudaStream_t st) {
  const int slices = cdiv(V, kBfNB);
  const int chunks = cdiv(rows, kBfMaxTiles * 128);
  CUtensorMap tmA, tmB;
  DIC_TRY(make_tmap_bf16(&tmA, h_tmp, rows, kBfH, h_ld, 128));
  DIC_TRY(make_tmap_bf16(&tmB, w_out, V, kBfH, kBfH, kBfNB));
  static DeviceOnce attr_set;
  if (int dev_ = 0; attr_set.need(&dev_)) {
    DIC_CUDA(cudaFuncSetAttribute(beam_logits_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)beam_logits_smem_bytes(kBfMaxTiles)));
    attr_set.mark(dev_);
  }
  BeamLogitsArgs la;
  la.bias = b_out; la.part = part; la.logits = logits; la.rows = rows; la.V = V; la.slices = slices; la.trace = g_trace_host;
  ProfScope prof(P_BEAM_SELECT, st, (double)rows * V * 2);
  DIC_CUDA(launch_pdl(beam_logits_stats_kernel, dim3(slices, chunks), dim3(kBfThreads),
                      beam_logits_smem_bytes(kBfMaxTiles), st, tmA, tmB, la));
  DIC_LAUNCH_CHECK();
  return 0;
}

template <typename ST>
inline int launch_beam_select_reorder(float* logits, float* part, BeamMergeArgs mg, int rows, int V, int K, cudaStream_t st) {
  mg.part = part; mg.logits = logits; mg.rows = rows; mg.V = V; mg.slices = cdiv(V, kBfNB); mg.trace = g_trace_host;
  switch (K) {
    case 3: DIC_CUDA(launch_pdl(beam_select_reorder_kernel<ST, 3>, dim3(rows / K), dim3(256), 0, st, mg)); break;
    case 5: DIC_CUDA(launch_pdl(beam_select_reorder_kernel<ST, 5>, dim3(rows / K), dim3(256), 0, st, mg)); break;
    default: DIC_FAIL(-4, "beam_fused: beam %d not instantiated", K);
  }
  DIC_LAUNCH_CHECK();
  return 0;
}

template <typename ST>
inline int launch_beam_fused(const void* h_tmp, const void* w_out, const float* b_out, float* logits, float* part,
                             const BeamMergeArgs& mg, int rows, int V, int K, cudaStream_t st) {
  DIC_TRY(launch_beam_logits_stats(h_tmp, kBfH, w_out, b_out, logits, part, rows, V, st));
  return launch_beam_select_reorder<ST>(logits, part, mg, rows, V, K, st);
}

}  // namespace dic
