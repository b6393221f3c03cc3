// LSTM cell of one decoder step as ONE kernel (bf16 mode):
//   gates = [emb | beta*z | h] . [W_ih | W_hh]^T + (b_ih + b_hh)      (nn.LSTMCell, depth_models.py:122,193)
//   c' = sig(f) c + sig(i) tanh(g) ;  h' = sig(o) tanh(c')            pointwise, in the GEMM's epilogue
//
// The GEMM has M = batch rows, N = 4H = 512, K = E+D+H = 2304: a handful of output tiles with a long
// contraction, so K is split 8 ways -- across the 8 CTAs of a thread-block CLUSTER.  Each CTA runs its K
// slice on tcgen05 into TMEM, parks the fp32 partial tile in its own shared memory, and after a cluster
// barrier every CTA reduces 1/8 of the tile's rows over the 8 partials through distributed shared memory
// (fixed order: deterministic) and applies the cell update to them.  This replaces the split-K GEMM that
// wrote 9 partial tiles to global memory plus a separate pointwise kernel that re-read them
// (4.0 + 1.1 + 3.3 us per step in the timeline) with one launch.
// The pointwise update needs the four gates of a hidden unit in one tile, so the weight pack holds a copy of
// [W_ih | W_hh] with its rows interleaved: 64-column tile nt = units [32 nt, 32 nt + 32) x (i, f, g, o).
#pragma once
#include "gemm_tc.cuh"
#include "lstm.cuh"

namespace dic {

constexpr int kGlSplits = 8;                 // cluster size = K splits
constexpr int kGlUnits = 32;                 // hidden units per N tile
constexpr int kGlBN = 4 * kGlUnits;          // x 4 gates = 128 columns: 2 x 4 tiles at batch 256 = 8 clusters, one per GPC
                                             // (16 clusters of 8 did not all become resident: a second wave, 13 us)
constexpr int kGlThreads = 256;              // warp 0 TMA, 1 MMA, 2 TMEM alloc, 4-7 epilogue; all 8 reduce
constexpr int kGlPartLd = kGlBN + 4;         // padded row of the fp32 partial tile (16-byte aligned rows, conflict-free 128-bit access)

struct GatesLstmArgs {
  LstmFwdArgs l;        // outputs / state of the pointwise part (gate_part, part_stride, splits unused)
  int rows, H, K;       // K = E + D + H
  int tiles_n;          // 4H / 64
  TraceRec* trace;
};

constexpr size_t gates_lstm_smem_bytes() {
  return 1024 + (size_t)kTcStages * (kTcBM * kTcBK * 2 + kGlBN * kTcBK * 2) + 256 + sizeof(float) * kTcBM * kGlPartLd;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ld_dsmem_f32x4(uint32_t local_addr, uint32_t cta) {
  uint32_t ra;
  float4 v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(cta));
  asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ra) : "memory");
  return v;
}

template <typename ST>
__global__ void __launch_bounds__(kGlThreads, 1)
gates_lstm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GatesLstmArgs p) {
  extern __shared__ uint8_t smem_raw[];
  Trace trace(p.trace);
  constexpr uint32_t A_BYTES = kTcBM * kTcBK * 2;
  constexpr uint32_t B_BYTES = kGlBN * kTcBK * 2;
  constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = kGlBN;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + kTcStages * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kTcStages + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * kTcStages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * kTcStages + 1);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const uint32_t part_addr = bar_base + 256;                                     // fp32 [128][kGlPartLd]
  float* part = reinterpret_cast<float*>(smem_raw + (part_addr - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = (int)cluster_ctarank();            // == blockIdx.x (cluster dims (8,1,1))
  const int tile = blockIdx.y;
  const int m0 = (tile / p.tiles_n) * kTcBM;
  const int nt = tile % p.tiles_n;
  const int n0 = nt * kGlBN;
  const int num_kb = (p.K + kTcBK - 1) / kTcBK;
  const int kb0 = (int)(((long long)num_kb * split) / kGlSplits);
  const int kb1 = (int)(((long long)num_kb * (split + 1)) / kGlSplits);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
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

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
        mbar_expect_tx(full_bar(stage), STAGE_BYTES);
        tma_load_2d(sa, &tmA, full_bar(stage), kb * kTcBK, m0);
        tma_load_2d(sb, &tmB, full_bar(stage), kb * kTcBK, n0);
        if (++stage == kTcStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kTcBM, kGlBN, false, false);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
        const uint64_t adesc = umma_desc_kmajor_sw128(sa), bdesc = umma_desc_kmajor_sw128(sb);
#pragma unroll
        for (int k = 0; k < kTcBK / 16; ++k)
          umma_bf16(tmem_base, adesc + 2u * k, bdesc + 2u * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        umma_commit(empty_bar(stage));
        if (++stage == kTcStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(tfull_bar);
    }
  } else if (warp >= 4) {
    // accumulator -> this CTA's partial tile in shared memory (a thread owns one row after tcgen05.ld)
    const int q = warp & 3;
    mbar_wait(tfull_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* prow = part + (size_t)(q * 32 + lane) * kGlPartLd;
    if (kb1 > kb0) {
#pragma unroll
      for (int c = 0; c < kGlBN / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4*>(prow + c * 32 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < kGlBN; j += 4) *reinterpret_cast<float4*>(prow + j) = make_float4(0.f, 0.f, 0.f, 0.f);   // empty K slice
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }

  // every CTA's partial tile is in place
  __syncwarp();
  cluster_sync_all();

  // ---- reduce + LSTM pointwise: CTA `split` owns rows [16 split, 16 split + 16) of the tile ----
  {
    const int H = p.H;
    for (int pi = threadIdx.x; pi < 16 * (kGlUnits / 4); pi += kGlThreads) {   // 16 rows x groups of 4 hidden units
      const int rr = pi / (kGlUnits / 4), ug = pi - rr * (kGlUnits / 4);
      const int trow = split * 16 + rr;
      const int row = m0 + trow;
      const int u0 = nt * kGlUnits + ug * 4;
      if (row >= p.rows || u0 >= H) continue;          // H % kGlUnits == 0: a group is in or out as a whole
      float g4[4][4];                                   // [gate][unit]
#pragma unroll
      for (int qg = 0; qg < 4; ++qg) {
        const uint32_t la = part_addr + (uint32_t)((trow * kGlPartLd + qg * kGlUnits + ug * 4) * 4);
        float4 v[kGlSplits];
#pragma unroll
        for (int c = 0; c < kGlSplits; ++c) v[c] = ld_dsmem_f32x4(la, (uint32_t)c);
        const float4 b = *reinterpret_cast<const float4*>(p.l.bias_g + qg * H + u0);
        float4 s = b;
#pragma unroll
        for (int c = 0; c < kGlSplits; ++c) { s.x += v[c].x; s.y += v[c].y; s.z += v[c].z; s.w += v[c].w; }   // fixed order
        g4[qg][0] = s.x; g4[qg][1] = s.y; g4[qg][2] = s.z; g4[qg][3] = s.w;
      }
      const size_t idx0 = (size_t)row * H + u0;
      const float4 cin = *reinterpret_cast<const float4*>(p.l.c_in + idx0);
      const float cprev[4] = {cin.x, cin.y, cin.z, cin.w};
      float ig[4], fg[4], gg[4], og[4], cn[4], hn[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        ig[e] = sigmoidf_acc(g4[0][e]);
        fg[e] = sigmoidf_acc(g4[1][e]);
        gg[e] = tanhf(g4[2][e]);
        og[e] = sigmoidf_acc(g4[3][e]);
        cn[e] = fg[e] * cprev[e] + ig[e] * gg[e];
        hn[e] = og[e] * tanhf(cn[e]);
      }
      store4<float>(p.l.c_out + idx0, cn);
      if (p.l.acts) {
        float* a = p.l.acts + (size_t)row * 4 * H + u0;
        store4<float>(a, ig); store4<float>(a + H, fg); store4<float>(a + 2 * H, gg); store4<float>(a + 3 * H, og);
      }
      store4<ST>(reinterpret_cast<ST*>(p.l.h_out) + (size_t)row * p.l.h_stride + u0, hn);
      if (p.l.h_f32_out) store4<float>(p.l.h_f32_out + idx0, hn);
      if (p.l.hdrop_out) {
        float hd[4] = {hn[0], hn[1], hn[2], hn[3]};
        if (p.l.mask) {
          const float4 mk = *reinterpret_cast<const float4*>(p.l.mask + idx0);
          hd[0] *= mk.x; hd[1] *= mk.y; hd[2] *= mk.z; hd[3] *= mk.w;
        }
        store4<ST>(reinterpret_cast<ST*>(p.l.hdrop_out) + idx0, hd);
      }
    }
  }

  // nobody leaves (and frees its shared memory) while a neighbour may still be reading it
  __syncwarp();
  cluster_sync_all();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
  }
  trace.end(TK_GEMM_TC + 200);
}

// OFF by default (DIC_FUSED_GATES=1 turns it on).  Measured at batch 256 (in-kernel timeline): this kernel 8.5 us
// per step against 4.0 (split-K GEMM over 144 CTAs) + 1.1 (launch gap) + 3.3 (pointwise kernel) = 8.4 us -- no
// gain.  16 clusters of 64-column tiles did not all become resident (a second wave, 13 us); 8 clusters of
// 128-column tiles are resident at once but use 64 SMs, and a CTA's time is dominated by getting its
// 144 KB of operands through TMA, not by the reduction (vectorising the DSMEM reads changed nothing).
inline bool gates_lstm_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DIC_FUSED_GATES"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1 && tc_enabled();
}
inline bool gates_lstm_eligible(int H, int K, const void* X, long long x_ld) {
  return gates_lstm_enabled() && H % kGlUnits == 0 && K % 8 == 0 && x_ld % 4 == 0 && (K + kTcBK - 1) / kTcBK >= kGlSplits &&
         tc_operand_ok(X, x_ld, 1);
}

// X [rows, K] bf16 (row stride x_ld), Wgp = gate-interleaved [4H, K] bf16 (PackLayout::Wgp)
inline int launch_gates_lstm(const bf16* X, long long x_ld, const bf16* Wgp, int rows, int H, int K,
                             const LstmFwdArgs& l, cudaStream_t st) {
  if (rows <= 0) return 0;
  CUtensorMap tmA, tmB;
  DIC_TRY(make_tmap_bf16(&tmA, X, rows, K, x_ld, kTcBM));
  DIC_TRY(make_tmap_bf16(&tmB, Wgp, 4LL * H, K, K, kGlBN));
  GatesLstmArgs p;
  p.l = l; p.rows = rows; p.H = H; p.K = K; p.tiles_n = 4 * H / kGlBN; p.trace = g_trace_host;
  const int tiles = cdiv(rows, kTcBM) * p.tiles_n;
  static DeviceOnce attr_set;
  if (int dev_ = 0; attr_set.need(&dev_)) {
    DIC_CUDA(cudaFuncSetAttribute(gates_lstm_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)gates_lstm_smem_bytes()));
    attr_set.mark(dev_);
  }
  ProfScope prof(P_LSTM, st);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(kGlSplits, tiles);
  cfg.blockDim = dim3(kGlThreads);
  cfg.dynamicSmemBytes = gates_lstm_smem_bytes();
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kGlSplits; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  DIC_CUDA(cudaLaunchKernelEx(&cfg, gates_lstm_kernel<bf16>, tmA, tmB, p));
  DIC_LAUNCH_CHECK();
  return 0;
}

}  // namespace dic
