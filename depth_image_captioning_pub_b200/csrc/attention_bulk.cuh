// Context pass for several beams per image (beam-search decoding): z_j = sum_l alpha_j[l] F[l,:], j < KB.
//
// With KB beams sharing one image the register-streaming kernel (attention.cuh) needs KB x 8 accumulators
// per thread next to its in-flight loads; at KB = 5 that meant 5 resident CTAs per SM, 1.4 waves and 27 us
// for 103 MB (which sits in L2 at 128 images per GPU), against 15 us with one beam.  Here the annotations
// go global -> shared memory through TMA (cp.async.bulk.tensor.2d boxes of 8 rows x 256 columns, mbarrier
// ring, one elected producer thread; per-row 512-byte bulk copies were tried first and capped at ~4 TB/s
// on the copy engine's request rate), so the bytes in flight live in shared memory instead of registers, and the 128
// consumer threads (4 columns x every other row each) read them back with 8-byte shared loads:
// 13 instructions per row for 4 columns x 5 beams (packed FFMA2).
#pragma once
#include "attention.cuh"
#include "gemm_tc.cuh"     // mbarrier helpers

namespace dic {

constexpr int kBulkRows = 8;        // annotation rows per stage
constexpr int kBulkStages = 6;      // 24 KB of bf16 annotations in flight per CTA, 7 CTAs per SM
constexpr int kBulkCols = 256;      // columns per CTA
constexpr int kBulkThreads = 160;   // 4 consumer warps + 1 producer warp

template <typename ST>
inline size_t attn_ctx_bulk_smem_bytes(int L) {
  // barriers | ring (reused for the final parity reduction: 64 x 8 x 4 floats = 8 KB <= ring) | alpha [L][8]
  return 128 + (size_t)kBulkStages * kBulkRows * kBulkCols * sizeof(ST) + sizeof(float) * 8 * (size_t)L + 128;
}

template <typename ST, int KB>
__global__ void __launch_bounds__(kBulkThreads) attn_context_bulk_kernel(const __grid_constant__ CUtensorMap tmF,
                                                                         const AttnFwdArgs p) {
  static_assert(KB >= 1 && KB <= 8, "KB");
  constexpr int AW = KB <= 4 ? 4 : 8;                    // alpha row width in shared memory (floats)
  constexpr uint32_t ROW_BYTES = kBulkCols * sizeof(ST);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Trace trace(p.trace);
  const int L = p.L, D = p.D, A = p.A;
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  unsigned char* gbase = smem_raw + (sbase - smem_u32(smem_raw));
  auto full_bar = [&](int s) { return sbase + 8u * s; };
  auto empty_bar = [&](int s) { return sbase + 8u * (kBulkStages + s); };
  const uint32_t ring = sbase + 128;
  unsigned char* ring_g = gbase + 128;
  float* al_s = reinterpret_cast<float*>(ring_g + (size_t)kBulkStages * kBulkRows * ROW_BYTES);   // [L][AW]
  float* red_s = reinterpret_cast<float*>(ring_g);          // [64][KB][4], aliases the ring once it is drained

  const int img = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int d0 = blockIdx.x * kBulkCols;
  const int cols = min(kBulkCols, D - d0);
  const int nstages = (L + kBulkRows - 1) / kBulkRows;

  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmF) : "memory");
    for (int s = 0; s < kBulkStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == 4) {
    // ===== producer: the annotations are static, so the ring is filled before the dependency wait =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int it = 0; it < nstages; ++it) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        // one box = 8 annotation rows x 256 columns; rows past this image / columns past D are either the
        // next image's (never read: consumers stop at L) or zero-filled, and always count as full bytes
        mbar_expect_tx(full_bar(stage), kBulkRows * ROW_BYTES);
        tma_load_2d(ring + (uint32_t)stage * kBulkRows * ROW_BYTES, &tmF, full_bar(stage), d0, img * L + it * kBulkRows);
        if (++stage == kBulkStages) { stage = 0; phase ^= 1; }
      }
    }
    pdl_trigger();
    trace.end(TK_CTX);
    return;
  }

  // ===== consumers =====
  pdl_wait();        // alpha and beta come from the preceding kernels of this step
  pdl_trigger();
  trace.mark();
  const int row0 = img * KB;
  for (int i = tid; i < L * AW; i += 128) {
    const int l = i / AW, j = i - l * AW;
    al_s[i] = j < KB ? p.alpha_out[(size_t)(row0 + j) * p.alpha_stride + l] : 0.f;
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");      // consumer warps only

  const int cg = tid & 63, rp = tid >> 6;             // 4 columns each, rows of parity rp within a stage
  const bool active = cg * 4 < cols;
  float acc[KB][4];
#pragma unroll
  for (int j = 0; j < KB; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[j][q] = 0.f;

  int stage = 0;
  uint32_t phase = 0;
  for (int it = 0; it < nstages; ++it) {
    mbar_wait(full_bar(stage), phase);
    const int l0 = it * kBulkRows;
    const ST* srow = reinterpret_cast<const ST*>(ring_g + (size_t)stage * kBulkRows * ROW_BYTES) + cg * 4;
    if (active) {
#pragma unroll
      for (int r = 0; r < kBulkRows / 2; ++r) {
        const int rr = 2 * r + rp;
        const int l = l0 + rr;
        if (l < L) {
          float v[4];
          if constexpr (sizeof(ST) == 2) {
            const uint2 raw = *reinterpret_cast<const uint2*>(srow + (size_t)rr * kBulkCols);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
            const float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
            v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
          } else {
            const float4 raw = *reinterpret_cast<const float4*>(srow + (size_t)rr * kBulkCols);
            v[0] = raw.x; v[1] = raw.y; v[2] = raw.z; v[3] = raw.w;
          }
          float al[AW];
          *reinterpret_cast<float4*>(al) = *reinterpret_cast<const float4*>(al_s + (size_t)l * AW);
          if constexpr (AW == 8) *reinterpret_cast<float4*>(al + 4) = *reinterpret_cast<const float4*>(al_s + (size_t)l * AW + 4);
#pragma unroll
          for (int j = 0; j < KB; ++j) ctx_fmaN<4>(al[j], v, acc[j]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar(stage));
    if (++stage == kBulkStages) { stage = 0; phase ^= 1; }
  }

  // combine the two row parities (fixed order), then the epilogue: z, beta * z
  asm volatile("bar.sync 1, 128;" ::: "memory");      // every consumer warp is done with the ring
  if (rp == 1) {
#pragma unroll
    for (int j = 0; j < KB; ++j)
      *reinterpret_cast<float4*>(red_s + ((size_t)cg * KB + j) * 4) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
  if (rp == 0 && active) {
    const int d = d0 + cg * 4;
#pragma unroll
    for (int j = 0; j < KB; ++j) {
      const float4 o = *reinterpret_cast<const float4*>(red_s + ((size_t)cg * KB + j) * 4);
      acc[j][0] += o.x; acc[j][1] += o.y; acc[j][2] += o.z; acc[j][3] += o.w;
      const int row = row0 + j;
      if (p.z_out) store4<float>(p.z_out + (size_t)row * D + d, acc[j]);
      float beta[4];
      load4<float>(p.hp + (size_t)row * (A + D) + A + d, beta);
      float zg[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) zg[q] = beta[q] * acc[j][q];
      store4<ST>(reinterpret_cast<ST*>(p.zg_out) + (size_t)row * p.zg_stride + d, zg);
    }
  }
  trace.end(TK_CTX);
}

template <typename ST, int KB>
inline int launch_attn_context_bulk(const AttnFwdArgs& p, int images, cudaStream_t st) {
  static DeviceOnce attr_set;
  if (int dev_ = 0; attr_set.need(&dev_)) {
    DIC_CUDA(cudaFuncSetAttribute(attn_context_bulk_kernel<ST, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set.mark(dev_);
  }
  // the annotations as a row-major [images*L, D] matrix; box = 256 columns x 8 rows, no swizzle
  CUtensorMap tmF;
  {
    PFN_tmapEncodeTiled enc = tmap_encoder();
    if (!enc) DIC_FAIL(-6, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t gdim[2] = {(cuuint64_t)p.D, (cuuint64_t)images * p.L};
    cuuint64_t gstr[1] = {(cuuint64_t)p.D * sizeof(ST)};
    cuuint32_t box[2] = {(cuuint32_t)kBulkCols, (cuuint32_t)kBulkRows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmF, sizeof(ST) == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     const_cast<void*>(p.F), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) DIC_FAIL(-6, "cuTensorMapEncodeTiled (annotations) failed with %d", (int)r);
  }
  ProfScope prof(P_ATTN_FWD, st, (double)images * p.L * (double)p.D * sizeof(ST));
  dim3 grid(cdiv(p.D, kBulkCols), images);
  AttnFwdArgs pc = p;
  pc.trace = g_trace_host;
  DIC_CUDA(launch_pdl(attn_context_bulk_kernel<ST, KB>, grid, dim3(kBulkThreads), attn_ctx_bulk_smem_bytes<ST>(p.L), st,
                      tmF, pc));
  DIC_LAUNCH_CHECK();
  return 0;
}

}  // namespace dic
