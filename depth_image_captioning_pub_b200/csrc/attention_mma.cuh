// Context pass for several beams per image on the tensor cores (bf16 mode, beam-search decoding):
//   z[j, :] = sum_l alpha_j[l] F[l, :],  j < KB <= 8      ==  a [16 x L] x [L x D] product per image.
//
// ncu on the FP32 versions of this pass (attention.cuh / attention_bulk.cuh, 25 us for 128 images x 5 beams,
// profiles/r01_ncu_ctx_bulk_kb5.txt) showed 47 warp instructions per 4 columns x 5 beams of one annotation
// row -- ten FFMA2 plus conversions, address arithmetic, loop control -- i.e. the pass is instruction-issue
// bound at 2.25 IPC while HBM sits at 45 %.  With warp-level MMA (mma.sync.m16n8k16, bf16 x bf16 -> fp32:
// rows 0..KB-1 of the 16-row A fragment are the beams' attention weights, the rest zero) one instruction
// covers 16 annotation rows x 8 columns for all beams, ~30x fewer instructions, and the time no longer depends
// on the beam count (24 us for 3, 5 or 8 beams).  It does NOT reach the 15 us of the single-beam register-
// streaming kernel on the same annotations: every TMA-fed variant of this pass (this one, the FP32 staged one,
// with 128-byte or 512-byte wide boxes, any L2 promotion) levels off near 4.5 TB/s at 7 resident CTAs per SM,
// while plain 16-byte LDG streaming reaches 7 TB/s (L2 hits included).  Open question for the next round.
//
// CTA = 256 columns of one image: 4 consumer warps (64 columns = one 128-byte-swizzled TMA box each) and one
// producer warp; ring of 3 stages x 16 rows; B fragments come out of the swizzled stage with
// ldmatrix.x4.trans, A fragments from a bf16 copy of alpha in shared memory.
#pragma once
#include "attention_bulk.cuh"

namespace dic {

constexpr int kMmaRows = 16;        // annotation rows per stage = K of one MMA
constexpr int kMmaStages = 3;
constexpr int kMmaCols = 256;
constexpr int kMmaThreads = 160;

inline size_t attn_ctx_mma_smem_bytes(int L) {
  const int nk = (L + kMmaRows - 1) / kMmaRows;
  // 1024 (alignment slack) | barriers 128 | ring | alpha bf16 [8][nk*16 + 8]
  return 1024 + 128 + (size_t)kMmaStages * kMmaRows * kMmaCols * 2 + (size_t)8 * (nk * kMmaRows + 8) * 2 + 64;
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int KB>
__global__ void __launch_bounds__(kMmaThreads) attn_context_mma_kernel(const __grid_constant__ CUtensorMap tmF,
                                                                       const AttnFwdArgs p) {
  static_assert(KB >= 1 && KB <= 8, "KB");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Trace trace(p.trace);
  const int L = p.L, D = p.D, A = p.A;
  const int nk = (L + kMmaRows - 1) / kMmaRows;
  const int LA = nk * kMmaRows + 8;                     // alpha row stride (bf16 elements); +8 staggers the banks
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;       // swizzle-128B boxes want 1024-byte alignment
  unsigned char* ring_g = smem_raw + (ring - smem_u32(smem_raw));
  constexpr uint32_t STAGE_BYTES = kMmaRows * kMmaCols * 2;           // 8 KB = 4 boxes of 16 rows x 128 bytes
  const uint32_t bar_base = ring + kMmaStages * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMmaStages + s); };
  bf16* al16 = reinterpret_cast<bf16*>(ring_g + kMmaStages * STAGE_BYTES + 128);      // [8][LA]

  const int img = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int d0 = blockIdx.x * kMmaCols;

  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmF) : "memory");
    for (int s = 0; s < kMmaStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == 4) {
    // ===== producer: the annotations are static, the ring is filled before the dependency wait =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ks = 0; ks < nk; ++ks) {
        mbar_wait(empty_bar(stage), phase ^ 1);
        mbar_expect_tx(full_bar(stage), STAGE_BYTES);
        // rows past this image / columns past D: the next image's rows (multiplied by alpha = 0) or zero fill
#pragma unroll
        for (int w = 0; w < 4; ++w)
          tma_load_2d(ring + stage * STAGE_BYTES + w * (kMmaRows * 128), &tmF, full_bar(stage), d0 + 64 * w,
                      img * L + ks * kMmaRows);
        if (++stage == kMmaStages) { stage = 0; phase ^= 1; }
      }
    }
    pdl_trigger();
    return;
  }

  // ===== consumers =====
  pdl_wait();        // alpha and beta come from the preceding kernels of this step
  pdl_trigger();
  trace.mark();
  const int row0 = img * KB;
  for (int i = tid; i < 8 * LA; i += 128) {
    const int j = i / LA, l = i - j * LA;
    float a = 0.f;
    if (j < KB && l < L) a = p.alpha_out[(size_t)(row0 + j) * p.alpha_stride + l];
    al16[i] = __float2bfloat16_rn(a);
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");      // consumer warps only

  float acc[8][4];
#pragma unroll
  for (int t = 0; t < 8; ++t)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[t][q] = 0.f;

  // A fragment addressing: row (beam) = lane / 4, k = (lane % 4) * 2 (+8 for the second half)
  const uint32_t* a_row = reinterpret_cast<const uint32_t*>(al16 + (size_t)(lane >> 2) * LA + (lane & 3) * 2);
  // ldmatrix.x4.trans lane addressing inside a warp's 16-row x 128-byte box (128B swizzle):
  //   matrix = lane / 8: rows (matrix & 1) * 8 + lane % 8, 16-byte chunk (n0 / 8) + (matrix >> 1)
  const int lrow = ((lane >> 3) & 1) * 8 + (lane & 7);
  const int lchunk = lane >> 4;
  int stage = 0;
  uint32_t phase = 0;
  for (int ks = 0; ks < nk; ++ks) {
    const uint32_t a0 = a_row[ks * 8];           // bf16 pairs: (k, k+1)
    const uint32_t a2 = a_row[ks * 8 + 4];       // (k+8, k+9)
    mbar_wait(full_bar(stage), phase);
    const uint32_t box = ring + stage * STAGE_BYTES + warp * (kMmaRows * 128) + lrow * 128;
#pragma unroll
    for (int pr = 0; pr < 4; ++pr) {             // pairs of 8-column tiles: columns pr*16 .. pr*16+15 of the warp's 64
      uint32_t b0, b1, b2, b3;
      ldmatrix_x4_trans(box + (uint32_t)(((pr * 2 + lchunk) ^ (lrow & 7)) * 16), b0, b1, b2, b3);
      mma_bf16_16816(acc[pr * 2], a0, 0u, a2, 0u, b0, b1);
      mma_bf16_16816(acc[pr * 2 + 1], a0, 0u, a2, 0u, b2, b3);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar(stage));
    if (++stage == kMmaStages) { stage = 0; phase ^= 1; }
  }

  // epilogue: c0, c1 of every tile are row (beam) lane / 4, columns (lane % 4) * 2 + {0, 1}
  const int j = lane >> 2;
  if (j < KB) {
    const int row = row0 + j;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int d = d0 + warp * 64 + t * 8 + (lane & 3) * 2;
      if (d < D) {
        if (p.z_out) *reinterpret_cast<float2*>(p.z_out + (size_t)row * D + d) = make_float2(acc[t][0], acc[t][1]);
        const float2 beta = *reinterpret_cast<const float2*>(p.hp + (size_t)row * (A + D) + A + d);
        *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<bf16*>(p.zg_out) + (size_t)row * p.zg_stride + d) =
            __floats2bfloat162_rn(beta.x * acc[t][0], beta.y * acc[t][1]);
      }
    }
  }
  trace.end(TK_CTX);
}

template <int KB>
inline int launch_attn_context_mma(const AttnFwdArgs& p, int images, cudaStream_t st) {
  static DeviceOnce attr_set;
  if (int dev_ = 0; attr_set.need(&dev_)) {
    DIC_CUDA(cudaFuncSetAttribute(attn_context_mma_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_set.mark(dev_);
  }
  CUtensorMap tmF;     // annotations as a row-major [images*L, D] bf16 matrix; box = 64 columns x 16 rows, 128B swizzle
  DIC_TRY(make_tmap_bf16(&tmF, p.F, (long long)images * p.L, p.D, p.D, kMmaRows));   // (L2 promotion: no effect measured)
  ProfScope prof(P_ATTN_FWD, st, (double)images * p.L * (double)p.D * 2);
  dim3 grid(cdiv(p.D, kMmaCols), images);
  AttnFwdArgs pc = p;
  pc.trace = g_trace_host;
  DIC_CUDA(launch_pdl(attn_context_mma_kernel<KB>, grid, dim3(kMmaThreads), attn_ctx_mma_smem_bytes(p.L), st, tmF, pc));
  DIC_LAUNCH_CHECK();
  return 0;
}

// dispatch shim: the tensor-core kernel exists for bf16 storage only
template <typename ST, int KB>
inline int launch_attn_context_mma_st(const AttnFwdArgs& p, int images, cudaStream_t st) {
  if constexpr (sizeof(ST) == 2) return launch_attn_context_mma<KB>(p, images, st);
  else return launch_attn_context_bulk<ST, KB>(p, images, st);
}

}  // namespace dic
