// Context pass for several beams per image on the tensor cores (bf16 mode, beam-search decoding):
//   z[j, :] = sum_l alpha_j[l] F[l, :],  j < KB <= 8      ==  a [16 x L] x [L x D] product per image.
//
// ncu on the FP32 versions of this pass (attention.cuh / attention_bulk.cuh, 25 us for 128 images x 5 beams,
// profiles/r01_ncu_ctx_bulk_kb5.txt) showed 47 warp instructions per 4 columns x 5 beams of one annotation
// row -- ten FFMA2 plus conversions, address arithmetic, loop control -- i.e. the pass is instruction-issue
// bound at 2.25 IPC while HBM sits at 45 %.  With warp-level MMA (mma.sync.m16n8k16, bf16 x bf16 -> fp32:
// rows 0..KB-1 of the 16-row A fragment are the beams' attention weights, the rest zero) one instruction
// covers 16 annotation rows x 8 columns for all beams, ~30x fewer instructions, and the time no longer depends
// on the beam count.
//
// Feeding the B fragments (profiles/r02_beam_phase_times.txt, 103 MB per launch): a 3-stage TMA ring with a
// producer warp (round 1) 22.3 us; register-fed LDG + byte permutes 29-31 us; the cp.async ring below with
// 4 warps x 3 stages (one wave of CTAs) 21.4 us, with 2 warps x 6 stages (two waves) 20.8 us = 4.95 TB/s.
// The single-beam register-streaming kernel moves the same bytes in 15.6 us; the difference is not the feed
// path (TMA and LSU agree) but the share of a CTA's life spent outside the streaming loop (alpha staging after
// the dependency wait, gate epilogue).
#pragma once
#include "attention_bulk.cuh"

namespace dic {

constexpr int kMmaRows = 16;        // annotation rows per stage = K of one MMA

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

// bf16 copy of the image's KB rows of attention weights in shared memory, [8][LA] (rows >= KB and columns >= L zero).
// All global loads of a thread are issued before the first conversion: the scalar loop this replaces made ~8
// dependent L2 round trips per thread (4 us per CTA, two waves of CTAs per launch).
template <int KB, int NTHR = 128>
__device__ __forceinline__ void stage_alpha16(const AttnFwdArgs& p, bf16* al16, int row0, int LA, int L, int tid) {
  const int chunks = LA / 4;                         // LA is a multiple of 8
  const bool vec = (L & 3) == 0 && (p.alpha_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(p.alpha_out) & 15) == 0;
  constexpr int IT = 512 / NTHR;                     // IT x NTHR threads x 4 columns >= 8 x 216 at the reference shape
  for (int base = 0; base < 8 * chunks; base += IT * NTHR) {
    float4 v[IT];
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      const int i = base + it * NTHR + tid;
      v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < 8 * chunks) {
        const int j = i / chunks, l = (i - j * chunks) * 4;
        if (j < KB && l < L) {
          const float* src = p.alpha_out + (size_t)(row0 + j) * p.alpha_stride + l;
          if (vec) v[it] = *reinterpret_cast<const float4*>(src);
          else {
            v[it].x = src[0];
            if (l + 1 < L) v[it].y = src[1];
            if (l + 2 < L) v[it].z = src[2];
            if (l + 3 < L) v[it].w = src[3];
          }
        }
      }
    }
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      const int i = base + it * NTHR + tid;
      if (i < 8 * chunks) {
        uint2 pk;
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&pk);
        h2[0] = __floats2bfloat162_rn(v[it].x, v[it].y);
        h2[1] = __floats2bfloat162_rn(v[it].z, v[it].w);
        *reinterpret_cast<uint2*>(al16 + (size_t)i * 4) = pk;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// cp.async-fed variant.  Same [16 x L] x [L x D] product per image and the same ldmatrix / MMA consumer loop, but
// the annotations reach shared memory through the LSU path (cp.async.cg, 16 bytes per lane, four full 128-byte row
// segments per warp instruction) instead of TMA boxes: each warp owns the 64 columns it consumes, keeps its own
// ring of kCpaStages x (16 rows x 128 bytes) and never meets a block-wide barrier.  The swizzle of the TMA box
// (16-byte chunk c of row r at chunk c ^ (r & 7)) is reproduced by the destination addresses, so the B-fragment
// loads stay conflict free.
// Geometry: WARPS x 64 columns per CTA, STAGES ring slots per warp.  (4, 3) = 24 KB + alpha per CTA, 7 CTAs per SM made
// the 1024 CTAs of 128 images ONE wave, every CTA in the same phase (alpha staging, streaming, gate epilogue) at the
// same time.  (2, 6) = the same shared memory and twice the CTAs: two waves, so the second wave's prologue overlaps
// the first wave's streaming, and 10 KB instead of 4 KB in flight per warp.
template <int WARPS, int STAGES>
inline size_t attn_ctx_mma_cpa_smem_bytes(int L) {
  const int nk = (L + kMmaRows - 1) / kMmaRows;
  return 128 + (size_t)STAGES * kMmaRows * (WARPS * 64) * 2 + (size_t)8 * (nk * kMmaRows + 8) * 2 + 64;
}

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

template <int KB, int WARPS, int STAGES>
__global__ void __launch_bounds__(WARPS * 32) attn_context_mma_cpa_kernel(const AttnFwdArgs p) {
  static_assert(KB >= 1 && KB <= 8, "KB");
  constexpr int NTHR = WARPS * 32, COLS = WARPS * 64;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Trace trace(p.trace);
  const int L = p.L, D = p.D, A = p.A;
  const int nk = (L + kMmaRows - 1) / kMmaRows;
  const int LA = nk * kMmaRows + 8;
  const uint32_t ring = (smem_u32(smem_raw) + 127u) & ~127u;          // the swizzle here is by row index, not by address
  unsigned char* ring_g = smem_raw + (ring - smem_u32(smem_raw));
  constexpr uint32_t WBOX = kMmaRows * 128;                          // one warp's stage: 16 rows x 128 bytes
  constexpr uint32_t STAGE_BYTES = WARPS * WBOX;
  bf16* al16 = reinterpret_cast<bf16*>(ring_g + STAGES * STAGE_BYTES);          // [8][LA]

  const int img = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int dw = blockIdx.x * COLS + warp * 64;                        // first column of this warp
  const bf16* Fimg = reinterpret_cast<const bf16*>(p.F) + (size_t)img * L * D;
  const int c8 = lane & 7, r4 = lane >> 3;                             // this lane's chunk and row (mod 4) of a stage
  const bool col_ok = dw + c8 * 8 + 7 < D;
  // Address generation is hoisted out of the stream: the four destination offsets of a lane inside a warp's stage
  // are constants, the four source pointers advance by 16 rows per step.  (Recomputing them per step was 55 of the
  // 85 instructions of the main loop, and ncu put the kernel at 37 % issue activity x 7.4 M warp instructions =
  // 17 of its 21 us.)  Only the last step of an image can run past its L rows: it alone carries the row checks.
  uint32_t doff[4];
  const char* sp[4];
  const size_t step_bytes = (size_t)kMmaRows * D * 2;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int r = q * 4 + r4;
    doff[q] = (uint32_t)(r * 128 + ((c8 ^ (r & 7)) * 16));
    sp[q] = reinterpret_cast<const char*>(Fimg + (col_ok ? dw + c8 * 8 : 0) + (size_t)r * D);
  }
  const uint32_t wring = ring + warp * WBOX;
  const uint32_t ok_bytes = col_ok ? 16u : 0u;
  // issue step `ks` into ring slot `stage`; sp[] must point at step ks (the caller advances it)
  auto issue = [&](int ks, int stage) {
    const uint32_t box = wring + stage * STAGE_BYTES;
    if (ks + 1 < nk) {
#pragma unroll
      for (int q = 0; q < 4; ++q) cp_async16_zfill(box + doff[q], sp[q], ok_bytes);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const bool ok = ks * kMmaRows + q * 4 + r4 < L;
        cp_async16_zfill(box + doff[q], ok ? sp[q] : reinterpret_cast<const char*>(Fimg), ok ? ok_bytes : 0u);
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) sp[q] += step_bytes;
  };
  // the annotations are static: the ring is filled before the dependency wait
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) issue(s, s);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  if (!p.late_wait) pdl_wait();        // alpha and beta come from the preceding kernels of this step
  pdl_trigger();
  trace.mark();
  const int row0 = img * KB;
  // the gate of this thread's output columns: requested now, used in the epilogue (the round trip hides behind the pass)
  const int j = lane >> 2;
  float2 beta[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int d = dw + t * 8 + (lane & 3) * 2;
    beta[t] = (j < KB && d < D) ? __ldcg(reinterpret_cast<const float2*>(p.hp + (size_t)(row0 + j) * (A + D) + A + d))
                                : make_float2(0.f, 0.f);
  }
  stage_alpha16<KB, NTHR>(p, al16, row0, LA, L, tid);
  __syncthreads();

  float acc[8][4];
#pragma unroll
  for (int t = 0; t < 8; ++t)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[t][q] = 0.f;
  const uint32_t* a_row = reinterpret_cast<const uint32_t*>(al16 + (size_t)(lane >> 2) * LA + (lane & 3) * 2);
  const int lrow = ((lane >> 3) & 1) * 8 + (lane & 7);
  const int lchunk = lane >> 4;
  int stage = 0;
  for (int ks = 0; ks < nk; ++ks) {
    const uint32_t a0 = a_row[ks * 8];
    const uint32_t a2 = a_row[ks * 8 + 4];
    asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 2) : "memory");           // step ks has landed (this lane's part)
    __syncwarp();                                                                    // ... and every other lane's
    // refill the slot consumed in the previous iteration (all lanes are past its ldmatrix reads: __syncwarp above)
    {
      const int kn = ks + STAGES - 1;
      if (kn < nk) issue(kn, stage == 0 ? STAGES - 1 : stage - 1);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    const uint32_t box = ring + stage * STAGE_BYTES + warp * WBOX + lrow * 128;
#pragma unroll
    for (int pr = 0; pr < 4; ++pr) {
      uint32_t b0, b1, b2, b3;
      ldmatrix_x4_trans(box + (uint32_t)(((pr * 2 + lchunk) ^ (lrow & 7)) * 16), b0, b1, b2, b3);
      mma_bf16_16816(acc[pr * 2], a0, 0u, a2, 0u, b0, b1);
      mma_bf16_16816(acc[pr * 2 + 1], a0, 0u, a2, 0u, b2, b3);
    }
    if (++stage == STAGES) stage = 0;
  }

  if (j < KB) {
    const int row = row0 + j;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int d = dw + t * 8 + (lane & 3) * 2;
      if (d < D) {
        if (p.z_out) *reinterpret_cast<float2*>(p.z_out + (size_t)row * D + d) = make_float2(acc[t][0], acc[t][1]);
        *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<bf16*>(p.zg_out) + (size_t)row * p.zg_stride + d) =
            __floats2bfloat162_rn(beta[t].x * acc[t][0], beta[t].y * acc[t][1]);
      }
    }
  }
  trace.end(TK_CTX);
  // late_wait: ONE CTA sits out the predecessor -- enough for "this grid complete => predecessor complete".  (With
  // every CTA waiting here the first wave kept its slots until the predecessor had finished and the second wave
  // started 16 us late: profiles/r02_beam_lookahead.txt.)
  if (p.late_wait && blockIdx.x == 0 && blockIdx.y == 0) pdl_wait();
}

template <int KB, int WARPS, int STAGES>
inline int launch_attn_context_mma_cpa(const AttnFwdArgs& p, int images, cudaStream_t st) {
  static DeviceOnce attr_set;
  if (int dev_ = 0; attr_set.need(&dev_)) {
    DIC_CUDA(cudaFuncSetAttribute(attn_context_mma_cpa_kernel<KB, WARPS, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  100 * 1024));
    attr_set.mark(dev_);
  }
  ProfScope prof(P_ATTN_FWD, st, (double)images * p.L * (double)p.D * 2);
  dim3 grid(cdiv(p.D, WARPS * 64), images);
  AttnFwdArgs pc = p;
  pc.trace = g_trace_host;
  DIC_CUDA(launch_pdl(attn_context_mma_cpa_kernel<KB, WARPS, STAGES>, grid, dim3(WARPS * 32),
                      attn_ctx_mma_cpa_smem_bytes<WARPS, STAGES>(p.L), st, pc));
  DIC_LAUNCH_CHECK();
  return 0;
}

// dispatch shim: the tensor-core kernel exists for bf16 storage only (the caller checks D % 8 == 0)
template <typename ST, int KB>
inline int launch_attn_context_mma_st(const AttnFwdArgs& p, int images, cudaStream_t st) {
  if constexpr (sizeof(ST) == 2) return launch_attn_context_mma_cpa<KB, 2, 6>(p, images, st);
  else return launch_attn_context_bulk<ST, KB>(p, images, st);
}

}  // namespace dic
