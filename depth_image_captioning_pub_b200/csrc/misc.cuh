// Memory-bound helpers: RGB+depth fusion, embedding gather/scatter, column sums, weight packing.
#pragma once
#include "common.cuh"

namespace dic {

// ---- K0a: Fsum = F_rgb (+ F_depth), mean over L ------------------------------------------------
// features.add(depth_features) and features.mean(dim=1) (depth_models.py:163,166) in one pass.
// Grid (D chunks of 1024 columns, B); each thread owns 4 columns and walks the L rows.
template <typename TIN, typename ST>
__global__ void __launch_bounds__(256) fuse_feats_kernel(const TIN* __restrict__ rgb,
                                                         const TIN* __restrict__ dep, ST* __restrict__ fsum,
                                                         float* __restrict__ meanF, bf16* __restrict__ mean16,
                                                         int L, int D) {
  pdl_wait();          // launched programmatically (launch_pdl): nothing of a predecessor is touched before this
  pdl_trigger();
  const int b = blockIdx.y;
  const int d = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (d >= D) return;
  const size_t base = (size_t)b * L * D + d;
  float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int l = 0; l < L; ++l) {
    float v[4];
    const size_t o = base + (size_t)l * D;
    if (sizeof(TIN) == 4) {
      float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(rgb) + o);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
      if (dep) {
        float4 c = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(dep) + o);
        v[0] += c.x; v[1] += c.y; v[2] += c.z; v[3] += c.w;
      }
    } else {
      uint2 a = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(rgb) + o);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&a);
      float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
      v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
      if (dep) {
        uint2 c = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(dep) + o);
        const __nv_bfloat162* g = reinterpret_cast<const __nv_bfloat162*>(&c);
        float2 g0 = __bfloat1622float2(g[0]), g1 = __bfloat1622float2(g[1]);
        v[0] += g0.x; v[1] += g0.y; v[2] += g1.x; v[3] += g1.y;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) s[q] += v[q];
    if (fsum) {
      if (sizeof(ST) == 4) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(fsum) + o) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
        uint2 r;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
        h[0] = __floats2bfloat162_rn(v[0], v[1]);
        h[1] = __floats2bfloat162_rn(v[2], v[3]);
        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(fsum) + o) = r;
      }
    }
  }
  const float inv = 1.f / (float)L;
  *reinterpret_cast<float4*>(meanF + (size_t)b * D + d) =
      make_float4(s[0] * inv, s[1] * inv, s[2] * inv, s[3] * inv);
  if (mean16) {     // bf16 copy: A operand of the init_linear tensor-core GEMMs
    uint2 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
    h[0] = __floats2bfloat162_rn(s[0] * inv, s[1] * inv);
    h[1] = __floats2bfloat162_rn(s[2] * inv, s[3] * inv);
    *reinterpret_cast<uint2*>(mean16 + (size_t)b * D + d) = r;
  }
}

// fsum == nullptr -> only the mean is produced (the caller aliases Fsum to the input).
template <typename ST>
inline int launch_fuse_feats(const void* rgb, const void* dep, int feat_bf16, ST* fsum, float* meanF,
                             bf16* mean16, int B, int L, int D, cudaStream_t st) {
  dim3 grid(cdiv(D, 1024), B);
  ProfScope prof(P_FUSE, st);
  if (feat_bf16)
    DIC_CUDA(launch_pdl(fuse_feats_kernel<bf16, ST>, grid, dim3(256), 0, st, reinterpret_cast<const bf16*>(rgb),
                        reinterpret_cast<const bf16*>(dep), fsum, meanF, mean16, L, D));
  else
    DIC_CUDA(launch_pdl(fuse_feats_kernel<float, ST>, grid, dim3(256), 0, st, reinterpret_cast<const float*>(rgb),
                        reinterpret_cast<const float*>(dep), fsum, meanF, mean16, L, D));
  DIC_LAUNCH_CHECK();
  return 0;
}

// ---- column sums: dst[c] = sum_r src[r*ld + c] (bias gradients) ---------------------------------
// Grid (column tiles of 32, row chunks); partial sums meet in dst through atomicAdd, so dst
// must be zeroed first (launcher does it).
__global__ void __launch_bounds__(256) colsum_kernel(const void* __restrict__ src, int src_bf16, int R,
                                                     int C, long long ld, int rows_per_block,
                                                     float* __restrict__ dst) {
  pdl_wait();          // launched programmatically (launch_pdl): nothing of a predecessor is touched before this
  pdl_trigger();
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(R, r0 + rows_per_block);
  float s = 0.f;
  if (c < C) {
    int r = r0 + ry;
    for (; r + 56 < r1; r += 64) {      // 8 independent loads in flight per thread
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = ld_as_float(src, (size_t)(r + 8 * i) * ld + c, src_bf16);
#pragma unroll
      for (int i = 0; i < 8; ++i) s += v[i];
    }
    for (; r < r1; r += 8) s += ld_as_float(src, (size_t)r * ld + c, src_bf16);
  }
  red[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][cx];
    atomicAdd(dst + c, t);
  }
}

// wide bf16 matrices (d_logits [N,V], G [T*B, 4H+A+D]): a thread owns 8 consecutive columns (16-byte
// loads, a warp covers 512 contiguous bytes of a row), 8 row groups per CTA, 4 rows in flight per thread
__global__ void __launch_bounds__(256) colsum_bf16x8_kernel(const bf16* __restrict__ src, int R, int C, long long ld,
                                                            int rows_per_block, float* __restrict__ dst) {
  pdl_wait();          // launched programmatically (launch_pdl): nothing of a predecessor is touched before this
  pdl_trigger();
  __shared__ float red[8][32][9];
  const int lane = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 256 + lane * 8;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(R, r0 + rows_per_block);
  float s[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) s[q] = 0.f;
  if (c < C) {
    int r = r0 + ry;
    for (; r + 24 < r1; r += 32) {
      Raw8<bf16> raw[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) raw[i].load_stream(src + (size_t)(r + 8 * i) * ld + c);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v[8];
        raw[i].unpack(v);
#pragma unroll
        for (int q = 0; q < 8; ++q) s[q] += v[q];
      }
    }
    for (; r < r1; r += 8) {
      float v[8];
      load8_stream<bf16>(src + (size_t)r * ld + c, v);
#pragma unroll
      for (int q = 0; q < 8; ++q) s[q] += v[q];
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) red[ry][lane][q] = s[q];
  __syncthreads();
  // 256 threads = 256 columns of the CTA
  const int cl = threadIdx.x, ln = cl >> 3, q = cl & 7;
  const int cc = blockIdx.x * 256 + cl;
  if (cc < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][ln][q];
    atomicAdd(dst + cc, t);
  }
}

// zeroed: the caller has cleared dst already (the backward clears all its accumulation targets in one launch)
inline int launch_colsum(const void* src, int src_bf16, int R, int C, long long ld, float* dst,
                         cudaStream_t st, bool zeroed = false) {
  ProfScope prof(P_COLSUM, st);
  if (!zeroed) DIC_CUDA(cudaMemsetAsync(dst, 0, sizeof(float) * C, st));
  if (R <= 0 || C <= 0) return 0;
  if (src_bf16 && C >= 256 && C % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int ctiles = cdiv(C, 256);
    int chunks = cdiv(148 * 8, ctiles);                 // ~8 CTAs per SM
    if (chunks > cdiv(R, 32)) chunks = cdiv(R, 32);
    if (chunks < 1) chunks = 1;
    const int rpb = cdiv(cdiv(R, chunks), 8) * 8;
    dim3 grid(ctiles, cdiv(R, rpb));
    DIC_CUDA(launch_pdl(colsum_bf16x8_kernel, grid, dim3(256), 0, st, reinterpret_cast<const bf16*>(src), R, C, ld, rpb, dst));
    DIC_LAUNCH_CHECK();
    return 0;
  }
  int chunks = cdiv(R, 64);              // one trip of 8 loads per thread: these are latency-bound launches of a few MB
  if (chunks > 1024) chunks = 1024;
  const int rpb = cdiv(R, chunks);
  dim3 grid(cdiv(C, 32), cdiv(R, rpb));
  DIC_CUDA(launch_pdl(colsum_kernel, grid, dim3(256), 0, st, src, src_bf16, R, C, ld, rpb, dst));
  DIC_LAUNCH_CHECK();
  return 0;
}

// ---- embedding gather for teacher forcing: X[t][b][0:E] = Emb[captions[b,t]] (depth_models.py:160)
template <typename ST>
__global__ void __launch_bounds__(256) embed_gather_tf_kernel(const ST* __restrict__ emb,
                                                              const int64_t* __restrict__ captions,
                                                              int cap_stride, ST* __restrict__ X,
                                                              long long x_row, long long x_step, int B,
                                                              int E, int V, StepSizes sizes, int T) {
  pdl_wait();          // launched programmatically (launch_pdl): nothing of a predecessor is touched before this
  pdl_trigger();
  const int t = blockIdx.y;
  const int n = sizes.n[t];
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n * E; i += gridDim.x * 256) {
    const int b = i / E, e = i - b * E;
    long long tok = captions[(size_t)b * cap_stride + t];
    if (tok < 0 || tok >= V) tok = 0;  // out-of-range ids are a caller bug; stay in bounds
    X[(size_t)t * x_step + (size_t)b * x_row + e] = emb[(size_t)tok * E + e];
  }
}

// ---- embedding gradient: dEmb[captions[b,t]] += dX[t][b]  (dEmb zeroed by the caller) -------------
__global__ void __launch_bounds__(256) embed_scatter_add_kernel(const float* __restrict__ dX,
                                                                const int64_t* __restrict__ captions,
                                                                int cap_stride, float* __restrict__ dEmb,
                                                                int B, int E, int V, StepSizes sizes,
                                                                int T) {
  pdl_wait();          // launched programmatically (launch_pdl): nothing of a predecessor is touched before this
  pdl_trigger();
  const int t = blockIdx.y;
  const int n = sizes.n[t];
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n * E; i += gridDim.x * 256) {
    const int b = i / E, e = i - b * E;
    const long long tok = captions[(size_t)b * cap_stride + t];
    if (tok < 0 || tok >= V) continue;
    atomicAdd(dEmb + (size_t)tok * E + e, dX[((size_t)t * B + b) * E + e]);
  }
}

// ---- 2-D strided copy with dtype conversion (weight packing) ------------------------------------
__global__ void __launch_bounds__(256) copy2d_kernel(const float* __restrict__ src, long long src_ld,
                                                     void* __restrict__ dst, long long dst_ld,
                                                     int dst_bf16, int R, int C) {
  const size_t n = (size_t)R * C;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const size_t r = i / C, c = i - r * C;
    st_from_float(dst, r * dst_ld + c, src[r * src_ld + c], dst_bf16);
  }
}

// contiguous fp32 -> bf16 cast, 8 elements per thread (16-byte stores)
__global__ void __launch_bounds__(256) cast_bf16_vec_kernel(const float* __restrict__ src, bf16* __restrict__ dst,
                                                            size_t n8) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n8; i += (size_t)gridDim.x * 256) {
    float v[8];
    load8_stream<float>(src + i * 8, v);
    store8<bf16>(dst + i * 8, v);
  }
}

inline int launch_copy2d(const float* src, long long src_ld, void* dst, long long dst_ld, int dst_bf16,
                         int R, int C, cudaStream_t st) {
  const size_t n = (size_t)R * C;
  if (n == 0) return 0;
  if (dst_bf16 && src_ld == C && dst_ld == C && n % 8 == 0 && n >= (1u << 16) &&
      ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    size_t n8 = n / 8;
    int blocks = (int)((n8 + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    cast_bf16_vec_kernel<<<blocks, 256, 0, st>>>(src, reinterpret_cast<bf16*>(dst), n8);
    DIC_LAUNCH_CHECK();
    return 0;
  }
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  copy2d_kernel<<<blocks, 256, 0, st>>>(src, src_ld, dst, dst_ld, dst_bf16, R, C);
  DIC_LAUNCH_CHECK();
  return 0;
}

// ---- weight packing in ONE launch: a table of 2-D copy/convert jobs (blockIdx.y = job) ----------------
// (17 separate copy kernels cost ~55 us of launch latency per training step; the data is 10 MB)
struct PackJob {
  const float* src;
  const float* src2;       // optional second addend (b_ih + b_hh)
  void* dst;
  long long src_ld, dst_ld;
  int R, C, dst_bf16;
  int gate_H;              // > 0: destination row of source row q*gate_H + u is (u/U)*4U + q*U + u%U, U = gate_U
  int gate_U;
};
constexpr int kMaxPackJobs = 24;
struct PackJobs {
  PackJob j[kMaxPackJobs];
  int n;
};
__global__ void __launch_bounds__(256) pack_jobs_kernel(const PackJobs jobs) {
  pdl_wait();          // launched programmatically (launch_pdl): nothing of a predecessor is touched before this
  pdl_trigger();
  const PackJob& jb = jobs.j[blockIdx.y];
  const size_t n = (size_t)jb.R * jb.C;
  const bool dense = jb.src_ld == jb.C && jb.dst_ld == jb.C;
  if (dense && jb.dst_bf16 && !jb.src2 && jb.gate_H == 0 && n % 8 == 0 &&
      ((reinterpret_cast<uintptr_t>(jb.src) | reinterpret_cast<uintptr_t>(jb.dst)) & 15) == 0) {
    const size_t n8 = n / 8;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n8; i += (size_t)gridDim.x * 256) {
      float v[8];
      load8_stream<float>(jb.src + i * 8, v);
      store8<bf16>(reinterpret_cast<bf16*>(jb.dst) + i * 8, v);
    }
    return;
  }
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    const size_t r = i / jb.C, c = i - r * jb.C;
    float v = jb.src[r * jb.src_ld + c];
    if (jb.src2) v += jb.src2[r * jb.src_ld + c];
    size_t rd = r;
    if (jb.gate_H > 0) {
      const size_t q = r / jb.gate_H, u = r - q * jb.gate_H;
      rd = (u / jb.gate_U) * 4 * jb.gate_U + q * jb.gate_U + (u % jb.gate_U);
    }
    st_from_float(jb.dst, rd * jb.dst_ld + c, v, jb.dst_bf16);
  }
}

// ---- fused multi-tensor AdamW (the reference steps torch.optim.AdamW over the decoder
// and depth-encoder parameters, depth_train.py:136-137,221).  One launch for up to kMaxOptTensors
// tensors (blockIdx.y = tensor); decoupled weight decay, bias-corrected moments, no amsgrad:
//   p *= 1 - lr*wd ;  m += (g - m)(1 - b1) ;  v = b2 v + (1 - b2) g^2 ;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
constexpr int kMaxOptTensors = 32;
struct AdamWJobs {
  float* p[kMaxOptTensors];
  const float* g[kMaxOptTensors];
  float* m[kMaxOptTensors];
  float* v[kMaxOptTensors];
  long long n[kMaxOptTensors];
  int count;
  float lr_wd;        // lr * weight_decay
  float one_m_b1, b2, one_m_b2;
  float step_size;    // lr / (1 - b1^t)
  float inv_sqrt_bc2; // 1 / sqrt(1 - b2^t)
  float eps;
};
__global__ void __launch_bounds__(256) adamw_kernel(const AdamWJobs a) {
  pdl_wait();          // launched programmatically (launch_pdl): nothing of a predecessor is touched before this
  pdl_trigger();
  const int j = blockIdx.y;
  float* __restrict__ p = a.p[j];
  const float* __restrict__ g = a.g[j];
  float* __restrict__ m = a.m[j];
  float* __restrict__ v = a.v[j];
  const long long n = a.n[j];
  const long long stride = (long long)gridDim.x * 256;
  const bool vec = (n % 4 == 0) && (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                                      reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0);
  auto upd = [&](float& pw, float gw, float& mw, float& vw) {
    pw *= 1.f - a.lr_wd;
    mw += (gw - mw) * a.one_m_b1;
    vw = vw * a.b2 + a.one_m_b2 * gw * gw;
    const float denom = sqrtf(vw) * a.inv_sqrt_bc2 + a.eps;
    pw -= a.step_size * (mw / denom);
  };
  if (vec) {
    const long long n4 = n / 4;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += stride) {
      float4 pw = reinterpret_cast<float4*>(p)[i];
      const float4 gw = reinterpret_cast<const float4*>(g)[i];
      float4 mw = reinterpret_cast<float4*>(m)[i];
      float4 vw = reinterpret_cast<float4*>(v)[i];
      upd(pw.x, gw.x, mw.x, vw.x); upd(pw.y, gw.y, mw.y, vw.y);
      upd(pw.z, gw.z, mw.z, vw.z); upd(pw.w, gw.w, mw.w, vw.w);
      reinterpret_cast<float4*>(p)[i] = pw;
      reinterpret_cast<float4*>(m)[i] = mw;
      reinterpret_cast<float4*>(v)[i] = vw;
    }
  } else {
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += stride) upd(p[i], g[i], m[i], v[i]);
  }
}

__global__ void __launch_bounds__(256) add_vec_kernel(const float* a, const float* b, float* dst, int n) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) dst[i] = a[i] + b[i];
}

// ---- backward of init_linear, operand prep: dhc16[b, 0:H] = dh0, [H:2H] = dc0 (bf16), and the bias
// gradient (column sums over the batch) in the same pass.  Grid 2H/32 CTAs of 256 threads (32 columns x 8
// row groups).
__global__ void __launch_bounds__(256) dhc_prep_kernel(const float* __restrict__ dh, const float* __restrict__ dc,
                                                       bf16* __restrict__ dhc16, float* __restrict__ db, int B, int H) {
  pdl_wait();          // launched programmatically (launch_pdl): nothing of a predecessor is touched before this
  pdl_trigger();
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;          // column in [0, 2H)
  float s = 0.f;
  if (c < 2 * H) {
    const float* src = c < H ? dh + c : dc + (c - H);
    int r = ry;
    for (; r + 56 < B; r += 64) {        // 8 independent loads in flight per thread (the one-load-per-trip loop took 11 us)
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = src[(size_t)(r + 8 * i) * H];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        s += v[i];
        dhc16[(size_t)(r + 8 * i) * 2 * H + c] = __float2bfloat16_rn(v[i]);
      }
    }
    for (; r < B; r += 8) {
      const float v = src[(size_t)r * H];
      s += v;
      dhc16[(size_t)r * 2 * H + c] = __float2bfloat16_rn(v);
    }
  }
  red[ry][cx] = s;
  __syncthreads();
  if (ry == 0 && c < 2 * H) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][cx];
    db[c] = t;
  }
}

// ---- dL/dF accumulation -------------------------------------------------------------------------
//   dF[b,l,d] (+)= sum_t alpha_t[b,l] * dz_t[b,d] + dmeanF[b,d] / L
// the context path (attention.py:93) and the mean path (depth_models.py:166) of the annotation
// gradient; the encoder_att path (datt1 . W_enc) is written first by the tensor-core GEMM, this
// kernel adds the rest in place.  Memory bound (one read + one write of dF).
// Grid (D/512, B); 8 warps x 64 columns; each thread owns 2 columns and keeps its dz_t values for a
// tile of up to 32 steps in registers; alpha^T sits in shared memory as [L][32].
constexpr int kDfeatTT = 32;
template <typename ST>
__global__ void __launch_bounds__(256) dfeat_accumulate_kernel(float* __restrict__ dF, const float* __restrict__ alphas,
                                                               const ST* __restrict__ DZ,
                                                               const float* __restrict__ dmeanF, int B, int L, int D,
                                                               int T, int accumulate, bf16* __restrict__ out16) {
  // out16 != null: the final sum is written as bf16 to out16 (annotations are bf16) and dF is only
  // the fp32 accumulation buffer the GEMM wrote; otherwise dF is updated in place.
  extern __shared__ __align__(16) float al_s[];   // [L][kDfeatTT]
  const int b = blockIdx.y;
  const int d = blockIdx.x * 512 + threadIdx.x * 2;
  const bool active = d < D;
  float* out = dF + (size_t)b * L * D + d;
  bf16* o16 = out16 ? out16 + (size_t)b * L * D + d : nullptr;
  const float inv_l = 1.f / (float)L;
  for (int t0 = 0; t0 < T; t0 += kDfeatTT) {
    const int tn = min(kDfeatTT, T - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < L * kDfeatTT; i += 256) {
      const int l = i / kDfeatTT, t = i - l * kDfeatTT;
      al_s[i] = t < tn ? alphas[((size_t)b * T + t0 + t) * L + l] : 0.f;
    }
    float dz0[kDfeatTT], dz1[kDfeatTT];
#pragma unroll
    for (int t = 0; t < kDfeatTT; ++t) {
      dz0[t] = 0.f; dz1[t] = 0.f;
      if (active && t < tn) {
        const ST* src = DZ + ((size_t)(t0 + t) * B + b) * D + d;
        dz0[t] = to_f<ST>(src[0]);
        dz1[t] = to_f<ST>(src[1]);
      }
    }
    float m0 = 0.f, m1 = 0.f;
    if (active && t0 == 0 && dmeanF) {
      m0 = dmeanF[(size_t)b * D + d] * inv_l;
      m1 = dmeanF[(size_t)b * D + d + 1] * inv_l;
    }
    __syncthreads();
    if (active) {
      const bool rmw = accumulate || t0 > 0;
      constexpr int RB = 7;   // rows in flight per thread (the kernel is a read-modify-write stream)
      for (int l0 = 0; l0 < L; l0 += RB) {
        float2 acc[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          acc[r] = make_float2(0.f, 0.f);
          if (rmw && l0 + r < L) acc[r] = *reinterpret_cast<const float2*>(out + (size_t)(l0 + r) * D);
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          if (l0 + r >= L) continue;
          float2 a2 = acc[r];
          a2.x += m0; a2.y += m1;
          const float4* ar = reinterpret_cast<const float4*>(al_s + (l0 + r) * kDfeatTT);
#pragma unroll
          for (int t4 = 0; t4 < kDfeatTT / 4; ++t4) {
            const float4 a = ar[t4];
            a2.x = fmaf(a.x, dz0[4 * t4], a2.x);     a2.y = fmaf(a.x, dz1[4 * t4], a2.y);
            a2.x = fmaf(a.y, dz0[4 * t4 + 1], a2.x); a2.y = fmaf(a.y, dz1[4 * t4 + 1], a2.y);
            a2.x = fmaf(a.z, dz0[4 * t4 + 2], a2.x); a2.y = fmaf(a.z, dz1[4 * t4 + 2], a2.y);
            a2.x = fmaf(a.w, dz0[4 * t4 + 3], a2.x); a2.y = fmaf(a.w, dz1[4 * t4 + 3], a2.y);
          }
          if (o16 && t0 + kDfeatTT >= T)
            *reinterpret_cast<__nv_bfloat162*>(o16 + (size_t)(l0 + r) * D) = __floats2bfloat162_rn(a2.x, a2.y);
          else
            *reinterpret_cast<float2*>(out + (size_t)(l0 + r) * D) = a2;
        }
      }
    }
  }
}

template <typename ST>
inline int launch_dfeat_accumulate(float* dF, const float* alphas, const ST* DZ, const float* dmeanF, int B, int L,
                                   int D, int T, int accumulate, bf16* out16, cudaStream_t st) {
  static DeviceOnce attr_set;
  if (int dev_ = 0; attr_set.need(&dev_)) {
    DIC_CUDA(cudaFuncSetAttribute(dfeat_accumulate_kernel<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set.mark(dev_);
  }
  ProfScope prof(P_DFEAT, st, 2.0 * (double)B * L * D * sizeof(float));
  dim3 grid(cdiv(D, 512), B);
  dfeat_accumulate_kernel<ST><<<grid, 256, sizeof(float) * L * kDfeatTT, st>>>(dF, alphas, DZ, dmeanF, B, L, D, T,
                                                                            accumulate, out16);
  DIC_LAUNCH_CHECK();
  return 0;
}

}  // namespace dic
