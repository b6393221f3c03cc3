// Fused additive-attention step kernels (the HBM-bound heart of the decoder timestep).
//
// Forward, one launch per timestep, replaces the ATen sequence
//   add, relu_, matmul(+bias), softmax, mul -> [B,L,D] temporary, sum, f_beta sigmoid, mul, cat
// of attention.py:84-93 + depth_models.py:189-192 with a single pass over the annotations:
//   e[l]  = relu(att1[l,:] + att2) . w_full + b_full
//   alpha = softmax_L(e)        | softmax((e+g)/temp)   | one_hot(argmax(e+g))
//   z     = sum_l alpha[l] F[l,:]          (streamed, 16-byte coalesced loads, fp32 registers)
//   zg    = beta * z            -> written straight into the LSTM input row ([emb | zg | h])
// att1 = F.W_enc^T + b_enc is loop invariant and hoisted out of the time loop (K0).
//
// Grid = (D chunks of 512 columns, images).  A CTA serves the KB rows (beams) that share one
// image, so the annotations are read once per image-step, not once per beam.
#pragma once
#include "common.cuh"

namespace dic {

constexpr int kAttnThreads = 256;
constexpr int kAttnDChunk = 512;   // columns per CTA (64 threads x 8 columns)
constexpr int kAttnRowGroups = kAttnThreads / (kAttnDChunk / 8);  // 4

struct AttnFwdArgs {
  const void* F;        // [images, L, D] ST
  const void* att1;     // [images, L, A] ST
  const float* hp;      // [rows, A+D] fp32: att2 | beta
  const float* w_full;  // [A]
  const float* b_full;  // [1]
  const float* u;       // [rows, L] uniform draws or null
  float* alpha_out;     // row r at alpha_out + r*alpha_stride, or null
  long long alpha_stride;
  float* z_out;         // [rows, D] fp32 or null (saved for backward)
  void* zg_out;         // ST, row r at zg_out + r*zg_stride
  long long zg_stride;
  int L, D, A;
  int mode;
  float inv_temp;
};

inline size_t attn_fwd_smem_bytes(int L, int A, int KB) {
  // w[A] | att2[KB][A] | e[KB][L] | part[3][KB][DCH] | idx[KB]
  const size_t Lp = (size_t)(L + 3) & ~(size_t)3;
  return sizeof(float) * ((size_t)A + (size_t)KB * A + (size_t)KB * Lp +
                          (size_t)(kAttnRowGroups - 1) * KB * kAttnDChunk) + sizeof(int) * 8;
}

template <typename ST, int KB>
__global__ void __launch_bounds__(kAttnThreads) attn_step_kernel(const AttnFwdArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* w_s = reinterpret_cast<float*>(smem_raw);
  float* att2_s = w_s + p.A;                 // [KB][A]
  float* e_s = att2_s + KB * p.A;            // [KB][L]
  const int Lp = (p.L + 3) & ~3;             // row pitch of e_s (keeps part_s 16-byte aligned)
  float* part_s = e_s + KB * Lp;             // [3][KB][DCH]
  int* pos_s = reinterpret_cast<int*>(part_s + (kAttnRowGroups - 1) * KB * kAttnDChunk);

  const int img = blockIdx.y;
  const int d0 = blockIdx.x * kAttnDChunk;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = p.L, D = p.D, A = p.A;
  const int row0 = img * KB;

  for (int i = tid; i < A; i += kAttnThreads) w_s[i] = p.w_full[i];
  for (int i = tid; i < KB * A; i += kAttnThreads) {
    const int j = i / A, a = i - j * A;
    att2_s[i] = p.hp[(size_t)(row0 + j) * (A + D) + a];
  }
  __syncthreads();

  // ---- phase 1: energies.  One warp per annotation row, lanes across A (4 columns each).
  const ST* att1 = reinterpret_cast<const ST*>(p.att1) + (size_t)img * L * A;
  const float b_full = p.b_full[0];
  for (int l = warp; l < L; l += kAttnThreads / 32) {
    float acc[KB];
#pragma unroll
    for (int j = 0; j < KB; ++j) acc[j] = 0.f;
    for (int a = lane * 4; a < A; a += 128) {
      float v[4];
      if (sizeof(ST) == 2) {
        uint2 r = *reinterpret_cast<const uint2*>(att1 + (size_t)l * A + a);
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&r);
        float2 f0 = __bfloat1622float2(h2[0]), f1 = __bfloat1622float2(h2[1]);
        v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
      } else {
        float4 r = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(att1) + (size_t)l * A + a);
        v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
      }
#pragma unroll
      for (int j = 0; j < KB; ++j) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          acc[j] = fmaf(fmaxf(v[q] + att2_s[j * A + a + q], 0.f), w_s[a + q], acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < KB; ++j) {
      const float s = warp_sum(acc[j]);
      if (lane == 0) e_s[j * Lp + l] = s + b_full;
    }
  }
  __syncthreads();

  // ---- phase 2: alpha.  Warp j normalises row j (KB <= 8 warps).
  if (warp < KB) {
    const int j = warp;
    float* e = e_s + j * Lp;
    const float* u = p.u ? p.u + (size_t)(row0 + j) * L : nullptr;
    if (p.mode == DIC_ATTN_GUMBEL_MAX) {
      // one_hot(argmax(e + g)), g = -log(-log u); ties -> lowest index (torch.argmax)
      float best = -INFINITY;
      int bi = 0x7fffffff;
      for (int l = lane; l < L; l += 32) {
        const float v = e[l] + (-logf(-logf(u[l])));
        if (v > best) { best = v; bi = l; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (bi == 0x7fffffff) bi = 0;   // all NaN / -inf guard
      for (int l = lane; l < L; l += 32) e[l] = (l == bi) ? 1.f : 0.f;
      if (lane == 0) pos_s[j] = bi;
    } else {
      float m = -INFINITY;
      for (int l = lane; l < L; l += 32) {
        float v = e[l];
        if (p.mode == DIC_ATTN_GUMBEL_SOFTMAX) v = (v + (-logf(-logf(u[l])))) * p.inv_temp;
        e[l] = v;
        m = fmaxf(m, v);
      }
      m = warp_max(m);
      float s = 0.f;
      for (int l = lane; l < L; l += 32) {
        const float v = expf(e[l] - m);
        e[l] = v;
        s += v;
      }
      s = warp_sum(s);
      const float inv = 1.f / s;
      for (int l = lane; l < L; l += 32) e[l] *= inv;
    }
    if (blockIdx.x == 0 && p.alpha_out) {
      float* ao = p.alpha_out + (size_t)(row0 + j) * p.alpha_stride;
      for (int l = lane; l < L; l += 32) ao[l] = e[l];
    }
  }
  __syncthreads();

  // ---- phase 3: context over this CTA's D chunk.  64 threads x 8 columns per row pass,
  // 4 row groups; each thread keeps KB x 8 fp32 accumulators.
  const int cg = tid & 63, rg = tid >> 6;
  const int d = d0 + cg * 8;
  const bool active = d < D;
  float acc[KB][8];
#pragma unroll
  for (int j = 0; j < KB; ++j)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[j][q] = 0.f;

  const ST* F = reinterpret_cast<const ST*>(p.F) + (size_t)img * L * D;
  if (active) {
    if (p.mode == DIC_ATTN_GUMBEL_MAX) {
      // one-hot alpha: the weighted sum is a single-row gather (row group 0 only)
      if (rg == 0) {
#pragma unroll
        for (int j = 0; j < KB; ++j) {
          float v[8];
          load8<ST>(F + (size_t)pos_s[j] * D + d, v);
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[j][q] = v[q];
        }
      }
    } else {
      constexpr int UNR = 7;
      int l = rg;
      for (; l + (UNR - 1) * kAttnRowGroups < L; l += UNR * kAttnRowGroups) {
        float v[UNR][8];
#pragma unroll
        for (int r = 0; r < UNR; ++r)
          load8_stream<ST>(F + (size_t)(l + r * kAttnRowGroups) * D + d, v[r]);
#pragma unroll
        for (int r = 0; r < UNR; ++r) {
#pragma unroll
          for (int j = 0; j < KB; ++j) {
            const float al = e_s[j * Lp + l + r * kAttnRowGroups];
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[j][q] = fmaf(al, v[r][q], acc[j][q]);
          }
        }
      }
      for (; l < L; l += kAttnRowGroups) {
        float v[8];
        load8_stream<ST>(F + (size_t)l * D + d, v);
#pragma unroll
        for (int j = 0; j < KB; ++j) {
          const float al = e_s[j * Lp + l];
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[j][q] = fmaf(al, v[q], acc[j][q]);
        }
      }
    }
  }
  // deterministic cross-group reduction: groups 1..3 park their partials, group 0 adds in order
  if (rg > 0) {
#pragma unroll
    for (int j = 0; j < KB; ++j) {
      float* dst = part_s + ((size_t)(rg - 1) * KB + j) * kAttnDChunk + cg * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[j][4], acc[j][5], acc[j][6], acc[j][7]);
    }
  }
  __syncthreads();
  if (rg == 0 && active) {
#pragma unroll
    for (int j = 0; j < KB; ++j) {
#pragma unroll
      for (int g = 0; g < kAttnRowGroups - 1; ++g) {
        const float* src = part_s + ((size_t)g * KB + j) * kAttnDChunk + cg * 8;
        const float4 a = *reinterpret_cast<const float4*>(src);
        const float4 b = *reinterpret_cast<const float4*>(src + 4);
        acc[j][0] += a.x; acc[j][1] += a.y; acc[j][2] += a.z; acc[j][3] += a.w;
        acc[j][4] += b.x; acc[j][5] += b.y; acc[j][6] += b.z; acc[j][7] += b.w;
      }
      const int row = row0 + j;
      if (p.z_out) store8<float>(p.z_out + (size_t)row * D + d, acc[j]);
      float beta[8];
      load8<float>(p.hp + (size_t)row * (A + D) + A + d, beta);
      float zg[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) zg[q] = beta[q] * acc[j][q];
      store8<ST>(reinterpret_cast<ST*>(p.zg_out) + (size_t)row * p.zg_stride + d, zg);
    }
  }
}

template <typename ST>
inline int launch_attn_step(const AttnFwdArgs& p, int images, int KB, cudaStream_t st) {
  if (images <= 0) return 0;
  dim3 grid(cdiv(p.D, kAttnDChunk), images);
  const size_t smem = attn_fwd_smem_bytes(p.L, p.A, KB);
  // algorithmic bytes: annotations + att1 once per image-step
  ProfScope prof(P_ATTN_FWD, st, (double)images * p.L * ((double)p.D + p.A) * sizeof(ST));
#define DIC_ATTN_CASE(K)                                                                          \
  case K: {                                                                                       \
    static bool attr_set = false;                                                                 \
    if (!attr_set) {                                                                              \
      DIC_CUDA(cudaFuncSetAttribute(attn_step_kernel<ST, K>,                                      \
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));    \
      attr_set = true;                                                                            \
    }                                                                                             \
    attn_step_kernel<ST, K><<<grid, kAttnThreads, smem, st>>>(p);                                 \
    break;                                                                                        \
  }
  switch (KB) {
    DIC_ATTN_CASE(1)
    DIC_ATTN_CASE(2)
    DIC_ATTN_CASE(3)
    DIC_ATTN_CASE(4)
    DIC_ATTN_CASE(5)
    DIC_ATTN_CASE(6)
    DIC_ATTN_CASE(7)
    DIC_ATTN_CASE(8)
    default:
      DIC_FAIL(-4, "attn_step: rows per image %d not in 1..8", KB);
  }
#undef DIC_ATTN_CASE
  DIC_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Backward of one attention step (training, one row per image).
//   dz      = dzg * beta                  -> DZ (ST), consumed post-loop by dF += alpha^T dz
//   dbeta'  = dzg * z * beta (1-beta)     -> G[:, gcol_beta + d]     (ST)
//   dalpha  = F . dz (+ external d_alphas)            (second pass over the annotations)
//   de      = alpha * (dalpha - sum alpha dalpha) * inv_temp         -> de_out
//   datt2   = w * sum_l de[l] 1[pre>0]    -> G[:, gcol_att2 + a]     (ST)
//   dw_full partial = sum_l de[l] relu(pre[l,:]) ; db_full partial = sum_l de[l]
// ------------------------------------------------------------------------------------------
constexpr int kAttnBwdThreads = 512;

struct AttnBwdArgs {
  const void* F;         // [B, L, D] ST
  const void* att1;      // [B, L, A] ST
  const float* hp;       // [B, A+D] fp32 att2 | beta   (this step)
  const float* z;        // [B, D] fp32                 (this step)
  const float* dzg;      // [B, D] fp32
  const float* alpha;    // row b at alpha + b*alpha_stride
  long long alpha_stride;
  const float* dalpha;   // external gradient, same addressing, or null
  const float* w_full;   // [A]
  void* G;               // ST [B, g_stride]
  long long g_stride;
  int gcol_att2, gcol_beta;
  void* DZ;              // ST [B, D]
  float* de_out;         // [B, L]
  float* dwfull_part;    // [B, A]
  float* dbfull_part;    // [B]
  int L, D, A;
  float inv_temp;
};

inline size_t attn_bwd_smem_bytes(int L, int D, int A) {
  // dz[D] | dal[L] | al[L] | w[A] | att2[A] | red[2][4][A] | scratch[64]
  return sizeof(float) * ((size_t)D + 2 * (size_t)L + 2 * (size_t)A + 8 * (size_t)A + 64);
}

template <typename ST>
__global__ void __launch_bounds__(kAttnBwdThreads) attn_bwd_kernel(const AttnBwdArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int L = p.L, D = p.D, A = p.A;
  float* dz_s = reinterpret_cast<float*>(smem_raw);
  float* dal_s = dz_s + D;
  float* al_s = dal_s + L;
  float* w_s = al_s + L;
  float* att2_s = w_s + A;
  float* red_s = att2_s + A;       // [2][4][A]
  float* scratch = red_s + 8 * A;  // 64

  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* hp = p.hp + (size_t)b * (A + D);
  ST* G = reinterpret_cast<ST*>(p.G) + (size_t)b * p.g_stride;

  for (int d = tid; d < D; d += kAttnBwdThreads) {
    const float beta = hp[A + d];
    const float g = p.dzg[(size_t)b * D + d];
    const float zz = p.z[(size_t)b * D + d];
    const float dz = g * beta;
    dz_s[d] = dz;
    reinterpret_cast<ST*>(p.DZ)[(size_t)b * D + d] = from_f<ST>(dz);
    G[p.gcol_beta + d] = from_f<ST>(g * zz * beta * (1.f - beta));
  }
  for (int a = tid; a < A; a += kAttnBwdThreads) {
    w_s[a] = p.w_full[a];
    att2_s[a] = hp[a];
  }
  for (int l = tid; l < L; l += kAttnBwdThreads) al_s[l] = p.alpha[(size_t)b * p.alpha_stride + l];
  __syncthreads();

  // dalpha[l] = F[l,:] . dz  -- one warp per annotation row, 16-byte loads
  const ST* F = reinterpret_cast<const ST*>(p.F) + (size_t)b * L * D;
  for (int l = warp; l < L; l += kAttnBwdThreads / 32) {
    float s = 0.f;
    for (int d = lane * 8; d < D; d += 256) {
      float v[8];
      load8_stream<ST>(F + (size_t)l * D + d, v);
#pragma unroll
      for (int q = 0; q < 8; ++q) s = fmaf(v[q], dz_s[d + q], s);
    }
    s = warp_sum(s);
    if (lane == 0) {
      if (p.dalpha) s += p.dalpha[(size_t)b * p.alpha_stride + l];
      dal_s[l] = s;
    }
  }
  __syncthreads();

  // softmax backward
  float part = 0.f;
  for (int l = tid; l < L; l += kAttnBwdThreads) part += al_s[l] * dal_s[l];
  const float dot = block_sum(part, scratch);
  float desum = 0.f;
  for (int l = tid; l < L; l += kAttnBwdThreads) {
    const float de = al_s[l] * (dal_s[l] - dot) * p.inv_temp;
    dal_s[l] = de;  // reuse as de
    p.de_out[(size_t)b * L + l] = de;
    desum += de;
  }
  const float dbf = block_sum(desum, scratch);  // includes the __syncthreads that publishes de
  if (tid == 0) p.dbfull_part[b] = dbf;

  // relu mask pass over att1: thread (a, row group)
  const ST* att1 = reinterpret_cast<const ST*>(p.att1) + (size_t)b * L * A;
  const int rg = tid >> 7, a0 = tid & 127;   // 4 row groups x 128 columns
  for (int ab = 0; ab < A; ab += 128) {
    const int a = ab + a0;
    float s1 = 0.f, s2 = 0.f;
    if (a < A) {
      const float a2 = att2_s[a];
      for (int l = rg; l < L; l += 4) {
        const float pre = to_f<ST>(att1[(size_t)l * A + a]) + a2;
        if (pre > 0.f) {
          const float de = dal_s[l];
          s1 += de;
          s2 = fmaf(de, pre, s2);
        }
      }
      red_s[(0 * 4 + rg) * A + a] = s1;
      red_s[(1 * 4 + rg) * A + a] = s2;
    }
  }
  __syncthreads();
  for (int a = tid; a < A; a += kAttnBwdThreads) {
    const float s1 = red_s[0 * A + a] + red_s[1 * A + a] + red_s[2 * A + a] + red_s[3 * A + a];
    const float s2 = red_s[4 * A + a] + red_s[5 * A + a] + red_s[6 * A + a] + red_s[7 * A + a];
    G[p.gcol_att2 + a] = from_f<ST>(w_s[a] * s1);
    p.dwfull_part[(size_t)b * A + a] = s2;
  }
}

template <typename ST>
inline int launch_attn_bwd(const AttnBwdArgs& p, int rows, cudaStream_t st) {
  if (rows <= 0) return 0;
  const size_t smem = attn_bwd_smem_bytes(p.L, p.D, p.A);
  static bool attr_set = false;
  if (!attr_set) {
    DIC_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  200 * 1024));
    attr_set = true;
  }
  ProfScope prof(P_ATTN_BWD, st, (double)rows * p.L * ((double)p.D + p.A) * sizeof(ST));
  attn_bwd_kernel<ST><<<rows, kAttnBwdThreads, smem, st>>>(p);
  DIC_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Post-loop: datt1[b,l,a] = w[a] * sum_t de_t[b,l] 1[att1[b,l,a] + att2_t[b,a] > 0]
// (the relu output is recomputed from att1 + att2 instead of being saved per step: saving it
// would cost B*L*A*4 bytes per step, SURVEY.md section 7 hard part 9).
// Grid (L chunks of 32 rows, B); 256 threads = 2 row groups x 128 columns.
// ------------------------------------------------------------------------------------------
struct Datt1Args {
  const void* att1;     // [B,L,A] ST
  const float* hp_all;  // [T,B,A+D] fp32
  const float* de_all;  // [T,B,L] fp32
  const float* w_full;  // [A]
  void* datt1;          // [B,L,A] ST
  int B, L, D, A, T;
  StepSizes sizes;
};

template <typename ST>
__global__ void __launch_bounds__(256) datt1_kernel(const Datt1Args p) {
  constexpr int RPT = 16;  // rows per thread
  const int b = blockIdx.y;
  const int l0 = blockIdx.x * (2 * RPT);
  const int rg = threadIdx.x >> 7, a0 = threadIdx.x & 127;
  int Tb = 0;
  for (int t = 0; t < p.T; ++t) Tb += (p.sizes.n[t] > b) ? 1 : 0;
  const ST* att1 = reinterpret_cast<const ST*>(p.att1) + (size_t)b * p.L * p.A;
  ST* out = reinterpret_cast<ST*>(p.datt1) + (size_t)b * p.L * p.A;
  for (int ab = 0; ab < p.A; ab += 128) {
    const int a = ab + a0;
    if (a >= p.A) continue;
    float v[RPT], acc[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int l = l0 + rg + 2 * r;
      v[r] = l < p.L ? to_f<ST>(att1[(size_t)l * p.A + a]) : 0.f;
      acc[r] = 0.f;
    }
    for (int t = 0; t < Tb; ++t) {
      const float a2 = p.hp_all[((size_t)t * p.B + b) * (p.A + p.D) + a];
      const float* de = p.de_all + ((size_t)t * p.B + b) * p.L;
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const int l = l0 + rg + 2 * r;
        if (l < p.L && v[r] + a2 > 0.f) acc[r] += de[l];
      }
    }
    const float w = p.w_full[a];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int l = l0 + rg + 2 * r;
      if (l < p.L) out[(size_t)l * p.A + a] = from_f<ST>(w * acc[r]);
    }
  }
}

template <typename ST>
inline int launch_datt1(const Datt1Args& p, cudaStream_t st) {
  dim3 grid(cdiv(p.L, 32), p.B);
  ProfScope prof(P_DATT1, st, (double)p.B * p.L * p.A * sizeof(ST) * 2);
  datt1_kernel<ST><<<grid, 256, 0, st>>>(p);
  DIC_LAUNCH_CHECK();
  return 0;
}

}  // namespace dic
