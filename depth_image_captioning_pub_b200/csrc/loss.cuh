// Fused caption-loss head: the training loop's
//   loss = F.cross_entropy(packed_logits, packed_targets, ignore_index=<null>)
//        + lam * ((1 - alphas.sum(dim=1)) ** 2).mean()           (depth_train.py:210-216)
// and its gradients w.r.t. the logits and the attention weights, in four launches, one pass over the
// logits.  The ATen sequence it replaces (log_softmax forward + backward, nll forward + backward, the
// bf16 copy of d_logits for the tensor-core GEMMs, the regulariser's elementwise chain) moved ~1.2 GB
// per step; this moves the logits once in and d_logits once out.
#pragma once
#include "common.cuh"

namespace dic {

struct LossArgs {
  const void* logits;       // [N, V] packed (time-major) rows, fp32 or (bf16 mode) bf16
  int logits_bf16;
  const int64_t* captions;  // [B, cap_stride]; target of packed row (t, b) is captions[b, t+1]
  int cap_stride;
  StepSizes sizes;
  int T, B, N, V, L;
  int ignore_index;
  const float* alphas;      // [B, T, L] or null
  float lam;
  float* loss;              // [1]
  void* d_logits;           // [N, V] ST (may alias logits when ST = float)
  float* d_alphas;          // [B, T, L] or null
  float* nll;               // [N]   workspace
  float* count;             // [1]   workspace: number of non-ignored targets
  float* regsq;             // [B]   workspace
};

__device__ __forceinline__ int loss_target(const LossArgs& p, int r) {
  int t = 0, off = 0;
  while (t + 1 < p.T && r >= off + p.sizes.n[t]) { off += p.sizes.n[t]; ++t; }
  const int b = r - off;
  return (int)p.captions[(size_t)b * p.cap_stride + t + 1];
}

// number of non-ignored targets (the mean's denominator), one CTA
__global__ void __launch_bounds__(1024) loss_count_kernel(const LossArgs p) {
  pdl_wait();          // launched programmatically (launch_pdl): nothing of a predecessor is touched before this
  pdl_trigger();
  __shared__ float scratch[64];
  float c = 0.f;
  for (int r = threadIdx.x; r < p.N; r += 1024) c += (loss_target(p, r) != p.ignore_index) ? 1.f : 0.f;
  c = block_sum(c, scratch);
  if (threadIdx.x == 0) p.count[0] = c;
}

// one CTA per packed row: row staged in shared memory with 16-byte loads, log-sum-exp, nll,
// d_logits = (softmax - onehot) / count written in the storage dtype of the mode
constexpr int kCeThreads = 256;     // threads per logits row (512 measured 7% slower)
template <typename ST>
__global__ void __launch_bounds__(kCeThreads) loss_ce_row_kernel(const LossArgs p, int staged) {
  pdl_wait();          // launched programmatically (launch_pdl): nothing of a predecessor is touched before this
  pdl_trigger();
  extern __shared__ __align__(16) float row_s[];
  __shared__ float scratch[64];
  const int r = blockIdx.x, tid = threadIdx.x, V = p.V;
  const float* lg = reinterpret_cast<const float*>(p.logits) + (size_t)r * V;        // fp32 view (when not bf16)
  const int tgt = loss_target(p, r);
  const bool valid = tgt != p.ignore_index;
  ST* out = reinterpret_cast<ST*>(p.d_logits) + (size_t)r * V;
  if (staged) {
    // one trip to global memory: the row is staged with 16-byte loads while the running max is taken;
    // exp(x - max) is computed once and kept in shared memory for the gradient pass
    float m = -INFINITY;
    if (p.logits_bf16) {
      // bf16 logits (fused training step in bf16 mode): 16-byte loads of 8 values, widened into the fp32 row
      const bf16* lb = reinterpret_cast<const bf16*>(p.logits) + (size_t)r * V;
      if ((V & 7) == 0) {
        const int n8 = V / 8;
        for (int i0 = tid; i0 < n8; i0 += 4 * kCeThreads) {        // four 16-byte loads in flight per thread
          uint4 raw[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (i0 + u * kCeThreads < n8) raw[u] = *reinterpret_cast<const uint4*>(lb + (size_t)(i0 + u * kCeThreads) * 8);   // plain loads (in-place d_logits)
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * kCeThreads;
            if (i < n8) {
              const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[u]);
              const float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
              const float2 f2 = __bfloat1622float2(h[2]), f3 = __bfloat1622float2(h[3]);
              *reinterpret_cast<float4*>(row_s + i * 8) = make_float4(f0.x, f0.y, f1.x, f1.y);
              *reinterpret_cast<float4*>(row_s + i * 8 + 4) = make_float4(f2.x, f2.y, f3.x, f3.y);
              m = fmaxf(m, fmaxf(fmaxf(fmaxf(f0.x, f0.y), fmaxf(f1.x, f1.y)), fmaxf(fmaxf(f2.x, f2.y), fmaxf(f3.x, f3.y))));
            }
          }
        }
      } else {
        for (int v = tid; v < V; v += kCeThreads) { const float x = __bfloat162float(lb[v]); row_s[v] = x; m = fmaxf(m, x); }
      }
    } else if ((V & 3) == 0) {
      const float4* src4 = reinterpret_cast<const float4*>(lg);
      float4* dst4 = reinterpret_cast<float4*>(row_s);
      const int n4 = V / 4;
      int i = tid;
      for (; i + 3 * kCeThreads < n4; i += 4 * kCeThreads) {      // four 16-byte loads in flight per thread
        float4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = src4[i + u * kCeThreads];   // plain loads: the row may be overwritten below
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          dst4[i + u * kCeThreads] = x[u];
          m = fmaxf(fmaxf(m, fmaxf(x[u].x, x[u].y)), fmaxf(x[u].z, x[u].w));
        }
      }
      for (; i < n4; i += kCeThreads) {
        const float4 x = src4[i];
        dst4[i] = x;
        m = fmaxf(fmaxf(m, fmaxf(x.x, x.y)), fmaxf(x.z, x.w));
      }
    } else {
      for (int v = tid; v < V; v += kCeThreads) {
        const float x = lg[v];
        row_s[v] = x;
        m = fmaxf(m, x);
      }
    }
    m = block_max(m, scratch);               // its barriers also publish row_s
    const float x_tgt = valid ? ((tgt >= 0 && tgt < V) ? row_s[tgt] : nanf("")) : 0.f;    // out-of-vocabulary target: NaN loss
    __syncthreads();                         // everyone has read x_tgt before row_s is overwritten
    float s = 0.f;
    if ((V & 3) == 0) {
      float4* r4 = reinterpret_cast<float4*>(row_s);
      for (int i = tid; i < V / 4; i += kCeThreads) {
        float4 x = r4[i];
        // bf16 mode: ex2.approx exponentials (the result is rounded to bf16 below); fp32 parity mode: expf
        if constexpr (sizeof(ST) == 2) { x.x = __expf(x.x - m); x.y = __expf(x.y - m); x.z = __expf(x.z - m); x.w = __expf(x.w - m); }
        else { x.x = expf(x.x - m); x.y = expf(x.y - m); x.z = expf(x.z - m); x.w = expf(x.w - m); }
        r4[i] = x;
        s += (x.x + x.y) + (x.z + x.w);
      }
    } else {
      for (int v = tid; v < V; v += kCeThreads) {
        const float e = expf(row_s[v] - m);
        row_s[v] = e;
        s += e;
      }
    }
    s = block_sum(s, scratch);
    if (tid == 0) p.nll[r] = valid ? (m + logf(s) - x_tgt) : 0.f;
    const float scale = valid ? 1.f / p.count[0] : 0.f;
    const float ps = scale / s;              // softmax * scale = e * ps
    if ((V & 7) == 0) {
      for (int v0 = tid * 8; v0 < V; v0 += kCeThreads * 8) {
        const float4 e0 = *reinterpret_cast<const float4*>(row_s + v0);
        const float4 e1 = *reinterpret_cast<const float4*>(row_s + v0 + 4);
        float g[8] = {e0.x * ps, e0.y * ps, e0.z * ps, e0.w * ps, e1.x * ps, e1.y * ps, e1.z * ps, e1.w * ps};
        if (tgt >= v0 && tgt < v0 + 8) g[tgt - v0] -= scale;
        store8<ST>(out + v0, g);
      }
    } else {
      for (int v = tid; v < V; v += kCeThreads) out[v] = from_f<ST>(row_s[v] * ps - ((v == tgt) ? scale : 0.f));
    }
    return;
  }
  // rows too long for shared memory: three passes over global memory
  float m = -INFINITY;
  for (int v = tid; v < V; v += kCeThreads) m = fmaxf(m, lg[v]);
  m = block_max(m, scratch);
  float s = 0.f;
  for (int v = tid; v < V; v += kCeThreads) s += expf(lg[v] - m);
  s = block_sum(s, scratch);
  const float lse = m + logf(s);
  if (tid == 0) p.nll[r] = valid ? ((tgt >= 0 && tgt < V) ? (lse - lg[tgt]) : nanf("")) : 0.f;
  const float scale = valid ? 1.f / p.count[0] : 0.f;
  for (int v = tid; v < V; v += kCeThreads)
    out[v] = from_f<ST>((expf(lg[v] - lse) - ((v == tgt) ? 1.f : 0.f)) * scale);
}

// bf16 logits, V <= 256 x 40 (the reference's 10k vocabulary): the row lives in REGISTERS -- five 16-byte
// loads of 8 bf16 per thread -- so a row costs two block reductions and nothing else between its load and
// its store: no shared-memory staging, 2 instead of 8 barriers, more rows resident per SM.  The shared-memory
// kernel above took 80 us for 102 MB in + 102 MB out (its CTAs are latency chains of four barrier-separated
// passes with 5 resident per SM).
constexpr int kCeRegChunks = 5;     // 16-byte chunks per thread: 256 threads x 5 x 8 = 10240 columns
__global__ void __launch_bounds__(256) loss_ce_row_reg_kernel(const LossArgs p) {
  pdl_wait();          // launched programmatically (launch_pdl): nothing of a predecessor is touched before this
  pdl_trigger();
  __shared__ float scratch[64];
  const int r = blockIdx.x, tid = threadIdx.x, V = p.V;
  const bf16* lb = reinterpret_cast<const bf16*>(p.logits) + (size_t)r * V;
  bf16* out = reinterpret_cast<bf16*>(p.d_logits) + (size_t)r * V;
  const int tgt = loss_target(p, r);
  const bool valid = tgt != p.ignore_index;
  const int n8 = V / 8;
  const bool in_range = tgt >= 0 && tgt < V;
  const bool owner = valid && in_range && (tgt / 8) % 256 == tid;     // the thread whose chunk holds the target column
  const float x_tgt = owner ? __bfloat162float(lb[tgt]) : 0.f;        // read up front: d_logits is written in place
  uint4 raw[kCeRegChunks];
#pragma unroll
  for (int u = 0; u < kCeRegChunks; ++u) {
    const int i = tid + u * 256;
    if (i < n8) raw[u] = *reinterpret_cast<const uint4*>(lb + (size_t)i * 8);      // plain loads (in-place d_logits)
  }
  // The loops below were instruction-issue bound (SASS: ~20 instructions per logit -- a per-element target compare in
  // two passes, a per-element padding select, __expf = three multiplies + compare + ex2 -- 46 us for a pass whose
  // HBM floor is 32 us).  Now: the target column is handled by its owner thread outside the loops, padding chunks
  // are filled once, exp(x - m) is one FMA into ex2.approx.
  float x[kCeRegChunks][8];
  float m = -INFINITY;
#pragma unroll
  for (int u = 0; u < kCeRegChunks; ++u) {
    const int i = tid + u * 256;
    if (i < n8) {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw[u]);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __bfloat1622float2(h[q]);
        x[u][2 * q] = f.x;
        x[u][2 * q + 1] = f.y;
        m = fmaxf(m, fmaxf(f.x, f.y));
      }
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q) x[u][q] = -INFINITY;
    }
  }
  m = block_max(m, scratch);
  constexpr float kLog2e = 1.4426950408889634f;
  const float m2 = m * kLog2e;
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int u = 0; u < kCeRegChunks; ++u) {
#pragma unroll
    for (int q = 0; q < 8; q += 2) {
      const float e0 = ex2_approx(fmaf(x[u][q], kLog2e, -m2));          // ex2(-inf) = 0 for the padding lanes
      const float e1 = ex2_approx(fmaf(x[u][q + 1], kLog2e, -m2));
      x[u][q] = e0; x[u][q + 1] = e1;
      s0 += e0; s1 += e1;
    }
  }
  const float s = block_sum(s0 + s1, scratch);
  // the thread that holds the target column reports the row's nll and patches that one gradient element
  if (owner) p.nll[r] = m + logf(s) - x_tgt;
  if (tid == 0 && !valid) p.nll[r] = 0.f;
  if (tid == 0 && valid && !in_range) p.nll[r] = nanf("");     // a target outside the vocabulary is a caller bug: make it visible
  const float scale = valid ? 1.f / p.count[0] : 0.f;
  const float ps = scale / s;
#pragma unroll
  for (int u = 0; u < kCeRegChunks; ++u) {
    const int i = tid + u * 256;
    if (i < n8) {
      float g[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) g[q] = x[u][q] * ps;
      store8<bf16>(out + (size_t)i * 8, g);
    }
  }
  if (owner) out[tgt] = __float2bfloat16_rn(ex2_approx(fmaf(x_tgt, kLog2e, -m2)) * ps - scale);    // same thread, after its chunk store
}

// doubly-stochastic regulariser, one CTA per image: S[l] = sum_t alpha[b,t,l];
// regsq[b] = sum_l (1-S)^2 ; d_alpha[b,t,l] = -2 lam (1-S[l]) / (B L) for every t
__global__ void __launch_bounds__(256) loss_reg_kernel(const LossArgs p) {
  pdl_wait();          // launched programmatically (launch_pdl): nothing of a predecessor is touched before this
  pdl_trigger();
  __shared__ float scratch[64];
  const int b = blockIdx.x;
  const float* al = p.alphas + (size_t)b * p.T * p.L;
  const float gs = -2.f * p.lam / ((float)p.B * (float)p.L);
  float sq = 0.f;
  for (int l = threadIdx.x; l < p.L; l += 256) {
    float S = 0.f;
    for (int t = 0; t < p.T; ++t) S += al[(size_t)t * p.L + l];
    const float d = 1.f - S;
    sq += d * d;
    if (p.d_alphas) {
      const float g = gs * d;
      for (int t = 0; t < p.T; ++t) p.d_alphas[((size_t)b * p.T + t) * p.L + l] = g;
    }
  }
  sq = block_sum(sq, scratch);
  if (threadIdx.x == 0) p.regsq[b] = sq;
}

// loss = sum(nll)/count + lam * sum(regsq)/(B L); one CTA, fixed reduction order
__global__ void __launch_bounds__(1024) loss_final_kernel(const LossArgs p) {
  pdl_wait();          // launched programmatically (launch_pdl): nothing of a predecessor is touched before this
  pdl_trigger();
  __shared__ float scratch[64];
  float a = 0.f;
  for (int r = threadIdx.x; r < p.N; r += 1024) a += p.nll[r];
  a = block_sum(a, scratch);
  float q = 0.f;
  if (p.alphas && p.lam != 0.f) {
    for (int b = threadIdx.x; b < p.B; b += 1024) q += p.regsq[b];
    q = block_sum(q, scratch);
  }
  if (threadIdx.x == 0) {
    const float cnt = p.count[0];
    float loss = cnt > 0.f ? a / cnt : nanf("");          // F.cross_entropy: all targets ignored -> nan
    if (p.alphas && p.lam != 0.f) loss += p.lam * q / ((float)p.B * (float)p.L);
    p.loss[0] = loss;
  }
}

// upstream gradient of the scalar loss (read on the device: no host sync); a no-op when it is 1
template <typename ST>
__global__ void __launch_bounds__(256) loss_scale_kernel(const float* __restrict__ g, ST* __restrict__ dl, size_t n1,
                                                         float* __restrict__ da, size_t n2) {
  const float s = g[0];
  if (s == 1.f) return;
  const size_t stride = (size_t)gridDim.x * 256;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n1; i += stride) dl[i] = from_f<ST>(to_f<ST>(dl[i]) * s);
  if (da)
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n2; i += stride) da[i] *= s;
}

inline bool ce_reg_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DIC_CE_REG"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

inline size_t loss_workspace_bytes(int N, int B) {
  return align_up(sizeof(float) * (size_t)N, 256) + 256 + align_up(sizeof(float) * (size_t)B, 256);
}

template <typename ST>
inline int launch_caption_loss(LossArgs p, void* workspace, cudaStream_t st) {
  char* ws = reinterpret_cast<char*>(workspace);
  p.nll = reinterpret_cast<float*>(ws);
  p.count = reinterpret_cast<float*>(ws + align_up(sizeof(float) * (size_t)p.N, 256));
  p.regsq = reinterpret_cast<float*>(ws + align_up(sizeof(float) * (size_t)p.N, 256) + 256);
  DIC_CUDA(launch_pdl(loss_count_kernel, dim3(1), dim3(1024), 0, st, p));
  DIC_LAUNCH_CHECK();
  const size_t row_bytes = sizeof(float) * (size_t)p.V;
  const int staged = row_bytes <= 200 * 1024 ? 1 : 0;
  if (!staged && (p.logits == p.d_logits || p.logits_bf16))
    DIC_FAIL(-4, "caption_loss: in-place / bf16 logits need V <= 51200");
  static DeviceOnce attr_set;
  if (int dev_ = 0; attr_set.need(&dev_)) {
    DIC_CUDA(cudaFuncSetAttribute(loss_ce_row_kernel<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set.mark(dev_);
  }
  {
    ProfScope prof(P_LOSS, st, (double)p.N * p.V * ((p.logits_bf16 ? 2 : 4) + sizeof(ST)));
    if (sizeof(ST) == 2 && p.logits_bf16 && p.V % 8 == 0 && p.V <= 256 * 8 * kCeRegChunks && ce_reg_enabled())
      DIC_CUDA(launch_pdl(loss_ce_row_reg_kernel, dim3(p.N), dim3(256), 0, st, p));
    else
      DIC_CUDA(launch_pdl(loss_ce_row_kernel<ST>, dim3(p.N), dim3(kCeThreads), staged ? row_bytes : 0, st, p, staged));
    DIC_LAUNCH_CHECK();
  }
  if (p.alphas && p.lam != 0.f) {
    DIC_CUDA(launch_pdl(loss_reg_kernel, dim3(p.B), dim3(256), 0, st, p));
    DIC_LAUNCH_CHECK();
  }
  DIC_CUDA(launch_pdl(loss_final_kernel, dim3(1), dim3(1024), 0, st, p));
  DIC_LAUNCH_CHECK();
  return 0;
}

}  // namespace dic
