// Additive-attention step kernels (the HBM-bound heart of the decoder timestep).
//
// Forward, per timestep, replaces the ATen sequence
//   add, relu_, matmul(+bias), softmax, mul -> [B,L,D] temporary, sum, f_beta sigmoid, mul, cat
// of attention.py:84-93 + depth_models.py:189-192 with two launches:
//   (a) attn_alpha_kernel   e[l] = relu(att1[l,:] + att2) . w_full + b_full
//                           alpha = softmax_L(e) | softmax((e+g)/temp) | one_hot(argmax(e+g))
//       one CTA per image (reads the 50 KB att1 slab once for all beams of the image);
//   (b) attn_context_kernel z = sum_l alpha[l] F[l,:],  zg = beta * z
//       a pure streaming pass over the annotations: grid = (D/256 column chunks, images),
//       128 threads, 16-byte coalesced loads with 7 rows in flight per thread, fp32 register
//       accumulators, deterministic cross-warp reduction in shared memory; zg is written
//       straight into the LSTM input row ([emb | zg | h]).
// att1 = F.W_enc^T + b_enc is loop invariant and hoisted out of the time loop (K0).
// A context CTA serves the KB rows (beams) that share one image, so the annotations are read
// once per image-step, not once per beam.
//
// (An earlier single-kernel version ran energies, softmax and the context pass in one CTA; ncu
// showed 34% of HBM peak: every CTA idled the memory system during its energy/softmax phases and
// the 1024-CTA grid was 1.15 waves.  See profiles/.)
#pragma once
#include "common.cuh"

namespace dic {

constexpr int kAlphaThreads = 512;
constexpr int kCtxThreads = 128;
constexpr int kCtxCols = 256;                       // columns per context CTA (32 lanes x 8)
constexpr int kCtxGroups = kCtxThreads / 32;        // 4 row groups (one warp each)
constexpr int kCtxUnroll = 7;                       // rows in flight per thread

struct AttnFwdArgs {
  const void* F;        // [images, L, D] ST
  const void* att1;     // [images, L, A] ST
  const float* hp;      // [rows, A+D] fp32: att2 | beta
  const float* w_full;  // [A]
  const float* b_full;  // [1]
  const float* u;       // [rows, L] uniform draws or null
  float* alpha_out;     // row r at alpha_out + r*alpha_stride (required: (a) -> (b) hand-off)
  long long alpha_stride;
  bf16* alpha16_out;    // optional bf16 copy (A operand of the fused dL/dF GEMM), row r at + r*alpha16_stride
  long long alpha16_stride;
  int alpha16_width;    // padded row width Lp >= L: columns [L, Lp) are written as zeros
  // Per-image hand-off alpha kernel -> context kernel (optional): the alpha CTA of image i publishes `epoch`
  // in ready[i] (release) once its weights are written, and the context CTAs of image i poll it (acquire)
  // instead of waiting for the whole alpha grid to drain and flush.
  unsigned int* ready;  // [images] or null
  unsigned int epoch;
  float* z_out;         // [rows, D] fp32 or null (saved for backward)
  void* zg_out;         // ST, row r at zg_out + r*zg_stride
  long long zg_stride;
  int L, D, A;
  int mode;
  float inv_temp;
  int rpi;              // rows (beams) per image for the alpha kernel's row -> image map (0 = KB)
  int skip_alpha;       // host side: the attention weights were already written by the fused head kernel (attn_head.cuh)
  int late_wait;        // decode context kernels: the kernel launched right before this one is NOT a producer of its
                        // inputs (they come from the launch before that, which the predecessor itself waited for before
                        // it let this grid start): run concurrently with it, and make the programmatic-dependency wait
                        // the LAST thing one CTA does, so that this grid's completion still implies the predecessor's
  TraceRec* trace;
};

// Normalisation of one row of energies by one warp (in place in shared memory), then the stores of the
// attention weights: softmax | softmax((e+g)/temp) | one_hot(argmax(e+g))  (attention.py:90, :12-25, :34-48)
// FAST (bf16 storage, the fused head kernel): exp(v - m) as one FMA into ex2.approx (relative error 2^-22)
template <bool FAST = false>
__device__ __forceinline__ void attn_normalise_row(const AttnFwdArgs& p, float* e, int row, int lane) {
  const int L = p.L;
  const float* u = p.u ? p.u + (size_t)row * L : nullptr;
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
  } else if (L <= 256) {
    // the row lives in registers between the passes (8 values per lane): same operations in the same order as the
    // shared-memory loop below, without its three store -> load round trips (1.6 -> 0.9 us of the head kernel)
    float v[8];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int l = lane + 32 * i;
      v[i] = -INFINITY;
      if (l < L) {
        float x = e[l];
        if (p.mode == DIC_ATTN_GUMBEL_SOFTMAX) x = (x + (-logf(-logf(u[l])))) * p.inv_temp;
        v[i] = x;
        m = fmaxf(m, x);
      }
    }
    m = warp_max(m);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (lane + 32 * i < L) {
        v[i] = FAST ? ex2_approx(fmaf(v[i], 1.4426950408889634f, -m * 1.4426950408889634f)) : expf(v[i] - m);
        s += v[i];
      }
    }
    s = warp_sum(s);
    const float inv = 1.f / s;
    float* ao = p.alpha_out + (size_t)row * p.alpha_stride;
    bf16* a16 = p.alpha16_out ? p.alpha16_out + (size_t)row * p.alpha16_stride : nullptr;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int l = lane + 32 * i;
      const float a = l < L ? v[i] * inv : 0.f;
      if (l < L) ao[l] = a;
      if (a16 && l < p.alpha16_width) a16[l] = __float2bfloat16_rn(a);
    }
    return;
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
  float* ao = p.alpha_out + (size_t)row * p.alpha_stride;
  for (int l = lane; l < L; l += 32) ao[l] = e[l];
  if (p.alpha16_out) {
    bf16* a16 = p.alpha16_out + (size_t)row * p.alpha16_stride;
    for (int l = lane; l < p.alpha16_width; l += 32) a16[l] = __float2bfloat16_rn(l < L ? e[l] : 0.f);
  }
}

// ---------------------------------------------------------------------------------------------
// (a) energies + normalisation.  Half-warp per annotation row (16 lanes x 8 columns = 128 columns
// per pass), so one warp instruction reads two full 256-byte att1 rows.
// ---------------------------------------------------------------------------------------------
inline size_t attn_alpha_smem_bytes(int L, int A, int KB) {
  const size_t Lp = (size_t)(L + 3) & ~(size_t)3;
  return sizeof(float) * ((size_t)A + (size_t)KB * A + (size_t)KB * Lp);
}

template <typename ST, int KB>
__global__ void __launch_bounds__(kAlphaThreads) attn_alpha_kernel(const AttnFwdArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Trace trace(p.trace);
  float* w_s = reinterpret_cast<float*>(smem_raw);
  float* att2_s = w_s + p.A;                 // [KB][A]
  float* e_s = att2_s + KB * p.A;            // [KB][Lp]
  const int L = p.L, D = p.D, A = p.A;
  const int Lp = (L + 3) & ~3;
  // KB rows of one image per CTA, or (rpi > 0, KB == 1) one row per CTA with rpi rows sharing an image:
  // beam decoding at 128 images per GPU needs the 5x larger grid more than it needs the att1 reuse
  const int img = p.rpi > 0 ? blockIdx.x / p.rpi : blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = p.rpi > 0 ? blockIdx.x : img * KB;

  const ST* att1 = reinterpret_cast<const ST*>(p.att1) + (size_t)img * L * A;
  const int half = lane >> 4, hl = lane & 15;
  constexpr int RPW = 2 * (kAlphaThreads / 32);      // rows per CTA pass (32)
  constexpr int IT = 7;                              // 7 * 32 rows >= 196: 7 x 16-byte loads per thread

  // att1 is loop invariant (written once by K0): the first block of its loads is issued BEFORE the
  // programmatic-dependency wait, so their latency overlaps the tail of the kernel that produces att2.
  Raw8<ST> raw[IT];
  if (A == 128) {
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      const int l = it * RPW + warp * 2 + half;
      if (l < L) raw[it].load_stream(att1 + (size_t)l * A + hl * 8);
      else raw[it].zero();
    }
  }
  pdl_wait();
  pdl_trigger();
  trace.mark();

  for (int i = tid; i < A; i += kAlphaThreads) w_s[i] = p.w_full[i];
  for (int i = tid; i < KB * A; i += kAlphaThreads) {
    const int j = i / A, a = i - j * A;
    att2_s[i] = p.hp[(size_t)(row0 + j) * (A + D) + a];
  }
  __syncthreads();

  const float b_full = p.b_full[0];
  if (A == 128) {
    // Reference shape: one 16-byte load covers a lane's 8 columns.  The whole att1 slab of the image
    // is put in flight first (IT independent loads per thread), then consumed: the kernel is
    // a pure latency chain otherwise (one 50 KB slab per CTA).
    float w8[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) w8[q] = w_s[hl * 8 + q];
    for (int lb0 = 0; lb0 < L; lb0 += IT * RPW) {
      if (lb0 > 0) {      // L > 208: reload the (single) register block for the next 208 rows
#pragma unroll
        for (int it = 0; it < IT; ++it) {
          const int l = lb0 + it * RPW + warp * 2 + half;
          if (l < L) raw[it].load_stream(att1 + (size_t)l * A + hl * 8);
          else raw[it].zero();
        }
      }
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        const int l = lb0 + it * RPW + warp * 2 + half;
        float v[8];
        raw[it].unpack(v);
#pragma unroll
        for (int j = 0; j < KB; ++j) {
          float s = 0.f;
#pragma unroll
          for (int q = 0; q < 8; ++q) s = fmaf(fmaxf(v[q] + att2_s[j * A + hl * 8 + q], 0.f), w8[q], s);
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          if (hl == 0 && l < L) e_s[j * Lp + l] = s + b_full;
        }
      }
    }
  } else {
    for (int lb = 0; lb < L; lb += RPW) {
      const int l = lb + warp * 2 + half;
      float acc[KB];
#pragma unroll
      for (int j = 0; j < KB; ++j) acc[j] = 0.f;
      if (l < L) {
        for (int a = hl * 8; a < A; a += 128) {
          float v[8];
          if (a + 8 <= A) {
            load8<ST>(att1 + (size_t)l * A + a, v);
          } else {   // A % 8 == 4 tail
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = (a + q < A) ? to_f<ST>(att1[(size_t)l * A + a + q]) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < KB; ++j) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (a + q < A) acc[j] = fmaf(fmaxf(v[q] + att2_s[j * A + a + q], 0.f), w_s[a + q], acc[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < KB; ++j) {
        float s = acc[j];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (hl == 0 && l < L) e_s[j * Lp + l] = s + b_full;
      }
    }
  }
  __syncthreads();

  // normalisation: warp j owns row j (KB <= 8 warps)
  if (warp < KB) attn_normalise_row(p, e_s + warp * Lp, row0 + warp, lane);
  if (p.ready && p.rpi <= 1) {             // one CTA per image (rpi == 0: KB rows, rpi == 1: its single row)
    __threadfence();                       // the writers' stores are visible device-wide ...
    __syncthreads();                       // ... before the flag goes up
    if (tid == 0)
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p.ready + img), "r"(p.epoch) : "memory");
  }
  trace.end(TK_ALPHA);
}

// ---------------------------------------------------------------------------------------------
// (b) context: the streaming pass over the annotations
// ---------------------------------------------------------------------------------------------
// columns per thread / rows in flight of the context kernel: one beam per image streams with 16-byte
// loads; with several beams the accumulators (KB x CPT registers) decide the occupancy, and 8 columns
// per thread left 5 CTAs per SM = 1.4 waves of the 1024-CTA grid (32 us at 128 images x 5 beams against
// 15 us with one beam).  4 columns per thread, twice the rows in flight, keeps 8 CTAs resident.
__host__ __device__ constexpr int ctx_cpt(int KB) { return KB <= 2 ? 8 : 4; }
__host__ __device__ constexpr int ctx_unroll(int KB) { return KB <= 2 ? kCtxUnroll : (KB <= 5 ? 12 : 8); }

inline size_t attn_ctx_smem_bytes(int L, int KB) {
  const size_t Lp = (size_t)(L + 3) & ~(size_t)3;
  return sizeof(float) * ((size_t)KB * Lp + (size_t)(kCtxGroups - 1) * KB * 32 * ctx_cpt(KB)) + 32;
}

// acc[0..N) += al * v[0..N) as packed fp32x2 FMAs (sm_100 FFMA2; same rounding as fmaf per element).
// With several beams per image the pass is instruction-issue bound, not HBM bound: 5 beams x 8 columns
// = 40 scalar FMAs per 16-byte load.
template <int N>
__device__ __forceinline__ void ctx_fmaN(float al, const float (&v)[N], float (&acc)[N]) {
  const float2 a2 = make_float2(al, al);
#pragma unroll
  for (int q = 0; q < N; q += 2) {
    const float2 r = __ffma2_rn(a2, make_float2(v[q], v[q + 1]), make_float2(acc[q], acc[q + 1]));
    acc[q] = r.x;
    acc[q + 1] = r.y;
  }
}

template <typename ST, int KB>
__global__ void __launch_bounds__(kCtxThreads) attn_context_kernel(const AttnFwdArgs p) {
  constexpr int CPT = ctx_cpt(KB), U = ctx_unroll(KB), COLS = 32 * CPT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Trace trace(p.trace);
  const int L = p.L, D = p.D, A = p.A;
  const int Lp = (L + 3) & ~3;
  float* al_s = reinterpret_cast<float*>(smem_raw);          // [KB][Lp]
  float* part_s = al_s + KB * Lp;                            // [3][KB][COLS]
  int* pos_s = reinterpret_cast<int*>(part_s + (kCtxGroups - 1) * KB * COLS);

  const int img = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, rg = tid >> 5;
  const int row0 = img * KB;
  const int d = blockIdx.x * COLS + lane * CPT;
  const bool active = d < D;
  const ST* F = reinterpret_cast<const ST*>(p.F) + (size_t)img * L * D + d;

  // first batch of annotation rows goes in flight before anything else is touched
  RawV<ST, CPT> v0[U];
  const bool gmax = p.mode == DIC_ATTN_GUMBEL_MAX;
  const bool pre = active && !gmax && (rg + (U - 1) * kCtxGroups < L);
  if (pre) {
#pragma unroll
    for (int r = 0; r < U; ++r) v0[r].load_stream(F + (size_t)(rg + r * kCtxGroups) * D);
  }
  // the annotations are static; alpha and beta come from the preceding kernels of this step.  With the
  // per-image hand-off the CTA waits for ITS image's attention weights only (beta was written two launches ago:
  // the alpha CTAs passed their own dependency wait before this grid could start).
  if (p.ready) {
    if (tid == 0) {
      unsigned int v, spins = 0;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.ready + img) : "memory");
        if (++spins > (1u << 26)) __trap();
      } while (v < p.epoch);
    }
    __syncthreads();
  } else if (!p.late_wait) {
    pdl_wait();
  }
  pdl_trigger();
  trace.mark();

  for (int i = tid; i < KB * L; i += kCtxThreads) {
    const int j = i / L, l = i - j * L;
    const float a = p.alpha_out[(size_t)(row0 + j) * p.alpha_stride + l];
    al_s[j * Lp + l] = a;
    if (gmax && a == 1.f) pos_s[j] = l;
  }
  __syncthreads();

  float acc[KB][CPT];
#pragma unroll
  for (int j = 0; j < KB; ++j)
#pragma unroll
    for (int q = 0; q < CPT; ++q) acc[j][q] = 0.f;

  if (active) {
    if (gmax) {
      // one-hot alpha: the weighted sum is a single-row gather
      if (rg == 0) {
#pragma unroll
        for (int j = 0; j < KB; ++j) vloadN<ST, CPT>(F + (size_t)pos_s[j] * D, acc[j]);
      }
    } else {
      int l = rg;
      if (pre) {
#pragma unroll
        for (int r = 0; r < U; ++r) {
          float v[CPT];
          v0[r].unpack(v);
#pragma unroll
          for (int j = 0; j < KB; ++j) ctx_fmaN<CPT>(al_s[j * Lp + l + r * kCtxGroups], v, acc[j]);
        }
        l += U * kCtxGroups;
      }
      for (; l + (U - 1) * kCtxGroups < L; l += U * kCtxGroups) {
        RawV<ST, CPT> raw[U];
#pragma unroll
        for (int r = 0; r < U; ++r) raw[r].load_stream(F + (size_t)(l + r * kCtxGroups) * D);
#pragma unroll
        for (int r = 0; r < U; ++r) {
          float v[CPT];
          raw[r].unpack(v);
#pragma unroll
          for (int j = 0; j < KB; ++j) ctx_fmaN<CPT>(al_s[j * Lp + l + r * kCtxGroups], v, acc[j]);
        }
      }
      for (; l < L; l += kCtxGroups) {
        RawV<ST, CPT> raw;
        raw.load_stream(F + (size_t)l * D);
        float v[CPT];
        raw.unpack(v);
#pragma unroll
        for (int j = 0; j < KB; ++j) ctx_fmaN<CPT>(al_s[j * Lp + l], v, acc[j]);
      }
    }
  }
  // deterministic cross-warp reduction: warps 1..3 park their partials, warp 0 adds in order
  if (rg > 0) {
#pragma unroll
    for (int j = 0; j < KB; ++j) {
      float* dst = part_s + ((size_t)(rg - 1) * KB + j) * COLS + lane * CPT;
#pragma unroll
      for (int q = 0; q < CPT; q += 4)
        *reinterpret_cast<float4*>(dst + q) = make_float4(acc[j][q], acc[j][q + 1], acc[j][q + 2], acc[j][q + 3]);
    }
  }
  __syncthreads();
  if (rg == 0 && active) {
#pragma unroll
    for (int j = 0; j < KB; ++j) {
#pragma unroll
      for (int g = 0; g < kCtxGroups - 1; ++g) {
        const float* src = part_s + ((size_t)g * KB + j) * COLS + lane * CPT;
#pragma unroll
        for (int q = 0; q < CPT; q += 4) {
          const float4 a = *reinterpret_cast<const float4*>(src + q);
          acc[j][q] += a.x; acc[j][q + 1] += a.y; acc[j][q + 2] += a.z; acc[j][q + 3] += a.w;
        }
      }
      const int row = row0 + j;
      if (p.z_out) vstoreN<float, CPT>(p.z_out + (size_t)row * D + d, acc[j]);
      float beta[CPT];
      vloadN<float, CPT>(p.hp + (size_t)row * (A + D) + A + d, beta);
      float zg[CPT];
#pragma unroll
      for (int q = 0; q < CPT; ++q) zg[q] = beta[q] * acc[j][q];
      vstoreN<ST, CPT>(reinterpret_cast<ST*>(p.zg_out) + (size_t)row * p.zg_stride + d, zg);
    }
  }
  trace.end(TK_CTX);
  if (p.late_wait && blockIdx.x == 0 && blockIdx.y == 0) pdl_wait();      // see AttnFwdArgs.late_wait
}

template <typename ST, int KB>
inline int launch_attn_context_bulk(const AttnFwdArgs& p, int images, cudaStream_t st);   // attention_bulk.cuh
template <typename ST, int KB>
inline int launch_attn_context_mma_st(const AttnFwdArgs& p, int images, cudaStream_t st);  // attention_mma.cuh
inline int bulk_ctx_mode() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DIC_BULK_CTX"); v = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1; }
  return v;
}
inline bool bulk_ctx_enabled() { return bulk_ctx_mode() != 0; }

template <typename ST, int KB>
inline int launch_attn_step_kb(const AttnFwdArgs& p, int images, cudaStream_t st) {
  static DeviceOnce attr_set;
  if (int dev_ = 0; attr_set.need(&dev_)) {
    DIC_CUDA(cudaFuncSetAttribute(attn_alpha_kernel<ST, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    DIC_CUDA(cudaFuncSetAttribute(attn_context_kernel<ST, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set.mark(dev_);
  }
  if (!p.skip_alpha) {
    ProfScope prof(P_ATTN_ALPHA, st, (double)images * p.L * p.A * sizeof(ST));
    AttnFwdArgs pa = p;
    pa.trace = g_trace_host;
    static int per_image = -1;
    if (per_image < 0) { const char* e = getenv("DIC_ALPHA_PER_IMAGE"); per_image = (e && e[0] == '0') ? 0 : 1; }
    if (KB > 1 && per_image) {
      // one CTA per IMAGE computing its KB rows from one read of the att1 slab
      static DeviceOnce attr2;
      if (int dev_ = 0; attr2.need(&dev_)) {
        DIC_CUDA(cudaFuncSetAttribute(attn_alpha_kernel<ST, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr2.mark(dev_);
      }
      pa.rpi = 0;
      DIC_CUDA(launch_pdl(attn_alpha_kernel<ST, KB>, dim3(images), dim3(kAlphaThreads),
                          attn_alpha_smem_bytes(p.L, p.A, KB), st, pa));
    } else {
      // one CTA per ROW (beam): images*KB CTAs, the KB CTAs of an image share its att1 slab through L2
      pa.rpi = KB;
      DIC_CUDA(launch_pdl(attn_alpha_kernel<ST, 1>, dim3(images * KB), dim3(kAlphaThreads),
                          attn_alpha_smem_bytes(p.L, p.A, 1), st, pa));
    }
    DIC_LAUNCH_CHECK();
  }
  if (KB >= 3 && p.mode == DIC_ATTN_SOFT && bulk_ctx_enabled()) {
    // several beams per image: tensor-core variant for bf16 storage (attention_mma.cuh), TMA-staged FP32
    // variant otherwise (attention_bulk.cuh).  DIC_BULK_CTX=0 falls back to the register-streaming kernel,
    // DIC_BULK_CTX=2 forces the FP32 staged variant.
    // measured at 128 images (profiles/r01_beam_ctx_variants.txt): 3 beams 19.7 us staged FP32 vs 21.9 us MMA,
    // 5 beams 24.9 vs 24.1, 8 beams 31.5 vs 24.0 (the MMA variant does not depend on the beam count)
    if (sizeof(ST) == 2 && KB >= 4 && p.D % 8 == 0 && !p.z_out && bulk_ctx_mode() == 1)
      return launch_attn_context_mma_st<ST, KB>(p, images, st);
    return launch_attn_context_bulk<ST, KB>(p, images, st);
  }
  {
    // algorithmic bytes of the context pass: the annotations once per image-step
    ProfScope prof(P_ATTN_FWD, st, (double)images * p.L * (double)p.D * sizeof(ST));
    dim3 grid(cdiv(p.D, 32 * ctx_cpt(KB)), images);
    AttnFwdArgs pc = p;
    pc.trace = g_trace_host;
    DIC_CUDA(launch_pdl(attn_context_kernel<ST, KB>, grid, dim3(kCtxThreads), attn_ctx_smem_bytes(p.L, KB), st, pc));
    DIC_LAUNCH_CHECK();
  }
  return 0;
}

template <typename ST>
inline int launch_attn_step(const AttnFwdArgs& p, int images, int KB, cudaStream_t st) {
  if (images <= 0) return 0;
  if (!p.alpha_out) DIC_FAIL(-4, "attn_step: alpha buffer is required");
  switch (KB) {
    case 1: return launch_attn_step_kb<ST, 1>(p, images, st);
    case 2: return launch_attn_step_kb<ST, 2>(p, images, st);
    case 3: return launch_attn_step_kb<ST, 3>(p, images, st);
    case 4: return launch_attn_step_kb<ST, 4>(p, images, st);
    case 5: return launch_attn_step_kb<ST, 5>(p, images, st);
    case 6: return launch_attn_step_kb<ST, 6>(p, images, st);
    case 7: return launch_attn_step_kb<ST, 7>(p, images, st);
    case 8: return launch_attn_step_kb<ST, 8>(p, images, st);
    default: DIC_FAIL(-4, "attn_step: rows per image %d not in 1..8", KB);
  }
}

// ------------------------------------------------------------------------------------------
// Backward of one attention step (training, one row per image), again split into the streaming
// pass over the annotations and a small per-image kernel:
//   (a) attn_bwd_stream_kernel, grid (D/256, rows):
//         dz      = dzg * beta                -> DZ (ST), consumed post-loop by dF += alpha^T dz
//         dbeta'  = dzg * z * beta (1-beta)   -> G[:, gcol_beta + d]     (ST)
//         dalpha partial[chunk][b][l] = F[l, chunk] . dz[chunk]
//   (b) attn_bwd_small_kernel, one CTA per image:
//         dalpha  = sum_chunks partial (+ external d_alphas)
//         de      = alpha * (dalpha - sum alpha dalpha) * inv_temp       -> de_out
//         datt2   = w * sum_l de[l] 1[pre>0]  -> G[:, gcol_att2 + a]     (ST)
//         dw_full partial = sum_l de[l] relu(pre[l,:]) ; db_full partial = sum_l de[l]
// ------------------------------------------------------------------------------------------
struct AttnBwdArgs {
  const void* F;         // [B, L, D] ST
  const void* att1;      // [B, L, A] ST
  const float* hp;       // [B, A+D] fp32 att2 | beta   (this step)
  const float* z;        // [B, D] fp32                 (this step)
  float* dzg;            // [B, D] fp32; dzg_rezero: cleared again after it is read (the next step's split-K GEMM
                         // accumulates into it with red.add, two addends: order independent)
  int dzg_rezero;
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
  float* dal_part;       // [chunks][part_rows][L] fp32 scratch
  int part_rows;
  int L, D, A;
  float inv_temp;
  TraceRec* trace;
};

template <typename ST>
__global__ void __launch_bounds__(kCtxThreads) attn_bwd_stream_kernel(const AttnBwdArgs p) {
  const int L = p.L, D = p.D, A = p.A;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, rg = tid >> 5;
  const int d = chunk * kCtxCols + lane * 8;
  const bool active = d < D;
  const ST* F = reinterpret_cast<const ST*>(p.F) + (size_t)b * L * D + d;
  float* part = p.dal_part + ((size_t)chunk * p.part_rows + b) * L;
  Trace trace(p.trace);

  // `pre` is warp-uniform (rg is the warp index): the blocks below contain warp shuffles
  Raw8<ST> v0[kCtxUnroll];
  const bool pre = (rg + (kCtxUnroll - 1) * kCtxGroups < L);
  if (pre) {
#pragma unroll
    for (int r = 0; r < kCtxUnroll; ++r) {
      if (active) v0[r].load_stream(F + (size_t)(rg + r * kCtxGroups) * D);
      else v0[r].zero();
    }
  }
  // beta and z are forward state (static here): loaded before the wait as well
  float beta[8], zz[8];
  if (active) {
    load8<float>(p.hp + (size_t)b * (A + D) + A + d, beta);
    load8<float>(p.z + (size_t)b * D + d, zz);
  }
  pdl_wait();       // dzg comes from the preceding GEMM; the annotation loads above are static
  pdl_trigger();
  trace.mark();

  float dz[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) dz[q] = 0.f;
  if (active) {
    float g[8];
    load8<float>(p.dzg + (size_t)b * D + d, g);

    float db[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      dz[q] = g[q] * beta[q];
      db[q] = g[q] * zz[q] * beta[q] * (1.f - beta[q]);
    }
    if (rg == 0) {
      store8<ST>(reinterpret_cast<ST*>(p.DZ) + (size_t)b * D + d, dz);
      store8<ST>(reinterpret_cast<ST*>(p.G) + (size_t)b * p.g_stride + p.gcol_beta + d, db);
    }
  }
  if (p.dzg_rezero) {
    __syncthreads();          // every warp of the CTA has read its dzg values
    if (active && rg == 0) {
      const float zero8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      store8<float>(p.dzg + (size_t)b * D + d, zero8);
    }
  }

  int l = rg;
  if (pre) {
#pragma unroll
    for (int r = 0; r < kCtxUnroll; ++r) {
      float v[8];
      v0[r].unpack(v);
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) s = fmaf(v[q], dz[q], s);
      s = warp_sum(s);
      if (lane == 0) part[l + r * kCtxGroups] = s;
    }
    l += kCtxUnroll * kCtxGroups;
  }
  for (; l + (kCtxUnroll - 1) * kCtxGroups < L; l += kCtxUnroll * kCtxGroups) {
    Raw8<ST> raw[kCtxUnroll];
#pragma unroll
    for (int r = 0; r < kCtxUnroll; ++r) {
      if (active) raw[r].load_stream(F + (size_t)(l + r * kCtxGroups) * D);
      else raw[r].zero();
    }
#pragma unroll
    for (int r = 0; r < kCtxUnroll; ++r) {
      float v[8];
      raw[r].unpack(v);
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) s = fmaf(v[q], dz[q], s);
      s = warp_sum(s);
      if (lane == 0) part[l + r * kCtxGroups] = s;
    }
  }
  for (; l < L; l += kCtxGroups) {
    float s = 0.f;
    if (active) {
      float v[8];
      load8_stream<ST>(F + (size_t)l * D, v);
#pragma unroll
      for (int q = 0; q < 8; ++q) s = fmaf(v[q], dz[q], s);
    }
    s = warp_sum(s);
    if (lane == 0) part[l] = s;
  }
  trace.end(TK_BWD_STREAM);
}

constexpr int kBwdSmallThreads = 256;     // (512 threads x 115 registers = one CTA per SM, two waves: 12.6 us instead of 7.5)
constexpr int kBwdRowGroups = 16;

inline size_t attn_bwd_small_smem_bytes(int L, int A) {
  // de[L] | al[L] | w[A] | att2[A] | red[2][16][A] | scratch[64]
  return sizeof(float) * (2 * (size_t)L + 2 * (size_t)A + 2 * kBwdRowGroups * (size_t)A + 64);
}

template <typename ST>
__global__ void __launch_bounds__(kBwdSmallThreads, 2) attn_bwd_small_kernel(const AttnBwdArgs p, int chunks) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int L = p.L, D = p.D, A = p.A;
  float* de_s = reinterpret_cast<float*>(smem_raw);
  float* al_s = de_s + L;
  float* w_s = al_s + L;
  float* att2_s = w_s + A;
  float* red_s = att2_s + A;                          // [2][kBwdRowGroups][A]
  float* scratch = red_s + 2 * kBwdRowGroups * A;     // 64
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const float* hp = p.hp + (size_t)b * (A + D);
  ST* G = reinterpret_cast<ST*>(p.G) + (size_t)b * p.g_stride;
  Trace trace(p.trace);
  // att1 is loop invariant: its first block of loads (the whole slab at the reference shape) is issued
  // before the dependency wait and stays in flight through the softmax-backward reductions below
  constexpr int IT = 13;     // 13 x 16 row groups >= 196 rows
  const ST* att1 = reinterpret_cast<const ST*>(p.att1) + (size_t)b * L * A;
  Raw8<ST> raw0[IT];
  if (A == 128) {
    const int cg = tid & 15, rgp = tid >> 4;
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      const int l = it * kBwdRowGroups + rgp;
      if (l < L) raw0[it].load_stream(att1 + (size_t)l * A + cg * 8);
      else raw0[it].zero();
    }
  }
  pdl_wait();
  pdl_trigger();
  trace.mark();

  for (int a = tid; a < A; a += kBwdSmallThreads) {
    w_s[a] = p.w_full[a];
    att2_s[a] = hp[a];
  }
  float part = 0.f;
  for (int l = tid; l < L; l += kBwdSmallThreads) {
    float s = 0.f;
    for (int c = 0; c < chunks; ++c) s += p.dal_part[((size_t)c * p.part_rows + b) * L + l];
    if (p.dalpha) s += p.dalpha[(size_t)b * p.alpha_stride + l];
    const float al = p.alpha[(size_t)b * p.alpha_stride + l];
    al_s[l] = al;
    de_s[l] = s;         // dalpha for now
    part += al * s;
  }
  const float dot = block_sum(part, scratch);
  float desum = 0.f;
  for (int l = tid; l < L; l += kBwdSmallThreads) {
    const float de = al_s[l] * (de_s[l] - dot) * p.inv_temp;
    de_s[l] = de;
    p.de_out[(size_t)b * L + l] = de;
    desum += de;
  }
  const float dbf = block_sum(desum, scratch);   // its barriers also publish de_s
  if (tid == 0) p.dbfull_part[b] = dbf;

  // relu-mask pass over att1.  red_s holds [2][kBwdRowGroups][A] partial sums.
  if (A == 128) {
    // reference shape: 16 column groups (8 columns, one 16-byte load) x 16 row groups
    const int cg = tid & 15, rgp = tid >> 4;
    float a2[8], s1[8], s2[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { a2[q] = att2_s[cg * 8 + q]; s1[q] = 0.f; s2[q] = 0.f; }
    auto consume = [&](const Raw8<ST> (&raw)[IT], int l0) {
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        const int l = l0 + it * kBwdRowGroups + rgp;
        if (l < L) {
          float v[8];
          raw[it].unpack(v);
          const float de = de_s[l];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float pre = v[q] + a2[q];
            if (pre > 0.f) { s1[q] += de; s2[q] = fmaf(de, pre, s2[q]); }
          }
        }
      }
    };
    consume(raw0, 0);                                        // the block prefetched before the dependency wait
    for (int l0 = IT * kBwdRowGroups; l0 < L; l0 += IT * kBwdRowGroups) {     // L > 208 only
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        const int l = l0 + it * kBwdRowGroups + rgp;
        if (l < L) raw0[it].load_stream(att1 + (size_t)l * A + cg * 8);
        else raw0[it].zero();
      }
      consume(raw0, l0);
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      red_s[(0 * kBwdRowGroups + rgp) * A + cg * 8 + q] = s1[q];
      red_s[(1 * kBwdRowGroups + rgp) * A + cg * 8 + q] = s2[q];
    }
  } else {
    for (int i = tid; i < 2 * kBwdRowGroups * A; i += kBwdSmallThreads) red_s[i] = 0.f;
    __syncthreads();
    const int rg = tid >> 7, a0 = tid & 127;   // 2 row groups x 128 columns
    for (int ab = 0; ab < A; ab += 128) {
      const int a = ab + a0;
      if (a < A) {
        const float a2 = att2_s[a];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
        for (int l = rg; l < L; l += kBwdSmallThreads / 128) {
          const float pre = to_f<ST>(att1[(size_t)l * A + a]) + a2;
          if (pre > 0.f) {
            const float de = de_s[l];
            s1 += de;
            s2 = fmaf(de, pre, s2);
          }
        }
        red_s[(0 * kBwdRowGroups + rg) * A + a] = s1;
        red_s[(1 * kBwdRowGroups + rg) * A + a] = s2;
      }
    }
  }
  __syncthreads();
  for (int a = tid; a < A; a += kBwdSmallThreads) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int g = 0; g < kBwdRowGroups; ++g) {
      s1 += red_s[(0 * kBwdRowGroups + g) * A + a];
      s2 += red_s[(1 * kBwdRowGroups + g) * A + a];
    }
    G[p.gcol_att2 + a] = from_f<ST>(w_s[a] * s1);
    p.dwfull_part[(size_t)b * A + a] = s2;
  }
  trace.end(TK_BWD_SMALL);
}

template <typename ST>
inline int launch_attn_bwd(const AttnBwdArgs& p_in, int rows, cudaStream_t st) {
  if (rows <= 0) return 0;
  AttnBwdArgs p = p_in;
  p.trace = g_trace_host;
  const int chunks = cdiv(p.D, kCtxCols);
  {
    ProfScope prof(P_ATTN_BWD, st, (double)rows * p.L * (double)p.D * sizeof(ST));
    dim3 grid(chunks, rows);
    DIC_CUDA(launch_pdl(attn_bwd_stream_kernel<ST>, grid, dim3(kCtxThreads), 0, st, p));
    DIC_LAUNCH_CHECK();
  }
  {
    static DeviceOnce attr_set;
    if (int dev_ = 0; attr_set.need(&dev_)) {
      DIC_CUDA(cudaFuncSetAttribute(attn_bwd_small_kernel<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      attr_set.mark(dev_);
    }
    ProfScope prof(P_ATTN_BWD_SMALL, st, (double)rows * p.L * p.A * sizeof(ST));
    DIC_CUDA(launch_pdl(attn_bwd_small_kernel<ST>, dim3(rows), dim3(kBwdSmallThreads),
                        attn_bwd_small_smem_bytes(p.L, p.A), st, p, chunks));
    DIC_LAUNCH_CHECK();
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// Post-loop: datt1[b,l,a] = w[a] * sum_t de_t[b,l] 1[att1[b,l,a] + att2_t[b,a] > 0]
// (the relu output is recomputed from att1 + att2 instead of being saved per step: saving it
// would cost B*L*A*4 bytes per step, SURVEY.md section 7 hard part 9).
// Grid (L chunks of 32 rows, B).
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

// CTA = 32 annotation rows x 128 columns of one image, 128 threads; a thread owns 8 rows x 4 columns.
// att2_t and de_t tiles are staged in shared memory (32 steps at a time), so the inner loop is three
// 16-byte shared loads per step for 32 mask-and-add updates.  (The first version read de_t[l] and
// att2_t[a] from global memory per update and took 107 us for 13 MB of traffic.)
constexpr int kDatt1TT = 32;
template <typename ST>
__global__ void __launch_bounds__(128) datt1_kernel(const Datt1Args p) {
  pdl_wait();          // launched programmatically (launch_pdl): nothing of a predecessor is touched before this
  pdl_trigger();
  __shared__ __align__(16) float a2_s[kDatt1TT][128];
  __shared__ __align__(16) float de_s[kDatt1TT][32];
  const int b = blockIdx.y;
  const int l0 = blockIdx.x * 32;
  const int tid = threadIdx.x, cg = tid & 31, rq = tid >> 5;
  int Tb = 0;
  for (int t = 0; t < p.T; ++t) Tb += (p.sizes.n[t] > b) ? 1 : 0;
  const ST* att1 = reinterpret_cast<const ST*>(p.att1) + (size_t)b * p.L * p.A;
  ST* out = reinterpret_cast<ST*>(p.datt1) + (size_t)b * p.L * p.A;
  for (int ab = 0; ab < p.A; ab += 128) {
    const int a = ab + cg * 4;
    const bool col_ok = a < p.A;            // A % 4 == 0: a thread's 4 columns are in or out together
    float v[8][4], acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int l = l0 + rq * 8 + r;
#pragma unroll
      for (int q = 0; q < 4; ++q) { v[r][q] = 0.f; acc[r][q] = 0.f; }
      if (col_ok && l < p.L) load4<ST>(att1 + (size_t)l * p.A + a, v[r]);
    }
    for (int t0 = 0; t0 < Tb; t0 += kDatt1TT) {
      const int tn = min(kDatt1TT, Tb - t0);
      __syncthreads();
      for (int i = tid; i < tn * 128; i += 128) {
        const int t = i >> 7, c = i & 127;
        a2_s[t][c] = (ab + c < p.A) ? p.hp_all[((size_t)(t0 + t) * p.B + b) * (p.A + p.D) + ab + c] : 0.f;
      }
      for (int i = tid; i < tn * 32; i += 128) {
        const int t = i >> 5, r = i & 31;
        de_s[t][r] = (l0 + r < p.L) ? p.de_all[((size_t)(t0 + t) * p.B + b) * p.L + l0 + r] : 0.f;
      }
      __syncthreads();
      for (int t = 0; t < tn; ++t) {
        const float4 a2 = *reinterpret_cast<const float4*>(&a2_s[t][cg * 4]);
        const float4 d0 = *reinterpret_cast<const float4*>(&de_s[t][rq * 8]);
        const float4 d1 = *reinterpret_cast<const float4*>(&de_s[t][rq * 8 + 4]);
        const float de[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        // att1 + att2 > 0  <=>  att1 > -att2 exactly (a floating-point sum keeps the sign of the exact sum): four
        // negations per step instead of 32 additions -- the loop is instruction bound (compare + predicated add left)
        const float na2[4] = {-a2.x, -a2.y, -a2.z, -a2.w};
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (v[r][q] > na2[q]) acc[r][q] += de[r];
      }
    }
    if (col_ok) {
      float w[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) w[q] = p.w_full[a + q];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int l = l0 + rq * 8 + r;
        if (l < p.L) {
          float o[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) o[q] = w[q] * acc[r][q];
          store4<ST>(out + (size_t)l * p.A + a, o);
        }
      }
    }
  }
}

template <typename ST>
inline int launch_datt1(const Datt1Args& p, cudaStream_t st) {
  dim3 grid(cdiv(p.L, 32), p.B);
  ProfScope prof(P_DATT1, st, (double)p.B * p.L * p.A * sizeof(ST) * 2);
  DIC_CUDA(launch_pdl(datt1_kernel<ST>, grid, dim3(128), 0, st, p));
  DIC_LAUNCH_CHECK();
  return 0;
}

}  // namespace dic
