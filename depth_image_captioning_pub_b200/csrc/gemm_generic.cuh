// CUDA-core FMA GEMM with fully generic operand strides.
//
//   C[z][m,n] = act( sum_k A[z][m,k] * B[z][n,k]  (+ bias[z][n]) )  (+= when accumulate)
//
// A(m,k) = A[m*a_m + k*a_k], B(n,k) = B[n*b_n + k*b_k]: any transposition is a stride choice.
// Operands may be fp32 or bf16 (runtime flags), accumulation is always fp32 FMA, so this
// is the engine of the DIC_F32 parity mode and of every small / oddly shaped contraction.
// The large bf16 contractions go to the tcgen05 engine (gemm_tc.cuh).
//
// Split-K: gridDim.z = batch * splits.  split_mode 0 = atomicAdd into fp32 C (C pre-zeroed or
// accumulating), 1 = write partial s to C + s*split_stride (consumer reduces).
#pragma once
#include "common.cuh"

namespace dic {

struct GemmArgs {
  const void* A;
  const void* B;
  void* C;
  const float* bias;      // per-column, may be null
  int M, N, K;
  long long a_m, a_k, b_n, b_k;
  long long ldc;          // C row stride (elements)
  long long a_batch, b_batch, c_batch, bias_batch;  // element strides between batches
  int batch;              // >= 1
  int splits;             // >= 1
  int split_mode;         // 0 atomic, 1 partial buffers
  long long split_stride; // elements between partial buffers
  int a_bf16, b_bf16, c_bf16;
  int accumulate;         // C += (fp32 C only)
  float alpha;            // scales the product (not the bias)
  float bias_scale;
  int sig_lo, sig_hi;     // apply sigmoid to columns [sig_lo, sig_hi)
  int tag;                // call-site id for the timeline trace
  int bn;                 // tcgen05 engine: N tile width (32/64/128), 0 = chosen from the shape
  int fast_act;           // sigmoid through ex2.approx / rcp.approx (bf16 mode)
  int b_static;           // B is a packed weight written at least two launches ago: the tcgen05 engine may
                          // fetch its first tiles before the programmatic-dependency wait
  int prof;               // event-profile class of this launch (0 = the engine's default class)
  double prof_bytes;      // algorithmic bytes attributed to it
};

inline GemmArgs gemm_args_nt(const void* A, int a_bf16, long long lda, const void* B, int b_bf16,
                             long long ldb, void* C, int c_bf16, long long ldc, int M, int N,
                             int K, const float* bias) {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.A = A; g.B = B; g.C = C; g.bias = bias;
  g.M = M; g.N = N; g.K = K;
  g.a_m = lda; g.a_k = 1; g.b_n = ldb; g.b_k = 1; g.ldc = ldc;
  g.batch = 1; g.splits = 1;
  g.a_bf16 = a_bf16; g.b_bf16 = b_bf16; g.c_bf16 = c_bf16;
  g.alpha = 1.f; g.bias_scale = 1.f;
  return g;
}

template <int BM, int BN, int BK>
__global__ void __launch_bounds__(256) gemm_generic_kernel(const GemmArgs g) {
  constexpr int TM = BM / 16, TN = BN / 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  pdl_wait();
  pdl_trigger();
  const int z = blockIdx.z;
  const int batch = z / g.splits, split = z % g.splits;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  // K range of this split, in multiples of BK
  const int ktiles = (g.K + BK - 1) / BK;
  const int per = (ktiles + g.splits - 1) / g.splits;
  const int kt0 = split * per, kt1 = min(ktiles, kt0 + per);

  const char* Ab = reinterpret_cast<const char*>(g.A) + (size_t)batch * g.a_batch * (g.a_bf16 ? 2 : 4);
  const char* Bb = reinterpret_cast<const char*>(g.B) + (size_t)batch * g.b_batch * (g.b_bf16 ? 2 : 4);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // loader mapping: fastest-varying thread index follows the contiguous operand dimension
  const bool a_kfast = (g.a_k == 1);
  const bool b_kfast = (g.b_k == 1);

  for (int kt = kt0; kt < kt1; ++kt) {
    const int k0 = kt * BK;
#pragma unroll 4
    for (int i = tid; i < BM * BK; i += 256) {
      int mm, kk;
      if (a_kfast) { kk = i % BK; mm = i / BK; } else { mm = i % BM; kk = i / BM; }
      const int m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < g.M && k < g.K) v = ld_as_float(Ab, (size_t)m * g.a_m + (size_t)k * g.a_k, g.a_bf16);
      As[kk][mm] = v;
    }
#pragma unroll 4
    for (int i = tid; i < BN * BK; i += 256) {
      int nn, kk;
      if (b_kfast) { kk = i % BK; nn = i / BK; } else { nn = i % BN; kk = i / BN; }
      const int n = n0 + nn, k = k0 + kk;
      float v = 0.f;
      if (n < g.N && k < g.K) v = ld_as_float(Bb, (size_t)n * g.b_n + (size_t)k * g.b_k, g.b_bf16);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        float4 t = *reinterpret_cast<const float4*>(&As[kk][ty * 4 + i * 16]);
        a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        float4 t = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4 + j * 16]);
        b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // epilogue.  Row of acc[i][*]: m0 + ty*4 + (i/4)*64 + i%4 ; column: n0 + tx*4 + (j/4)*64 + j%4
  const bool first = (split == 0);
  const float* bias = g.bias ? g.bias + (size_t)batch * g.bias_batch : nullptr;
  char* Cb = reinterpret_cast<char*>(g.C);
  size_t cbase = (size_t)batch * g.c_batch;
  if (g.splits > 1 && g.split_mode == 1) cbase += (size_t)split * g.split_stride;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * 4 + (i >> 2) * 64 + (i & 3);
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * 4 + (j >> 2) * 64 + (j & 3);
      if (n >= g.N) continue;
      float v = acc[i][j] * g.alpha;
      if (bias && first) v += g.bias_scale * bias[n];
      const size_t ci = cbase + (size_t)m * g.ldc + n;
      if (g.splits > 1 && g.split_mode == 0) {
        atomicAdd(reinterpret_cast<float*>(Cb) + ci, v);
      } else {
        if (g.accumulate) v += reinterpret_cast<float*>(Cb)[ci];
        if (n >= g.sig_lo && n < g.sig_hi) v = sigmoidf_acc(v);
        st_from_float(Cb, ci, v, g.c_bf16);
      }
    }
  }
}

// Host launcher.  Picks the 128x128 tile for large outputs, 64x64 otherwise.
inline int gemm_generic(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return 0;
  if (g.splits > 1 && g.split_mode == 0 && (g.c_bf16 || g.sig_hi > g.sig_lo))
    DIC_FAIL(-4, "gemm_generic: atomic split-K needs fp32 C and no activation");
  if (g.accumulate && g.c_bf16) DIC_FAIL(-4, "gemm_generic: accumulate needs fp32 C");
  ProfScope prof(P_GEMM_FMA, st);
  const long long tiles128 = (long long)cdiv(g.M, 128) * cdiv(g.N, 128) * g.batch * g.splits;
  if (tiles128 >= 148 && g.M >= 128 && g.N >= 128) {
    dim3 grid(cdiv(g.N, 128), cdiv(g.M, 128), g.batch * g.splits);
    DIC_CUDA(launch_pdl(gemm_generic_kernel<128, 128, 16>, grid, dim3(256), 0, st, g));
  } else {
    dim3 grid(cdiv(g.N, 64), cdiv(g.M, 64), g.batch * g.splits);
    DIC_CUDA(launch_pdl(gemm_generic_kernel<64, 64, 32>, grid, dim3(256), 0, st, g));
  }
  DIC_LAUNCH_CHECK();
  return 0;
}

// choose a split count so that the grid covers the SMs ~2x; K tiles of 32
inline int pick_splits(int M, int N, int K, int batch = 1) {
  long long tiles = (long long)cdiv(M, 64) * cdiv(N, 64) * batch;
  int ktiles = cdiv(K, 32);
  int s = (int)((2 * 148 + tiles - 1) / tiles);
  if (s > ktiles) s = ktiles;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  return s;
}

}  // namespace dic
