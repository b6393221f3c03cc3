// Depth CNN encoder (reference Depth_CNN_endoder, depth_models.py:12-56), forward and backward:
//
//   conv1 1->128 k7 s3 | BN | ReLU | maxpool 3 | conv2 128->512 k3 | BN | ReLU | maxpool 3 | conv3 512->2048 k1 |
//   BN | ReLU | AdaptiveAvgPool2d(14) | permute -> [B, 196, 2048]
//
// With the reference's 224 x 224 depth maps the last feature map is 7 x 7, so AdaptiveAvgPool2d(14) is an exact
// 2 x 2 replication (output (i, j) averages the single input (i/2, j/2)): the annotation rows the decoder reads
// are written directly in its [B, L, D] layout and dtype, and the backward sums the four replicas of dL/dF.
//
// Everything is channels-last.  The three convolutions are GEMMs on the library's engines (tcgen05 in bf16 storage,
// FMA in the fp32 parity mode): im2col rows [B*Ho*Wo, k*k*Ci] (K index = tap * Ci + c, conv1 padded 49 -> 64)
// against re-packed weights [Co, k*k*Ci]; weight gradients are the transposed contractions over the same
// buffers, the data gradient of conv2 is a GEMM + col2im gather, conv3 is 1 x 1 (no im2col).  Batch-norm uses
// batch statistics in training (fp64 accumulation of per-block partial sums, running statistics updated with the
// module's momentum) and the running statistics in eval; BN + ReLU + pooling are one kernel per stage, and their
// backward is one routing/reduction kernel (dy through max-pool arg-max and the ReLU mask, per-channel sum(dy) and
// sum(dy * xhat)) plus one elementwise kernel (dx = gamma * invstd * (dy - mean(dy) - xhat * mean(dy * xhat))).
#pragma once
#include "common.cuh"
#include "layout.cuh"

namespace dic {

struct EncGeom {
  int B, Hi, Wi;          // input depth maps [B, Hi, Wi]
  int H1, W1, P1h, P1w;   // conv1 output, pooled
  int H2, W2, P2h, P2w;   // conv2 output, pooled (= conv3 output)
  int C1, C2, C3;         // 128, 512, 2048
  int K1p;                // padded K of conv1 (64)
  EncGeom(int B_, int Hi_, int Wi_) {
    B = B_; Hi = Hi_; Wi = Wi_;
    H1 = (Hi - 7) / 3 + 1; W1 = (Wi - 7) / 3 + 1;
    P1h = H1 / 3; P1w = W1 / 3;
    H2 = P1h - 2; W2 = P1w - 2;
    P2h = H2 / 3; P2w = W2 / 3;
    C1 = 128; C2 = 512; C3 = 2048; K1p = 64;
  }
  size_t M1() const { return (size_t)B * H1 * W1; }
  size_t M2() const { return (size_t)B * H2 * W2; }
  size_t M3() const { return (size_t)B * P2h * P2w; }
};

// workspace carve-up (ST = storage type of the mode)
struct EncLayout {
  size_t col1, out1, p1, col2, out2, p2, out3;           // forward (kept for backward)
  size_t w1p, w2p, w3p;                                   // re-packed weights (ST)
  size_t stats;                                           // double [3][2][2048]: sum, sumsq | backward: sum dy, sum dy*xhat
  size_t mean, invstd;                                    // float [3][2048]
  size_t dy3, dy2, dy1, dx16, dp2, dcol2, dp1, dwp;       // backward scratch
  size_t bytes;
  EncLayout(const EncGeom& g, int dtype) {
    const size_t es = dtype == DIC_BF16 ? 2 : 4;
    Carver c;
    col1 = c.take(g.M1() * g.K1p * es);
    out1 = c.take(g.M1() * g.C1 * 4);
    p1 = c.take((size_t)g.B * g.P1h * g.P1w * g.C1 * es);
    col2 = c.take(g.M2() * 9 * g.C1 * es);
    out2 = c.take(g.M2() * g.C2 * 4);
    p2 = c.take(g.M3() * g.C2 * es);
    out3 = c.take(g.M3() * g.C3 * 4);
    w1p = c.take((size_t)g.C1 * g.K1p * es);
    w2p = c.take((size_t)g.C2 * 9 * g.C1 * es);
    w3p = c.take((size_t)g.C3 * g.C2 * es);
    stats = c.take(sizeof(double) * 3 * 2 * 2048);
    mean = c.take(sizeof(float) * 3 * 2048);
    invstd = c.take(sizeof(float) * 3 * 2048);
    dy3 = c.take(g.M3() * g.C3 * 4);
    dy2 = c.take(g.M2() * g.C2 * 4);
    dy1 = c.take(g.M1() * g.C1 * 4);
    dx16 = c.take(g.M1() * g.C1 * 2);                     // bf16 copy of a stage's dx (GEMM operand), largest stage
    dp2 = c.take(g.M3() * g.C2 * 4);
    dcol2 = c.take(g.M2() * 9 * g.C1 * 4);
    dp1 = c.take((size_t)g.B * g.P1h * g.P1w * g.C1 * 4);
    dwp = c.take(sizeof(float) * (size_t)g.C2 * 9 * g.C1);   // packed weight gradient (largest: conv2)
    bytes = c.off;
  }
};

// ---- im2col: src [B, Hi, Wi, Ci] (channels-last) -> col [B*Ho*Wo, Kp], K index = (ky*k + kx)*Ci + c ----------
template <typename SRC, typename ST>
__global__ void __launch_bounds__(256) enc_im2col_kernel(const SRC* __restrict__ src, ST* __restrict__ col, int B, int Hi,
                                                         int Wi, int Ci, int k, int stride, int Ho, int Wo, int Kp) {
  const size_t total = (size_t)B * Ho * Wo * Kp;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int kk = (int)(i % Kp);
    const size_t row = i / Kp;
    float v = 0.f;
    if (kk < k * k * Ci) {
      const int c = kk % Ci, tap = kk / Ci;
      const int ky = tap / k, kx = tap - ky * k;
      const int xo = (int)(row % Wo);
      const size_t r2 = row / Wo;
      const int yo = (int)(r2 % Ho), b = (int)(r2 / Ho);
      v = to_f<SRC>(src[(((size_t)b * Hi + yo * stride + ky) * Wi + xo * stride + kx) * Ci + c]);
    }
    col[i] = from_f<ST>(v);
  }
}

// ---- conv weight re-pack: w [Co, Ci, k, k] fp32 -> wp [Co, Kp] ST (K index = tap*Ci + c), and the inverse for grads
template <typename ST>
__global__ void __launch_bounds__(256) enc_pack_w_kernel(const float* __restrict__ w, ST* __restrict__ wp, int Co, int Ci,
                                                         int kk, int Kp) {
  const int total = Co * Kp;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    const int q = i % Kp, n = i / Kp;
    float v = 0.f;
    if (q < kk * Ci) { const int c = q % Ci, tap = q / Ci; v = w[((size_t)n * Ci + c) * kk + tap]; }
    wp[i] = from_f<ST>(v);
  }
}
__global__ void __launch_bounds__(256) enc_unpack_dw_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int Co,
                                                            int Ci, int kk, int Kp) {
  const int total = Co * Ci * kk;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += gridDim.x * 256) {
    const int tap = i % kk, c = (i / kk) % Ci, n = i / (kk * Ci);
    dw[i] = dwp[(size_t)n * Kp + tap * Ci + c];
  }
}

// ---- per-channel sum / sum of squares of X [M, C] fp32 (fp64 accumulation across blocks) ------------------------
__global__ void __launch_bounds__(256) enc_stats_kernel(const float* __restrict__ X, size_t M, int C, int rows_per_block,
                                                        double* __restrict__ sum, double* __restrict__ sumsq) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  const size_t r0 = (size_t)blockIdx.y * rows_per_block;
  const size_t r1 = r0 + rows_per_block < M ? r0 + rows_per_block : M;
  float s = 0.f, q = 0.f, cs = 0.f, cq = 0.f;     // Kahan-compensated partial sums of this block's rows
  for (size_t r = r0; r < r1; ++r) {
    const float v = X[r * C + c];
    float y = v - cs, t = s + y; cs = (t - s) - y; s = t;
    y = v * v - cq; t = q + y; cq = (t - q) - y; q = t;
  }
  atomicAdd(sum + c, (double)s);
  atomicAdd(sumsq + c, (double)q);
}

// mean / invstd of the batch (training) or of the running statistics (eval); training also updates the running
// statistics like nn.BatchNorm2d (momentum, unbiased variance)
__global__ void __launch_bounds__(256) enc_bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq,
                                                              double count, int C, int training, float momentum, float eps,
                                                              float* __restrict__ running_mean, float* __restrict__ running_var,
                                                              float* __restrict__ mean, float* __restrict__ invstd) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  if (training) {
    const double m = sum[c] / count;
    double var = sumsq[c] / count - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  } else {
    mean[c] = running_mean[c];
    invstd[c] = rsqrtf(running_var[c] + eps);
  }
}

// ---- BN + ReLU + max-pool P (P = 1: none) + replication R: X [B, H, W, C] fp32 -> Y [B, (H/P)*R, (W/P)*R, C] OT ----
template <typename OT>
__global__ void __launch_bounds__(256) enc_bn_relu_pool_kernel(const float* __restrict__ X, const float* __restrict__ mean,
                                                               const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, OT* __restrict__ Y, int B, int H,
                                                               int W, int C, int P, int R) {
  const int Hp = H / P, Wp = W / P;
  const size_t total = (size_t)B * Hp * Wp * C;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int c = (int)(i % C);
    size_t r = i / C;
    const int xp = (int)(r % Wp); r /= Wp;
    const int yp = (int)(r % Hp);
    const int b = (int)(r / Hp);
    const float sc = gamma[c] * invstd[c], sh = beta[c] - mean[c] * sc;
    float best = -INFINITY;
    for (int ky = 0; ky < P; ++ky)
      for (int kx = 0; kx < P; ++kx) {
        const float v = X[(((size_t)b * H + yp * P + ky) * W + xp * P + kx) * C + c];
        best = fmaxf(best, fmaxf(fmaf(v, sc, sh), 0.f));
      }
    const OT o = from_f<OT>(best);
    for (int ry = 0; ry < R; ++ry)
      for (int rx = 0; rx < R; ++rx)
        Y[(((size_t)b * Hp * R + yp * R + ry) * (Wp * R) + xp * R + rx) * C + c] = o;
  }
}

// ---- backward through replication R, max-pool P and ReLU: dY [B, (H/P)*R, (W/P)*R, C] (GT) -> dy [B, H, W, C] fp32 ----
// (dy must be zero where no pooling window reaches: the caller clears it when H % P or W % P), plus the two
// per-channel BN reductions sum(dy) and sum(dy * xhat) in fp64.
template <typename GT>
__global__ void __launch_bounds__(256) enc_bwd_route_kernel(const GT* __restrict__ dY, const float* __restrict__ X,
                                                            const float* __restrict__ mean, const float* __restrict__ invstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            float* __restrict__ dy, int B, int H, int W, int C, int P, int R,
                                                            int win_per_block, double* __restrict__ s_dy,
                                                            double* __restrict__ s_dyx) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= C) return;
  const int Hp = H / P, Wp = W / P;
  const size_t nwin = (size_t)B * Hp * Wp;
  const size_t w0 = (size_t)blockIdx.y * win_per_block;
  const size_t w1 = w0 + win_per_block < nwin ? w0 + win_per_block : nwin;
  const float m = mean[c], is = invstd[c];
  const float sc = gamma[c] * is, sh = beta[c] - m * sc;
  float a1 = 0.f, a2 = 0.f;
  for (size_t w = w0; w < w1; ++w) {
    const int xp = (int)(w % Wp);
    const size_t r = w / Wp;
    const int yp = (int)(r % Hp), b = (int)(r / Hp);
    float g = 0.f;
    for (int ry = 0; ry < R; ++ry)
      for (int rx = 0; rx < R; ++rx)
        g += to_f<GT>(dY[(((size_t)b * Hp * R + yp * R + ry) * (Wp * R) + xp * R + rx) * C + c]);
    // arg-max of relu(bn(x)) over the window in torch's scan order (first maximum wins)
    float best = -INFINITY, xbest = 0.f;
    int kbest = 0;
    for (int ky = 0; ky < P; ++ky)
      for (int kx = 0; kx < P; ++kx) {
        const float v = X[(((size_t)b * H + yp * P + ky) * W + xp * P + kx) * C + c];
        const float a = fmaxf(fmaf(v, sc, sh), 0.f);
        if (a > best) { best = a; kbest = ky * P + kx; xbest = v; }
      }
    const float d = best > 0.f ? g : 0.f;      // ReLU mask
    for (int ky = 0; ky < P; ++ky)
      for (int kx = 0; kx < P; ++kx)
        dy[(((size_t)b * H + yp * P + ky) * W + xp * P + kx) * C + c] = (ky * P + kx == kbest) ? d : 0.f;
    a1 += d;
    a2 = fmaf(d, (xbest - m) * is, a2);
  }
  atomicAdd(s_dy + c, (double)a1);
  atomicAdd(s_dyx + c, (double)a2);
}

// ---- BN backward, elementwise: dx = gamma*invstd*(dy - mean(dy) - xhat*mean(dy*xhat)) (training) -----------------
// in place on dy (fp32) and, optionally, a bf16 copy for the tensor-core GEMMs; also dgamma / dbeta
template <typename ST>
__global__ void __launch_bounds__(256) enc_bwd_apply_kernel(float* __restrict__ dy, const float* __restrict__ X,
                                                            const float* __restrict__ mean, const float* __restrict__ invstd,
                                                            const float* __restrict__ gamma, const double* __restrict__ s_dy,
                                                            const double* __restrict__ s_dyx, double count, size_t M, int C,
                                                            ST* __restrict__ dx_st, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta) {
  const size_t total = M * C;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int c = (int)(i % C);
    const float is = invstd[c];
    const float xh = (X[i] - mean[c]) * is;
    const float m1 = (float)(s_dy[c] / count), m2 = (float)(s_dyx[c] / count);
    const float dx = gamma[c] * is * (dy[i] - m1 - xh * m2);
    dy[i] = dx;
    if (dx_st) dx_st[i] = from_f<ST>(dx);
    if (i < (size_t)C) { dgamma[c] = (float)s_dyx[c]; dbeta[c] = (float)s_dy[c]; }
  }
}

// ---- col2im gather for the 3x3 stride-1 convolution: dcol [B*Ho*Wo, 9*Ci] fp32 -> dsrc [B, Hi, Wi, Ci] fp32 ------
__global__ void __launch_bounds__(256) enc_col2im3_kernel(const float* __restrict__ dcol, float* __restrict__ dsrc, int B,
                                                          int Hi, int Wi, int Ci, int Ho, int Wo) {
  const size_t total = (size_t)B * Hi * Wi * Ci;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (size_t)gridDim.x * 256) {
    const int c = (int)(i % Ci);
    size_t r = i / Ci;
    const int x = (int)(r % Wi); r /= Wi;
    const int y = (int)(r % Hi);
    const int b = (int)(r / Hi);
    float s = 0.f;
    for (int ky = 0; ky < 3; ++ky) {
      const int yo = y - ky;
      if (yo < 0 || yo >= Ho) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int xo = x - kx;
        if (xo < 0 || xo >= Wo) continue;
        s += dcol[(((size_t)b * Ho + yo) * Wo + xo) * (9 * Ci) + (ky * 3 + kx) * Ci + c];
      }
    }
    dsrc[i] = s;
  }
}

inline int enc_grid(size_t total) {
  size_t g = (total + 255) / 256;
  if (g > 148u * 16u) g = 148u * 16u;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace dic
