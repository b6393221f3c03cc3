// LSTM cell pointwise kernels (torch.nn.LSTMCell semantics, gate order i,f,g,o;
// depth_models.py:122,193).  The gate pre-activations come from the [emb|zg|h] x [W_ih|W_hh]^T
// GEMM as split-K partial tiles; this kernel reduces them, adds b_ih+b_hh, applies the
// nonlinearities and writes h' directly where the next step's GEMMs read it.
#pragma once
#include "common.cuh"

namespace dic {

struct LstmFwdArgs {
  const float* gate_part;  // [splits][rows_alloc][4H] fp32
  long long part_stride;   // elements between partials
  int splits;
  const float* bias_g;     // [4H] = b_ih + b_hh
  const float* c_in;       // [rows, H]
  float* c_out;            // [rows, H]
  float* acts;             // [rows, 4H] post-activation i,f,g,o (saved for backward) or null
  void* h_out;             // ST, row r at h_out + r*h_stride
  long long h_stride;
  void* hdrop_out;         // ST [rows, H] = h * mask (logits GEMM operand) or null
  const float* mask;       // [rows, H] or null
  float* h_f32_out;        // optional fp32 copy [rows, H] or null
  int rows, H;
  TraceRec* trace;
};

template <typename ST>
__global__ void __launch_bounds__(256) lstm_fwd_kernel(const LstmFwdArgs p) {
  Trace trace(p.trace);
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const bool live = idx < p.rows * p.H;
  const int H = p.H;
  const int r = live ? idx / H : 0, j = live ? idx - r * H : 0;
  // the previous cell state, the dropout mask and the bias are at least two launches old: loaded before the wait
  float c_prev = 0.f, mk = 1.f, bias4[4] = {0.f, 0.f, 0.f, 0.f};
  if (live) {
    c_prev = p.c_in[idx];
    if (p.hdrop_out && p.mask) mk = p.mask[idx];
#pragma unroll
    for (int q = 0; q < 4; ++q) bias4[q] = p.bias_g[q * H + j];
  }
  pdl_wait();
  pdl_trigger();
  trace.mark();
  if (!live) { trace.end(TK_LSTM_FWD); return; }
  // reduce the split-K partial tiles in a fixed order (deterministic); all loads are issued
  // before the adds so the up-to-64 L2 reads of a thread overlap
  float g4[4];
  constexpr int kMaxSplits = 16;
  float part[4][kMaxSplits];
#pragma unroll
  for (int sp = 0; sp < kMaxSplits; ++sp) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      part[q][sp] = sp < p.splits ? p.gate_part[(size_t)sp * p.part_stride + (size_t)r * 4 * H + q * H + j] : 0.f;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float s = bias4[q];
#pragma unroll
    for (int sp = 0; sp < kMaxSplits; ++sp) s += part[q][sp];
    g4[q] = s;
  }
  const float ig = sigmoidf_acc(g4[0]);
  const float fg = sigmoidf_acc(g4[1]);
  const float gg = tanhf(g4[2]);
  const float og = sigmoidf_acc(g4[3]);
  const float c = fg * c_prev + ig * gg;
  const float h = og * tanhf(c);
  p.c_out[idx] = c;
  if (p.acts) {
    float* a = p.acts + (size_t)r * 4 * H;
    a[j] = ig; a[H + j] = fg; a[2 * H + j] = gg; a[3 * H + j] = og;
  }
  reinterpret_cast<ST*>(p.h_out)[(size_t)r * p.h_stride + j] = from_f<ST>(h);
  if (p.h_f32_out) p.h_f32_out[idx] = h;
  if (p.hdrop_out) {
    reinterpret_cast<ST*>(p.hdrop_out)[idx] = from_f<ST>(h * mk);
  }
  trace.end(TK_LSTM_FWD);
}

template <typename ST>
inline int launch_lstm_fwd(const LstmFwdArgs& p_in, cudaStream_t st) {
  if (p_in.rows <= 0) return 0;
  LstmFwdArgs p = p_in;
  p.trace = g_trace_host;
  ProfScope prof(P_LSTM, st);
  DIC_CUDA(launch_pdl(lstm_fwd_kernel<ST>, dim3(cdiv(p.rows * p.H, 256)), dim3(256), 0, st, p));
  DIC_LAUNCH_CHECK();
  return 0;
}

// Backward through the cell for the first `rows` (valid) batch rows of one step.
struct LstmBwdArgs {
  const float* dh_carry;  // [splits][B,H] split-K partials of dh from step t+1 (W_hh, f_beta, decoder_att paths)
  long long dh_stride;    // elements between partials
  int dh_splits;          // number of partials
  int dh_rows;            // rows [0, dh_rows) carry a gradient from step t+1 (bs_valid of t+1); others 0
  const float* dh_out;    // [rows,H] d(logits).W_out for this step's packed rows
  const float* mask;      // [rows,H] dropout mask or null
  float* dc_carry;        // [B,H] in/out
  const float* acts;      // [rows,4H]
  const float* c_new;     // [rows,H]
  const float* c_prev;    // [rows,H]
  void* G;                // ST, row r at G + r*g_stride, columns [0,4H) = d(pre-activations)
  long long g_stride;
  int rows, H;
  TraceRec* trace;
};

template <typename ST>
__global__ void __launch_bounds__(256) lstm_bwd_kernel(const LstmBwdArgs p) {
  Trace trace(p.trace);
  // Forward state and the dc carry of step t+1 were written at least two launches ago: loaded before the
  // dependency wait.  NOT dHout: for the last step it comes from the GEMM launched right before this kernel.
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const bool live = idx < p.rows * p.H;
  const int H = p.H;
  const int r = live ? idx / H : 0, j = live ? idx - r * H : 0;
  float ig = 0.f, fg = 0.f, gg = 0.f, og = 0.f, m = 1.f, dho = 0.f, cn = 0.f, cp = 0.f, dcc = 0.f;
  if (live) {
    const float* a = p.acts + (size_t)r * 4 * H;
    ig = a[j]; fg = a[H + j]; gg = a[2 * H + j]; og = a[3 * H + j];
    m = p.mask ? p.mask[idx] : 1.f;
    cn = p.c_new[idx];
    cp = p.c_prev[idx];
    dcc = p.dc_carry[idx];
  }
  pdl_wait();
  pdl_trigger();
  trace.mark();
  if (!live) { trace.end(TK_LSTM_BWD); return; }
  dho = p.dh_out[idx];
  float dh = dho * m;
  if (r < p.dh_rows) {
    constexpr int kMaxSplits = 16;
    float part[kMaxSplits];
#pragma unroll
    for (int sp = 0; sp < kMaxSplits; ++sp)
      part[sp] = sp < p.dh_splits ? p.dh_carry[(size_t)sp * p.dh_stride + idx] : 0.f;
    float carry = 0.f;
#pragma unroll
    for (int sp = 0; sp < kMaxSplits; ++sp) carry += part[sp];
    dh += carry;
  }
  const float tc = tanhf(cn);
  const float dc = dcc + dh * og * (1.f - tc * tc);
  const float d_o = dh * tc;
  const float d_i = dc * gg, d_g = dc * ig, d_f = dc * cp;
  p.dc_carry[idx] = dc * fg;
  ST* G = reinterpret_cast<ST*>(p.G) + (size_t)r * p.g_stride;
  G[j] = from_f<ST>(d_i * ig * (1.f - ig));
  G[H + j] = from_f<ST>(d_f * fg * (1.f - fg));
  G[2 * H + j] = from_f<ST>(d_g * (1.f - gg * gg));
  G[3 * H + j] = from_f<ST>(d_o * og * (1.f - og));
  trace.end(TK_LSTM_BWD);
}

// out[i] = sum_s parts[s][i]  (final dh0 after the time loop)
__global__ void __launch_bounds__(256) reduce_parts_kernel(const float* __restrict__ parts, long long stride, int splits,
                                                           float* __restrict__ out, int n) {
  pdl_wait();
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int sp = 0; sp < splits; ++sp) s += parts[(size_t)sp * stride + i];
  out[i] = s;
}

template <typename ST>
inline int launch_lstm_bwd(const LstmBwdArgs& p_in, cudaStream_t st) {
  if (p_in.rows <= 0) return 0;
  LstmBwdArgs p = p_in;
  p.trace = g_trace_host;
  ProfScope prof(P_LSTM, st);
  DIC_CUDA(launch_pdl(lstm_bwd_kernel<ST>, dim3(cdiv(p.rows * p.H, 256)), dim3(256), 0, st, p));
  DIC_LAUNCH_CHECK();
  return 0;
}

}  // namespace dic
