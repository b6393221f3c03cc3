// C-ABI entry points (include/dic.h) and the host-side orchestration of the decoder path.
// One translation unit: kernels are header templates, this file instantiates and sequences them.
#include "attention.cuh"
#include "attention_bulk.cuh"
#include "attention_mma.cuh"
#include "attn_head.cuh"
#include "beam_fused.cuh"
#include "common.cuh"
#include "decode.cuh"
#include "depth_encoder.cuh"
#include "dfeat_tc.cuh"
#include "dp_allreduce.cuh"
#include "gemm_generic.cuh"
#include "gemm_tc.cuh"
#include "gates_lstm.cuh"
#include "layout.cuh"
#include "loss.cuh"
#include "lstm.cuh"
#include "misc.cuh"

namespace dic {
thread_local char g_err[512] = {0};
// event recorded by the next backward once every PARAMETER gradient is enqueued (before the dL/dF GEMM):
// data-parallel callers start their all-reduce on it and overlap it with the dL/dF work
// (process-wide, not thread-local: PyTorch runs backward on an autograd worker thread, not on the thread that armed it)
static std::atomic<cudaEvent_t> g_grads_ready_event{nullptr};
// optional earlier hand-over points of the same backward (buckets of the flat gradient buffer, in its order):
// the vocabulary-projection gradients (last two tensors) are final before the time loop starts, everything
// but the encoder_att pair (first two tensors) is final before the datt1 / dW_enc contraction
static std::atomic<cudaEvent_t> g_grads_lin_event{nullptr};
static std::atomic<cudaEvent_t> g_grads_mid_event{nullptr};
constexpr int kAllReduceSms = 20;   // SMs the persistent dL/dF GEMM leaves to an overlapped all-reduce

static int check_dims(const dic_dims* d, int dtype) {
  if (!d) DIC_FAIL(-1, "dims is null");
  if (dtype != DIC_F32 && dtype != DIC_BF16) DIC_FAIL(-1, "bad dtype %d", dtype);
  if (d->L <= 0 || d->D <= 0 || d->A <= 0 || d->E <= 0 || d->H <= 0 || d->V <= 0)
    DIC_FAIL(-1, "non-positive dimension");
  if (d->D % 8 || d->A % 4 || d->E % 4 || d->H % 4) DIC_FAIL(-1, "need D%%8==0, A,E,H%%4==0");
  if (dtype == DIC_BF16 && (d->E % 8 || d->H % 8 || d->A % 8))
    DIC_FAIL(-1, "bf16 mode needs A,E,H %% 8 == 0");
  if (d->A > 1024 || d->L > 1024) DIC_FAIL(-1, "A and L must be <= 1024");
  return 0;
}

static int make_sizes(const int32_t* host_bs, int T, int B, StepSizes* s, int* total) {
  if (T <= 0 || T > DIC_MAX_STEPS) DIC_FAIL(-1, "T=%d out of range 1..%d", T, DIC_MAX_STEPS);
  int tot = 0, prev = B;
  for (int t = 0; t < DIC_MAX_STEPS; ++t) s->n[t] = 0;
  for (int t = 0; t < T; ++t) {
    const int n = host_bs[t];
    if (n <= 0 || n > prev) DIC_FAIL(-1, "batch_sizes must be positive and non-increasing (t=%d n=%d)", t, n);
    s->n[t] = n;
    prev = n;
    tot += n;
  }
  *total = tot;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// GEMM dispatch: tcgen05 engine for bf16 K-major operands of suitable shape, FMA engine otherwise.
// ---------------------------------------------------------------------------------------------
static int gemm(const GemmArgs& g, cudaStream_t st) {
  if (tc_gemm_eligible(g)) return tc_gemm(g, st);
  return gemm_generic(g, st);
}

// Several buffers cleared by ONE launch.  The backward used to enqueue a cudaMemsetAsync in front of each of its
// twelve split-K / accumulation targets and four device-to-device copies for the bias slices: every such node is
// its own small grid with a full (non-programmatic) dependency on both sides, ~3 us of stream time each.
constexpr int kZeroJobsMax = 20;
struct ZeroJobs {
  void* p[kZeroJobsMax];
  unsigned long long bytes[kZeroJobsMax];
  int n = 0;
  void add(void* ptr, size_t b) {
    if (ptr && b && n < kZeroJobsMax) { p[n] = ptr; bytes[n] = b; ++n; }
  }
};
__global__ void __launch_bounds__(256) zero_many_kernel(const ZeroJobs j) {
  pdl_wait();          // launched programmatically (launch_pdl): nothing of a predecessor is touched before this
  pdl_trigger();
  const int job = blockIdx.y;
  char* base = reinterpret_cast<char*>(j.p[job]);
  const unsigned long long nb = j.bytes[job];
  const unsigned long long i0 = (unsigned long long)blockIdx.x * 256 + threadIdx.x, stride = (unsigned long long)gridDim.x * 256;
  if (((reinterpret_cast<uintptr_t>(base) | nb) & 15) == 0) {
    uint4* q = reinterpret_cast<uint4*>(base);
    for (unsigned long long i = i0; i < nb / 16; i += stride) q[i] = make_uint4(0u, 0u, 0u, 0u);
  } else {                                     // fp32 buffers: always 4-byte aligned
    float* q = reinterpret_cast<float*>(base);
    for (unsigned long long i = i0; i < nb / 4; i += stride) q[i] = 0.f;
  }
}
static int launch_zero_many(const ZeroJobs& j, cudaStream_t st) {
  if (j.n == 0) return 0;
  DIC_CUDA(launch_pdl(zero_many_kernel, dim3(64, j.n), dim3(256), 0, st, j));
  DIC_LAUNCH_CHECK();
  return 0;
}
// column sums of G -> the four bias gradients they belong to (b_ih and b_hh share the gate slice)
__global__ void __launch_bounds__(256) bias_scatter_kernel(const float* __restrict__ gsum, float* b_ih, float* b_hh, float* dec_att_b,
                                                           float* fbeta_b, int H4, int A, int D) {
  pdl_wait();          // launched programmatically (launch_pdl): nothing of a predecessor is touched before this
  pdl_trigger();
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < H4) { const float v = gsum[i]; b_ih[i] = v; b_hh[i] = v; }
  else if (i < H4 + A) dec_att_b[i - H4] = gsum[i];
  else if (i < H4 + A + D) fbeta_b[i - H4 - A] = gsum[i];
}

// Dense fp32 output with a long contraction and few output tiles (weight gradients, dHout):
// split K so the grid covers the SMs about twice; C is zero-filled (here, or by the caller: `zeroed`) and
// accumulated atomically.
static int gemm_splitk(GemmArgs g, cudaStream_t st, bool zeroed = false) {
  int s = 1;
  if (g.ldc == g.N && !g.c_bf16 && !g.accumulate && g.batch == 1) {
    if (tc_gemm_eligible(g)) {
      const long long tiles = (long long)cdiv(g.M, kTcBM) * cdiv(g.N, 128);
      const int kb = cdiv(g.K, kTcBK);
      s = (int)((2 * 148 + tiles - 1) / tiles);
      if (s > kb / 4) s = kb / 4;     // keep >= 4 k-blocks per CTA so the TMA/MMA pipeline fills
    } else {
      s = pick_splits(g.M, g.N, g.K);
    }
    if (s < 1) s = 1;
  }
  g.splits = s;
  g.split_mode = 0;
  if (s > 1 && !zeroed) DIC_CUDA(cudaMemsetAsync(g.C, 0, sizeof(float) * (size_t)g.M * g.ldc, st));
  return gemm(g, st);
}

// ---------------------------------------------------------------------------------------------
// shared prologue: Fsum, meanF, att1 (K0)
// ---------------------------------------------------------------------------------------------
template <typename ST>
static int prologue(const dic_dims& d, const Pack& pk, const void* f_rgb, const void* f_depth,
                    int feat_dtype, int B, ST* Fsum, float* meanF, bf16* mean16, ST* att1, const ST** Fuse,
                    cudaStream_t st) {
  const int is_bf16 = sizeof(ST) == 2;
  const bool alias = (f_depth == nullptr) && ((feat_dtype == DIC_BF16) == (is_bf16 != 0));
  // base decoders with matching storage: no copy, only the mean (base_caption_models.py:118)
  DIC_TRY(launch_fuse_feats<ST>(f_rgb, f_depth, feat_dtype == DIC_BF16, alias ? nullptr : Fsum, meanF, mean16, B,
                                d.L, d.D, st));
  const ST* F = alias ? reinterpret_cast<const ST*>(f_rgb) : Fsum;
  *Fuse = F;
  // att1 = F . W_enc^T + b_enc   (attention.py:84, hoisted out of the time loop)
  GemmArgs g = gemm_args_nt(F, is_bf16, d.D, pk.Wenc(), is_bf16, d.D, att1, is_bf16, d.A, B * d.L, d.A,
                            d.D, pk.b_enc());
  g.prof = P_GEMM_ATT1;
  g.prof_bytes = (double)B * d.L * ((double)d.D + d.A) * sizeof(ST);
  DIC_TRY(gemm(g, st));
  return 0;
}

template <typename ST>
static int init_state_gemm(const dic_dims& d, const Pack& pk, const float* meanF, int B, float* h0,
                           float* c0, cudaStream_t st) {
  const int is_bf16 = sizeof(ST) == 2;
  // h0, c0 = chunk(init_linear(mean_l F), 2)   (depth_models.py:166-168); dense fp32 [B,H] outputs,
  // K = D is long and the output tiny, so the contraction is split across the SMs
  GemmArgs g = gemm_args_nt(meanF, 0, d.D, pk.Winit(), is_bf16, d.D, h0, 0, d.H, B, d.H, d.D, pk.b_init());
  DIC_TRY(gemm_splitk(g, st));
  const char* w2 = reinterpret_cast<const char*>(pk.Winit()) + (size_t)d.H * d.D * sizeof(ST);
  g = gemm_args_nt(meanF, 0, d.D, w2, is_bf16, d.D, c0, 0, d.H, B, d.H, d.D, pk.b_init() + d.H);
  DIC_TRY(gemm_splitk(g, st));
  return 0;
}

// bf16 mode: [h0 | c0] in ONE tensor-core GEMM (split-K) from the bf16 copy of mean_l F
static int init_state_gemm_tc(const dic_dims& d, const Pack& pk, const bf16* mean16, int B, float* hc0,
                              cudaStream_t st) {
  GemmArgs g = gemm_args_nt(mean16, 1, d.D, pk.Winit(), 1, d.D, hc0, 0, 2 * d.H, B, 2 * d.H, d.D, pk.b_init());
  return gemm_splitk(g, st);
}

// h-projection of a step: [att2 | beta] = [h W_dec^T + b_dec | sigmoid(h W_beta^T + b_beta)]
// (attention.py:85, depth_models.py:189)
template <typename ST>
static int hproj(const dic_dims& d, const Pack& pk, const ST* h, long long h_ld, int rows, float* HP,
                 cudaStream_t st) {
  const int is_bf16 = sizeof(ST) == 2;
  GemmArgs g = gemm_args_nt(h, is_bf16, h_ld, pk.Wdb(d), is_bf16, d.H, HP, 0, d.A + d.D, rows,
                            d.A + d.D, d.H, pk.bias_db());
  g.sig_lo = d.A;
  g.sig_hi = d.A + d.D;
  g.fast_act = is_bf16;
  g.tag = 1;
  g.b_static = 1;
  return gemm(g, st);
}

// gate pre-activations as split-K partials: [emb|zg|h] . [W_ih|W_hh]^T  (nn.LSTMCell)
template <typename ST>
static int gates_gemm(const dic_dims& d, const Pack& pk, const ST* X, long long XW, int rows,
                      int rows_alloc, float* gate_part, int* splits_out, cudaStream_t st) {
  const int is_bf16 = sizeof(ST) == 2;
  GemmArgs g = gemm_args_nt(X, is_bf16, XW, pk.Wg(), is_bf16, XW, gate_part, 0, 4 * d.H, rows, 4 * d.H,
                            (int)XW, nullptr);
  // only a handful of 128x128 output tiles exist (M = batch, N = 4H): split K across the SMs
  // (as many as it takes to cover the SMs about twice; large row counts need few or none)
  int s;
  if (tc_gemm_eligible(g)) {
    // 64-column tiles and just enough K splits for one wave: a CTA's time is fixed latency + the store
    // of its fp32 partial tile (128x128 took 2.1 us of a 4.5 us CTA), and lstm re-reads every partial
    const long long tiles = (long long)cdiv(rows, kTcBM) * cdiv(4 * d.H, 64);
    s = (int)(tc_num_sms() / tiles);
    const int kb = cdiv((int)XW, kTcBK);
    if (s > kb / 2) s = kb / 2;
    g.bn = 64;
  } else {
    s = pick_splits(rows, 4 * d.H, (int)XW);
  }
  if (s > kGateSplitsMax) s = kGateSplitsMax;
  if (s < 1) s = 1;
  g.splits = s;
  g.split_mode = 1;
  g.split_stride = (long long)rows_alloc * 4 * d.H;
  g.tag = 2;
  g.b_static = 1;
  *splits_out = s;
  return gemm(g, st);
}

// the same split-K gate GEMM over an arbitrary operand pair (look-ahead beam step: parent-order rows [beta.z | h]
// against the columns [E, E+D+H) of the packed [W_ih | W_hh])
template <typename ST>
static int gates_gemm_from(const dic_dims& d, const ST* Aop, long long lda, const ST* Bop, long long ldb, int Kdim,
                           int rows, int rows_alloc, float* gate_part, int* splits_out, cudaStream_t st) {
  const int is_bf16 = sizeof(ST) == 2;
  GemmArgs g = gemm_args_nt(Aop, is_bf16, lda, Bop, is_bf16, ldb, gate_part, 0, 4 * d.H, rows, 4 * d.H, Kdim, nullptr);
  int s;
  if (tc_gemm_eligible(g)) {
    // A CTA's time is what it pulls through TMA (~55 GB/s per SM): (128 + BN) x 64 x 2 bytes per K block, K blocks
    // = ceil(K / 64 / splits), splits = SMs / tiles.  640 rows: 64-column tiles = 40 tiles x 3 splits x 12 blocks x
    // 24 KB = 288 KB per CTA, 128-column tiles = 20 tiles x 7 splits x 5 blocks x 32 KB = 160 KB per CTA.
    const int kb = cdiv(Kdim, kTcBK);
    long long best = -1;
    s = 1;
    for (int bn = 64; bn <= 128; bn *= 2) {
      const long long tiles = (long long)cdiv(rows, kTcBM) * cdiv(4 * d.H, bn);
      int sp = (int)(tc_num_sms() / tiles);
      if (sp > kb / 2) sp = kb / 2;
      if (sp > kGateSplitsMax) sp = kGateSplitsMax;
      if (sp < 1) sp = 1;
      const long long bytes = (long long)(kTcBM + bn) * kTcBK * 2 * cdiv(kb, sp);
      if (best < 0 || bytes < best) { best = bytes; s = sp; g.bn = bn; }
    }
  } else {
    s = pick_splits(rows, 4 * d.H, Kdim);
  }
  if (s > kGateSplitsMax) s = kGateSplitsMax;
  if (s < 1) s = 1;
  g.splits = s;
  g.split_mode = 1;
  g.split_stride = (long long)rows_alloc * 4 * d.H;
  g.tag = 2;
  g.b_static = 1;
  *splits_out = s;
  return gemm(g, st);
}

// one LSTM step: cluster-fused GEMM + pointwise (bf16 mode, gates_lstm.cuh) or split-K GEMM + lstm kernel
template <typename ST>
static int lstm_step(const dic_dims& d, const Pack& pk, const ST* X, long long XW, int rows, int rows_alloc,
                     float* gate_part, LstmFwdArgs l, cudaStream_t st) {
  if constexpr (sizeof(ST) == 2) {
    // one wave of 8-CTA clusters only: with more tiles than that the split-K GEMM spreads the work better
    if (gates_lstm_eligible(d.H, (int)XW, X, XW) && cdiv(rows, kTcBM) * (4 * d.H / kGlBN) * kGlSplits <= tc_num_sms())
      return launch_gates_lstm(reinterpret_cast<const bf16*>(X), XW, reinterpret_cast<const bf16*>(pk.Wgp()), rows,
                               d.H, (int)XW, l, st);
  }
  int splits = 1;
  DIC_TRY(gates_gemm<ST>(d, pk, X, XW, rows, rows_alloc, gate_part, &splits, st));
  l.gate_part = gate_part;
  l.part_stride = (long long)rows_alloc * 4 * d.H;
  l.splits = splits;
  return launch_lstm_fwd<ST>(l, st);
}

static bool handoff_enabled() {
  static int v = -1;
  // off by default: three same-box A/B runs of 60 steps gave 2.989 ms (grid-level wait) vs 2.996 ms (hand-off)
  if (v < 0) { const char* e = getenv("DIC_HANDOFF"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

// Number of sub-batch streams of a time loop.  Measured on B200 (scripts/sub_sweep.py, gpurun_out/sweep1.log):
// 2 streams gain 1% on the training step and lose 10% on decode, more streams lose everywhere -- the
// streaming kernels fill every SM's register file (8 CTAs x 64 regs x 128 threads), so another
// stream's kernels cannot become resident next to them and only the launch count grows.  Hence 1
// unless the caller asks (dic_set_substreams); the facility stays for shapes where it pays.
static int pick_substreams(int images, int training) {
  (void)images; (void)training;
  const int o = g_sub_override.load();
  if (o > 0 && !g_prof.on) return o;    // per-kernel event timing wants kernels timed alone
  return 1;
}

// =============================================================================================
// training forward
// =============================================================================================
template <typename ST>
static int decoder_forward_impl(const dic_dims& d, int attn_mode, const void* pack, const void* f_rgb,
                                const void* f_depth, int feat_dtype, const int64_t* captions,
                                int cap_stride, const StepSizes& sizes, int total, int T, int B,
                                const float* u, float temp, const float* dropout_mask, void* logits,
                                int logits_bf16,
                                float* alphas, char* ws, cudaStream_t st) {
  const int dtype = sizeof(ST) == 2 ? DIC_BF16 : DIC_F32;
  const int is_bf16 = sizeof(ST) == 2;
  const Pack pk(pack, d, dtype);
  const TrainLayout lay(d, dtype, B, T);
  const size_t XW = lay.XW;
  ST* Fsum = reinterpret_cast<ST*>(ws + lay.Fsum);
  float* meanF = reinterpret_cast<float*>(ws + lay.meanF);
  ST* att1 = reinterpret_cast<ST*>(ws + lay.att1);
  ST* XH = reinterpret_cast<ST*>(ws + lay.XH);
  float* HP = reinterpret_cast<float*>(ws + lay.HP);
  float* Z = reinterpret_cast<float*>(ws + lay.Z);
  float* acts = reinterpret_cast<float*>(ws + lay.acts);
  float* c_all = reinterpret_cast<float*>(ws + lay.c_all);
  float* gate_part = reinterpret_cast<float*>(ws + lay.gate_part);
  ST* Hdrop = reinterpret_cast<ST*>(ws + lay.Hdrop);

  // invalid (b >= bs_valid) rows of the step buffers must stay finite zeros: the post-loop
  // weight-gradient GEMMs run over all T*B rows.
  // (a batch of equal-length captions writes every row of these buffers: nothing to clear)
  const bool ragged = sizes.n[T - 1] < B;
  if (ragged) DIC_CUDA(cudaMemsetAsync(XH, 0, ((size_t)T * B + B) * XW * sizeof(ST), st));
  bf16* alpha16 = is_bf16 ? reinterpret_cast<bf16*>(ws + lay.alpha16) : nullptr;
  if (alpha16 && ragged) DIC_CUDA(cudaMemsetAsync(alpha16, 0, (size_t)T * B * lay.Lp * 2, st));

  const ST* F = nullptr;
  bf16* mean16 = is_bf16 ? reinterpret_cast<bf16*>(ws + lay.meanF16) : nullptr;
  DIC_TRY(prologue<ST>(d, pk, f_rgb, f_depth, feat_dtype, B, Fsum, meanF, mean16, att1, &F, st));
  if (is_bf16) {
    float* hc0 = reinterpret_cast<float*>(ws + lay.hc0);
    DIC_TRY(init_state_gemm_tc(d, pk, mean16, B, hc0, st));
    DIC_TRY(launch_copy2d(hc0, 2 * d.H, XH + d.E + d.D, (long long)XW, 1, B, d.H, st));
    DIC_TRY(launch_copy2d(hc0 + d.H, 2 * d.H, c_all, d.H, 0, B, d.H, st));
  } else {
    float* h0 = reinterpret_cast<float*>(ws + lay.h0);
    DIC_TRY(init_state_gemm<ST>(d, pk, meanF, B, h0, c_all, st));
    DIC_TRY(launch_copy2d(h0, d.H, XH + d.E + d.D, (long long)XW, is_bf16, B, d.H, st));
  }

  // all-step embedding gather (depth_models.py:160)
  {
    dim3 grid(cdiv(B * d.E, 256), T);
    DIC_CUDA(launch_pdl(embed_gather_tf_kernel<ST>, grid, dim3(256), 0, st, reinterpret_cast<const ST*>(pk.Emb()), captions,
                        cap_stride, XH, (long long)XW, (long long)B * XW, B, d.E, d.V, sizes, T));
    DIC_LAUNCH_CHECK();
  }

  // per-image alpha -> context hand-off (soft / Gumbel-softmax; the one-hot Gumbel-max context kernel keeps the
  // grid-level wait): flags cleared once per call, epoch = step + 1
  unsigned int* ready = reinterpret_cast<unsigned int*>(ws + lay.ready);
  const bool handoff = handoff_enabled() && attn_mode != DIC_ATTN_GUMBEL_MAX && pdl_enabled();
  if (handoff) DIC_CUDA(cudaMemsetAsync(ready, 0, sizeof(unsigned int) * B, st));
  // time loop: S sub-batches of images on S streams (see common.cuh, "sub-batch streams")
  int offs[DIC_MAX_STEPS + 1];
  offs[0] = 0;
  for (int t = 0; t < T; ++t) offs[t + 1] = offs[t] + sizes.n[t];
  int r0s[kMaxSub + 1];
  const int S = sub_bounds(B, pick_substreams(B, 1), 8, r0s);
  cudaStream_t ss[kMaxSub];
  DIC_TRY(sub_fork(st, S, ss));
  for (int t = 0; t < T; ++t) {
    for (int sb = 0; sb < S; ++sb) {
      const int r0 = r0s[sb];
      const int n = (sizes.n[t] < r0s[sb + 1] ? sizes.n[t] : r0s[sb + 1]) - r0;
      if (n <= 0) continue;
      cudaStream_t sst = ss[sb];
      const int off = offs[t] + r0;                       // packed row of this sub-batch's first image
      ST* X = XH + ((size_t)t * B + r0) * XW;
      ST* Xn = XH + ((size_t)(t + 1) * B + r0) * XW;
      float* HPt = HP + ((size_t)t * B + r0) * (d.A + d.D);
      // bf16 storage at the reference shape: h-projection + energies + softmax in one launch (attn_head.cuh)
      const bool head = is_bf16 && !handoff && attn_head_eligible(d.A, d.H, d.D, d.L, 1);
      if (!head) DIC_TRY(hproj<ST>(d, pk, X + d.E + d.D, (long long)XW, n, HPt, sst));

      AttnFwdArgs a;
      memset(&a, 0, sizeof(a));
      a.F = F + (size_t)r0 * d.L * d.D; a.att1 = att1 + (size_t)r0 * d.L * d.A; a.hp = HPt;
      a.w_full = pk.w_full(); a.b_full = pk.b_full();
      a.u = u ? u + (size_t)off * d.L : nullptr;
      a.alpha_out = alphas + ((size_t)r0 * T + t) * d.L;
      a.alpha_stride = (long long)T * d.L;
      a.alpha16_out = alpha16 ? alpha16 + ((size_t)r0 * T + t) * lay.Lp : nullptr;
      a.alpha16_stride = (long long)T * lay.Lp;
      a.alpha16_width = lay.Lp;
      if (handoff) { a.ready = ready + r0; a.epoch = (unsigned int)(t + 1); }
      a.z_out = Z + ((size_t)t * B + r0) * d.D;
      a.zg_out = X + d.E;
      a.zg_stride = (long long)XW;
      a.L = d.L; a.D = d.D; a.A = d.A;
      a.mode = attn_mode;
      a.inv_temp = attn_mode == DIC_ATTN_GUMBEL_SOFTMAX ? 1.f / temp : 1.f;
      if (head) {
        HeadArgs hd;
        hd.h = reinterpret_cast<const bf16*>(X + d.E + d.D); hd.h_ld = (long long)XW;
        hd.Wdb = reinterpret_cast<const bf16*>(pk.Wdb(d)); hd.bias_db = pk.bias_db();
        hd.HP = HPt; hd.a = a; hd.rows = n;
        DIC_TRY(launch_attn_head(hd, 1, sst));
        a.skip_alpha = 1;
      }
      DIC_TRY(launch_attn_step<ST>(a, n, 1, sst));

      float* gp = gate_part + (size_t)r0 * 4 * d.H;

      LstmFwdArgs l;
      memset(&l, 0, sizeof(l));
      l.bias_g = pk.bias_g();
      l.c_in = c_all + ((size_t)t * B + r0) * d.H;
      l.c_out = c_all + ((size_t)(t + 1) * B + r0) * d.H;
      l.acts = acts + ((size_t)t * B + r0) * 4 * d.H;
      l.h_out = Xn + d.E + d.D; l.h_stride = (long long)XW;
      l.hdrop_out = Hdrop + (size_t)off * d.H;
      l.mask = dropout_mask ? dropout_mask + (size_t)off * d.H : nullptr;
      l.rows = n; l.H = d.H;
      DIC_TRY(lstm_step<ST>(d, pk, X, (long long)XW, n, B, gp, l, sst));
    }
  }
  DIC_TRY(sub_join(st, S));

  // logits for every packed row at once: linear(dropout(h)) (depth_models.py:197), written in
  // PackedSequence (time-major) order
  GemmArgs g = gemm_args_nt(Hdrop, is_bf16, d.H, pk.Wout(), is_bf16, d.H, logits, logits_bf16, d.V, total, d.V,
                            d.H, pk.b_out());
  g.prof = P_GEMM_LOGITS;
  g.prof_bytes = (double)total * d.V * (logits_bf16 ? 2 : 4);
  DIC_TRY(gemm(g, st));
  return 0;
}

// =============================================================================================
// training backward
// =============================================================================================
template <typename ST>
static int decoder_backward_impl(const dic_dims& d, int attn_mode, const void* pack,
                                 const int64_t* captions, int cap_stride, const StepSizes& sizes,
                                 int total, int T, int B, const void* d_logits, int dl_is_st, const float* d_alphas,
                                 const float* alphas, float temp, const float* dropout_mask,
                                 const dic_params& gr, void* d_feats, int dfeat_bf16, const void* f_rgb_alias, char* ws,
                                 cudaStream_t st) {
  const int dtype = sizeof(ST) == 2 ? DIC_BF16 : DIC_F32;
  const int is_bf16 = sizeof(ST) == 2;
  const Pack pk(pack, d, dtype);
  const TrainLayout lay(d, dtype, B, T);
  const long long XW = (long long)lay.XW, GW = (long long)lay.GW;
  const int H = d.H, A = d.A, D = d.D, E = d.E, L = d.L, V = d.V;
  const size_t TB = (size_t)T * B;
  const ST* F = f_rgb_alias ? reinterpret_cast<const ST*>(f_rgb_alias) : reinterpret_cast<ST*>(ws + lay.Fsum);
  float* meanF = reinterpret_cast<float*>(ws + lay.meanF);
  ST* att1 = reinterpret_cast<ST*>(ws + lay.att1);
  ST* XH = reinterpret_cast<ST*>(ws + lay.XH);
  float* HP = reinterpret_cast<float*>(ws + lay.HP);
  float* Z = reinterpret_cast<float*>(ws + lay.Z);
  float* acts = reinterpret_cast<float*>(ws + lay.acts);
  float* c_all = reinterpret_cast<float*>(ws + lay.c_all);
  ST* Hdrop = reinterpret_cast<ST*>(ws + lay.Hdrop);
  ST* G = reinterpret_cast<ST*>(ws + lay.G);
  ST* DZ = reinterpret_cast<ST*>(ws + lay.DZ);
  float* de = reinterpret_cast<float*>(ws + lay.de);
  float* dzg = reinterpret_cast<float*>(ws + lay.dzg);
  float* dh = reinterpret_cast<float*>(ws + lay.dh);
  float* dc = reinterpret_cast<float*>(ws + lay.dc);
  float* dHout = reinterpret_cast<float*>(ws + lay.dHout);
  float* dwfull_part = reinterpret_cast<float*>(ws + lay.dwfull_part);
  float* dbfull_part = reinterpret_cast<float*>(ws + lay.dbfull_part);
  ST* datt1 = reinterpret_cast<ST*>(ws + lay.datt1);
  float* dXemb = reinterpret_cast<float*>(ws + lay.dXemb);
  float* dmeanF = reinterpret_cast<float*>(ws + lay.dmeanF);

  if (attn_mode == DIC_ATTN_GUMBEL_MAX) DIC_FAIL(-1, "gumbel-max (Hard_sample) is a no_grad path");

  // rows of inactive (t, b) pairs must read as zeros in the post-loop contractions over all T*B rows;
  // with equal-length captions every row is written by the loop below and nothing needs clearing
  if (sizes.n[T - 1] < B) {
    DIC_CUDA(cudaMemsetAsync(G, 0, TB * GW * sizeof(ST), st));
    DIC_CUDA(cudaMemsetAsync(DZ, 0, TB * D * sizeof(ST), st));
    DIC_CUDA(cudaMemsetAsync(de, 0, TB * L * sizeof(float), st));
    DIC_CUDA(cudaMemsetAsync(dwfull_part, 0, sizeof(float) * TB * A, st));
    DIC_CUDA(cudaMemsetAsync(dbfull_part, 0, sizeof(float) * TB, st));
  }
  {
    // everything this backward accumulates into, cleared by one launch (see ZeroJobs)
    ZeroJobs z;
    z.add(dh, sizeof(float) * B * H);
    z.add(dc, sizeof(float) * B * H);
    z.add(dHout, sizeof(float) * (size_t)total * H);
    z.add(gr.lin_w, sizeof(float) * (size_t)V * H);
    z.add(gr.init_w, sizeof(float) * (size_t)2 * H * D);
    z.add(gr.w_ih, sizeof(float) * (size_t)4 * H * (E + D));
    z.add(gr.w_hh, sizeof(float) * (size_t)4 * H * H);
    z.add(gr.dec_att_w, sizeof(float) * (size_t)A * H);
    z.add(gr.fbeta_w, sizeof(float) * (size_t)D * H);
    z.add(gr.enc_att_w, sizeof(float) * (size_t)A * D);
    z.add(gr.embed_w, sizeof(float) * (size_t)V * E);
    z.add(dzg, sizeof(float) * (size_t)B * D);         // red.add target of the per-step dzg GEMM halves (re-zeroed by its reader)
    // destinations of the column-sum kernels (atomic partial sums)
    z.add(gr.lin_b, sizeof(float) * V);
    z.add(ws + lay.tmpvec, sizeof(float) * GW);
    z.add(gr.full_att_w, sizeof(float) * A);
    z.add(gr.full_att_b, sizeof(float) * 1);
    z.add(gr.enc_att_b, sizeof(float) * A);
    DIC_TRY(launch_zero_many(z, st));
  }

  // bf16 mode: the two contractions over d_logits take a bf16 copy (tensor-core operand)
  const void* dl = d_logits;
  int dl_bf16 = 0;
  if (is_bf16 && dl_is_st) {
    dl_bf16 = 1;          // the fused loss head already wrote d_logits in bf16
  } else if (is_bf16) {
    void* dl16 = ws + lay.dlogits16;
    DIC_TRY(launch_copy2d(reinterpret_cast<const float*>(d_logits), V, dl16, V, 1, total, V, st));
    dl = dl16;
    dl_bf16 = 1;
  }

  // dHout = d_logits . W_out   (all packed rows)
  {
    GemmArgs g = gemm_args_nt(dl, dl_bf16, V, pk.Wout(), is_bf16, 0, dHout, 0, H, total, H, V, nullptr);
    g.b_n = 1; g.b_k = H;
    DIC_TRY(gemm_splitk(g, st, true));
  }

  // linear (vocabulary projection): its gradients need d_logits and the saved dropout(h) only, so they are
  // final before the time loop starts -- a data-parallel caller reduces them under the whole loop
  {
    GemmArgs g = gemm_args_nt(dl, dl_bf16, 0, Hdrop, is_bf16, 0, gr.lin_w, 0, H, V, H, total, nullptr);
    g.a_m = 1; g.a_k = V; g.b_n = 1; g.b_k = H;
    DIC_TRY(gemm_splitk(g, st, true));
    DIC_TRY(launch_colsum(dl, dl_bf16, total, V, V, gr.lin_b, st, true));   // bf16 mode: half the bytes
    if (cudaEvent_t ev = g_grads_lin_event.exchange(nullptr)) DIC_CUDA(cudaEventRecord(ev, st));
  }

  const float inv_temp = attn_mode == DIC_ATTN_GUMBEL_SOFTMAX ? 1.f / temp : 1.f;
  float* dh_part = reinterpret_cast<float*>(ws + lay.dh_part);
  int dh_splits = cdiv((int)GW, kTcBK) / 4;      // >= 4 k-blocks of 64 per split
  if (dh_splits > kDhSplitsMax) dh_splits = kDhSplitsMax;
  if (dh_splits < 1) dh_splits = 1;
  const bool dzg_split_ok = is_bf16 && tc_enabled() && cdiv(4 * H, kTcBK) >= 2 * kDzgSplits && D % 8 == 0 &&
                            (long long)B * D >= 128LL * 128LL;
  int offs[DIC_MAX_STEPS + 1];
  offs[0] = 0;
  for (int t = 0; t < T; ++t) offs[t + 1] = offs[t] + sizes.n[t];
  int r0s[kMaxSub + 1];
  const int S = sub_bounds(B, pick_substreams(B, 1), 8, r0s);
  cudaStream_t ss[kMaxSub];
  DIC_TRY(sub_fork(st, S, ss));
  for (int t = T - 1; t >= 0; --t) {
    for (int sb = 0; sb < S; ++sb) {
      const int r0 = r0s[sb];
      const int n = (sizes.n[t] < r0s[sb + 1] ? sizes.n[t] : r0s[sb + 1]) - r0;
      if (n <= 0) continue;
      cudaStream_t sst = ss[sb];
      const int off = offs[t] + r0;
      ST* Gt = G + ((size_t)t * B + r0) * GW;
      float* HPt = HP + ((size_t)t * B + r0) * (A + D);
      float* dhp = dh_part + (size_t)r0 * H;
      float* dzg_s = dzg + (size_t)r0 * D;
      int dzg_splits = 1;

      LstmBwdArgs lb;
      memset(&lb, 0, sizeof(lb));
      lb.dh_carry = dhp; lb.dh_stride = (long long)B * H; lb.dh_splits = dh_splits;
      // rows (of this sub-batch) that were active one step later
      int nxt = (t + 1 < T) ? (sizes.n[t + 1] < r0s[sb + 1] ? sizes.n[t + 1] : r0s[sb + 1]) - r0 : 0;
      lb.dh_rows = nxt > 0 ? nxt : 0;
      lb.dh_out = dHout + (size_t)off * H;
      lb.mask = dropout_mask ? dropout_mask + (size_t)off * H : nullptr;
      lb.dc_carry = dc + (size_t)r0 * H;
      lb.acts = acts + ((size_t)t * B + r0) * 4 * H;
      lb.c_new = c_all + ((size_t)(t + 1) * B + r0) * H;
      lb.c_prev = c_all + ((size_t)t * B + r0) * H;
      lb.G = Gt; lb.g_stride = GW; lb.rows = n; lb.H = H;
      DIC_TRY(launch_lstm_bwd<ST>(lb, sst));

      // dzg = dgates . W_ih[:, E:E+D]
      {
        GemmArgs g = gemm_args_nt(Gt, is_bf16, GW, reinterpret_cast<const ST*>(pk.Wg()) + E, is_bf16, 0, dzg_s,
                                  0, D, n, D, 4 * H, nullptr);
        g.b_n = 1; g.b_k = XW;
        g.tag = 3;
        g.b_static = 1;
        // two K halves accumulated with red.add into a buffer that is zero on entry: 128 CTAs instead of 64.
        // The buffer is cleared once before the loop and again by the streaming kernel right after it has read
        // it (no memset node inside the PDL kernel chain); two addends onto zero: the result is order independent.
        dzg_splits = dzg_split_ok ? kDzgSplits : 1;
        g.splits = dzg_splits;
        g.split_mode = 0;
        DIC_TRY(gemm(g, sst));
      }

      AttnBwdArgs ab;
      memset(&ab, 0, sizeof(ab));
      ab.F = F + (size_t)r0 * L * D; ab.att1 = att1 + (size_t)r0 * L * A; ab.hp = HPt;
      ab.z = Z + ((size_t)t * B + r0) * D; ab.dzg = dzg_s; ab.dzg_rezero = dzg_splits > 1;
      ab.alpha = alphas + ((size_t)r0 * T + t) * L; ab.alpha_stride = (long long)T * L;
      ab.dalpha = d_alphas ? d_alphas + ((size_t)r0 * T + t) * L : nullptr;
      ab.w_full = pk.w_full();
      ab.G = Gt; ab.g_stride = GW; ab.gcol_att2 = 4 * H; ab.gcol_beta = 4 * H + A;
      ab.DZ = DZ + ((size_t)t * B + r0) * D;
      ab.de_out = de + ((size_t)t * B + r0) * L;
      ab.dwfull_part = dwfull_part + ((size_t)t * B + r0) * A;
      ab.dbfull_part = dbfull_part + ((size_t)t * B + r0);
      ab.dal_part = reinterpret_cast<float*>(ws + lay.dal_part) + (size_t)r0 * L;
      ab.part_rows = B;
      ab.L = L; ab.D = D; ab.A = A; ab.inv_temp = inv_temp;
      DIC_TRY(launch_attn_bwd<ST>(ab, n, sst));

      // dh_{t-1} = [dgates | datt2 | dbeta'] . [W_hh ; W_dec ; W_beta]
      {
        // K = 4H+A+D is long and the output tiny: split-K into partial buffers that the next
        // lstm_bwd (and the final reduce) sum in a fixed order -- no memset, no atomics
        GemmArgs g = gemm_args_nt(Gt, is_bf16, GW, pk.Whdb(), is_bf16, 0, dhp, 0, H, n, H, (int)GW, nullptr);
        g.b_n = 1; g.b_k = H;
        g.splits = dh_splits;
        g.split_mode = 1;
        g.split_stride = (long long)B * H;
        g.tag = 4;
        g.b_static = 1;
        DIC_TRY(gemm(g, sst));
      }
    }
  }
  DIC_TRY(sub_join(st, S));
  // dh0 (rows [0, bs_valid[0]) = all B rows)
  DIC_CUDA(launch_pdl(reduce_parts_kernel, dim3(cdiv(B * H, 256)), dim3(256), 0, st, (const float*)dh_part,
                      (long long)B * H, dh_splits, dh, B * H));
  DIC_LAUNCH_CHECK();

  // ---- post-loop: everything that is a sum over (t, b) is one contraction over T*B rows ----
  // init_linear: rows [0,H) from dh0, rows [H,2H) from dc0
  if (is_bf16) {
    // bf16 operands [dh0|dc0] and mean_l F -> two tensor-core GEMMs; the bias gradient rides on the cast
    bf16* dhc16 = reinterpret_cast<bf16*>(ws + lay.dhc16);
    const bf16* mean16 = reinterpret_cast<const bf16*>(ws + lay.meanF16);
    DIC_CUDA(launch_pdl(dhc_prep_kernel, dim3(cdiv(2 * H, 32)), dim3(256), 0, st, dh, dc, dhc16, gr.init_b, B, H));
    DIC_LAUNCH_CHECK();
    {
      GemmArgs g = gemm_args_nt(dhc16, 1, 0, mean16, 1, 0, gr.init_w, 0, D, 2 * H, D, B, nullptr);
      g.a_m = 1; g.a_k = 2 * H; g.b_n = 1; g.b_k = D;
      DIC_TRY(gemm_splitk(g, st, true));
    }
    if (d_feats) {
      GemmArgs m = gemm_args_nt(dhc16, 1, 2 * H, pk.Winit(), 1, 0, dmeanF, 0, D, B, D, 2 * H, nullptr);
      m.b_n = 1; m.b_k = D;
      DIC_TRY(gemm(m, st));
    }
  } else {
    for (int half = 0; half < 2; ++half) {
      const float* dsrc = half == 0 ? dh : dc;
      GemmArgs g = gemm_args_nt(dsrc, 0, 0, meanF, 0, 0, gr.init_w + (size_t)half * H * D, 0, D, H, D, B, nullptr);
      g.a_m = 1; g.a_k = H; g.b_n = 1; g.b_k = D;
      DIC_TRY(gemm_generic(g, st));
      DIC_TRY(launch_colsum(dsrc, 0, B, H, H, gr.init_b + half * H, st));
      if (d_feats) {
        GemmArgs m = gemm_args_nt(dsrc, 0, H, reinterpret_cast<const ST*>(pk.Winit()) + (size_t)half * H * D,
                                  is_bf16, 0, dmeanF, 0, D, B, D, H, nullptr);
        m.b_n = 1; m.b_k = D; m.accumulate = half;
        DIC_TRY(gemm_generic(m, st));
      }
    }
  }

  // biases of the LSTM / decoder_att / f_beta: column sums of G
  // (one pass over G, then the four slices are copied out)
  {
    float* gsum = reinterpret_cast<float*>(ws + lay.tmpvec);
    DIC_TRY(launch_colsum(G, is_bf16, (int)TB, (int)GW, GW, gsum, st, true));
    DIC_CUDA(launch_pdl(bias_scatter_kernel, dim3(cdiv(4 * H + A + D, 256)), dim3(256), 0, st, (const float*)gsum, gr.b_ih, gr.b_hh,
                        gr.dec_att_b, gr.fbeta_b, 4 * H, A, D));
    DIC_LAUNCH_CHECK();
  }

  auto wgrad = [&](const ST* Aop, long long a_ld, int M, const ST* Bop, long long b_ld, int N, int K,
                   float* C, long long ldc) -> int {
    GemmArgs g = gemm_args_nt(Aop, is_bf16, 0, Bop, is_bf16, 0, C, 0, ldc, M, N, K, nullptr);
    g.a_m = 1; g.a_k = a_ld; g.b_n = 1; g.b_k = b_ld;
    return gemm_splitk(g, st, true);
  };
  // [dW_ih | dW_hh] = dgates^T . [emb|zg|h]
  DIC_TRY(wgrad(G, GW, 4 * H, XH, XW, E + D, (int)TB, gr.w_ih, E + D));
  DIC_TRY(wgrad(G, GW, 4 * H, XH + E + D, XW, H, (int)TB, gr.w_hh, H));
  // dW_dec = datt2^T . h_prev ; dW_beta = dbeta'^T . h_prev
  DIC_TRY(wgrad(G + 4 * H, GW, A, XH + E + D, XW, H, (int)TB, gr.dec_att_w, H));
  DIC_TRY(wgrad(G + 4 * H + A, GW, D, XH + E + D, XW, H, (int)TB, gr.fbeta_w, H));

  // embedding: dX_emb = dgates . W_ih[:, :E], scattered by token id
  {
    GemmArgs g = gemm_args_nt(G, is_bf16, GW, pk.Wg(), is_bf16, 0, dXemb, 0, E, (int)TB, E, 4 * H, nullptr);
    g.b_n = 1; g.b_k = XW;
    DIC_TRY(gemm(g, st));
    // (gr.embed_w was cleared with the other accumulation targets at the top)
    dim3 grid(cdiv(B * E, 256), T);
    DIC_CUDA(launch_pdl(embed_scatter_add_kernel, grid, dim3(256), 0, st, dXemb, captions, cap_stride, gr.embed_w, B, E, V, sizes, T));
    DIC_LAUNCH_CHECK();
  }

  // full_att
  DIC_TRY(launch_colsum(dwfull_part, 0, (int)TB, A, A, gr.full_att_w, st, true));
  DIC_TRY(launch_colsum(dbfull_part, 0, (int)TB, 1, 1, gr.full_att_b, st, true));

  if (cudaEvent_t ev = g_grads_mid_event.exchange(nullptr)) DIC_CUDA(cudaEventRecord(ev, st));

  // encoder_att: datt1 summed over steps, then two contractions over B*L rows
  {
    Datt1Args da;
    da.att1 = att1; da.hp_all = HP; da.de_all = de; da.w_full = pk.w_full(); da.datt1 = datt1;
    da.B = B; da.L = L; da.D = D; da.A = A; da.T = T; da.sizes = sizes;
    DIC_TRY(launch_datt1<ST>(da, st));
    DIC_TRY(launch_colsum(datt1, is_bf16, B * L, A, A, gr.enc_att_b, st, true));
    DIC_TRY(wgrad(datt1, A, A, F, D, D, B * L, gr.enc_att_w, D));
  }

  const cudaEvent_t ev_all = g_grads_ready_event.exchange(nullptr);     // one shot
  if (ev_all) DIC_CUDA(cudaEventRecord(ev_all, st));

  // dL/dF = datt1 . W_enc + sum_t alpha_t (x) dz_t + dmeanF / L
  // (written in the annotations' dtype: fp32 in place, bf16 through the fp32 accumulation buffer dF32)
  if (d_feats && is_bf16 && dfeat_bf16 && dfeat_tc_eligible(A, D, lay.Lp)) {
    // one tensor-core GEMM over K = A + T, bf16 dF written once (dfeat_tc.cuh)
    DIC_TRY(launch_dfeat_tc(reinterpret_cast<const bf16*>(datt1), reinterpret_cast<const bf16*>(pk.Wenc()),
                            reinterpret_cast<const bf16*>(ws + lay.alpha16), lay.Lp,
                            reinterpret_cast<const bf16*>(DZ), dmeanF, reinterpret_cast<bf16*>(d_feats), B, L, D, A,
                            T, st, ev_all ? kAllReduceSms : 0));
  } else if (d_feats) {
    float* acc = dfeat_bf16 ? reinterpret_cast<float*>(ws + lay.dF32) : reinterpret_cast<float*>(d_feats);
    GemmArgs g = gemm_args_nt(datt1, is_bf16, A, pk.Wenc(), is_bf16, 0, acc, 0, D, B * L, D, A, nullptr);
    g.b_n = 1; g.b_k = D;
    DIC_TRY(gemm(g, st));
    DIC_TRY(launch_dfeat_accumulate<ST>(acc, alphas, DZ, dmeanF, B, L, D, T, 1,
                                        dfeat_bf16 ? reinterpret_cast<bf16*>(d_feats) : nullptr, st));
  }
  return 0;
}

// =============================================================================================
// decoding (greedy and beam share the step pipeline; beam = rows_per_image > 1 + selection)
// =============================================================================================
// DIC_BEAM_LOOKAHEAD=0: serial order (A/B switch of the measurement and of the parity test)
static bool lookahead_enabled() {
  const char* e = getenv("DIC_BEAM_LOOKAHEAD");
  return !(e && e[0] == '0') && !g_prof.on;      // per-kernel event timing wants kernels timed alone
}

template <typename ST>
static int decode_impl(const dic_dims& d, int attn_mode, const void* pack, const void* f_rgb,
                       const void* f_depth, int feat_dtype, int B, int K, bool beam, int start_id,
                       int end_id, int max_len, const float* u, int64_t* tokens, int32_t* lengths,
                       float* scores_out, float* alphas_out, float* logits_out, int32_t* back_out,
                       int32_t* toks_out, float* step_scores_out, float* lse_out, char* ws,
                       cudaStream_t st) {
  const int dtype = sizeof(ST) == 2 ? DIC_BF16 : DIC_F32;
  const int is_bf16 = sizeof(ST) == 2;
  const Pack pk(pack, d, dtype);
  const DecodeLayout lay(d, dtype, B, K);
  const long long XW = (long long)lay.XW;
  const int R = B * K;
  const int H = d.H, A = d.A, D = d.D, E = d.E, L = d.L, V = d.V;
  ST* Fsum = reinterpret_cast<ST*>(ws + lay.Fsum);
  float* meanF = reinterpret_cast<float*>(ws + lay.meanF);
  ST* att1 = reinterpret_cast<ST*>(ws + lay.att1);
  ST* XH = reinterpret_cast<ST*>(ws + lay.XH);
  float* HP = reinterpret_cast<float*>(ws + lay.HP);
  float* c = reinterpret_cast<float*>(ws + lay.c);
  float* c_tmp = reinterpret_cast<float*>(ws + lay.c_tmp);
  ST* h_tmp = reinterpret_cast<ST*>(ws + lay.h_tmp);
  float* h0 = reinterpret_cast<float*>(ws + lay.h0);
  float* c0 = reinterpret_cast<float*>(ws + lay.c0);
  float* gate_part = reinterpret_cast<float*>(ws + lay.gate_part);
  float* logits_ws = reinterpret_cast<float*>(ws + lay.logits);
  float* lse_ws = reinterpret_cast<float*>(ws + lay.lse);
  float* sc[2] = {reinterpret_cast<float*>(ws + lay.scores), reinterpret_cast<float*>(ws + lay.scores2)};
  uint8_t* fin[2] = {reinterpret_cast<uint8_t*>(ws + lay.fin), reinterpret_cast<uint8_t*>(ws + lay.fin2)};
  int32_t* back_ws = reinterpret_cast<int32_t*>(ws + lay.back);
  int32_t* tok_ws = reinterpret_cast<int32_t*>(ws + lay.tok);

  const ST* F = nullptr;
  bf16* mean16 = is_bf16 ? reinterpret_cast<bf16*>(ws + lay.meanF16) : nullptr;
  DIC_TRY(prologue<ST>(d, pk, f_rgb, f_depth, feat_dtype, B, Fsum, meanF, mean16, att1, &F, st));
  int hc_stride = H;
  if (is_bf16) {
    float* hc0 = reinterpret_cast<float*>(ws + lay.hc0);
    DIC_TRY(init_state_gemm_tc(d, pk, mean16, B, hc0, st));
    h0 = hc0; c0 = hc0 + H; hc_stride = 2 * H;
  } else {
    DIC_TRY(init_state_gemm<ST>(d, pk, meanF, B, h0, c0, st));
  }
  DIC_CUDA(cudaMemsetAsync(XH, 0, (size_t)2 * R * XW * sizeof(ST), st));
  decode_init_kernel<ST><<<cdiv(R * (E + H), 256), 256, 0, st>>>(
      h0, c0, hc_stride, reinterpret_cast<const ST*>(pk.Emb()), start_id, XH, XW, E + D, c, R, K, E, H);
  DIC_LAUNCH_CHECK();
  if (beam) {
    beam_state_init_kernel<<<cdiv(R, 256), 256, 0, st>>>(sc[0], fin[0], B, K);
    DIC_LAUNCH_CHECK();
  }

  // time loop over sub-batches of images on their own streams (common.cuh, "sub-batch streams"):
  // a decode step is ~9 short dependent kernels, so the sub-batches' chains overlap almost freely
  int i0s[kMaxSub + 1];
  const int S = sub_bounds(B, pick_substreams(B, 0), 4, i0s);
  cudaStream_t ss[kMaxSub];
  DIC_TRY(sub_fork(st, S, ss));
  // Look-ahead attention (fused bf16 beam step, one stream of images).  The attention of step t+1 needs h_t of
  // the row's PARENT only, and the beam selection merely permutes / duplicates the rows of an image.  So from the
  // first selection on nothing is reordered any more: every per-row tensor stays in parent order,
  //     ZH[t] = [ beta.z of step t+1 | h_t ]   (context kernel | LSTM kernel write their column ranges)
  // the gate GEMM of step t+1 runs on it as P = ZH[t] . [W_z | W_hh]^T, and the LSTM kernel of step t+1 follows the
  // backpointers: gates[r] = P[parent(r)] + etab[token(r)] + bias, c_prev = c[parent(r)] (decode.cuh
  // lstm_beam_kernel; etab = Emb . W_e^T once per call).  Launch order per step, ONE stream, programmatic launches:
  //     lstm(t) -> logits+stats(t) -> head(t+1) -> select(t) -> context(t+1) -> gate GEMM(t+1) -> lstm(t+1)
  // The context kernel (HBM bound, small CTAs) does not depend on the selection kernel launched right before it
  // (latency bound, one light CTA per image): it skips its dependency wait and the two run CONCURRENTLY.  Its inputs
  // come from the head kernel, which is complete by then because the selection kernel lets its dependents start only
  // after its own wait has returned; ONE context CTA waits at its very end instead (AttnFwdArgs.late_wait), so the
  // gate GEMM behind it (and the LSTM kernel behind that) still see the selection's backpointers.  Same arithmetic per
  // row as the serial order up to the place where the token's share of the gates is added (outside the GEMM here).
  // (Measured on the way, profiles/r02_beam_lookahead.txt: a two-stream version with events -- 5 us per cross-stream
  // hand-over, and the whole-SM logits CTAs starved behind the context grid; every context CTA waiting at its end --
  // the first wave keeps its slots; a gather kernel copying the parents' contexts into child order -- 3.8 us.)
  const bool fused_beam = beam && is_bf16 && !logits_out && !lse_out && beam_fused_eligible(H, V, K, R);
  const bool lookahead = fused_beam && S == 1 && !alphas_out && !u && attn_head_eligible(A, H, D, L, K) &&
                         D % 8 == 0 && lookahead_enabled();
  // Greedy decoding, same idea without the reorder: the head kernel of step t+1 follows the logits GEMM of step t, and
  // the context pass of step t+1 runs next to the arg-max + embedding kernel of step t (both write their own columns
  // of the next step's input rows):   gates -> lstm -> logits -> head(t+1) -> argmax_embed(t) -> context(t+1)
  const bool lookahead_g = !beam && is_bf16 && S == 1 && attn_mode == DIC_ATTN_SOFT && !u &&
                           attn_head_eligible(A, H, D, L, K) && lookahead_enabled();
  const long long ZW = (long long)D + H;
  ST* ZH = reinterpret_cast<ST*>(ws + lay.ZH);               // [2][R][D + H], parent order
  float* c_par = reinterpret_cast<float*>(ws + lay.c_par);   // [2][R][H]
  float* etab = reinterpret_cast<float*>(ws + lay.etab);
  if (lookahead) {
    GemmArgs g = gemm_args_nt(pk.Emb(), 1, E, pk.Wg(), 1, XW, etab, 0, 4 * H, V, 4 * H, E, nullptr);
    DIC_TRY(gemm(g, st));
  }
  for (int t = 0; t < max_len; ++t) {
    for (int sb = 0; sb < S; ++sb) {
      const int i0 = i0s[sb], Bs = i0s[sb + 1] - i0;      // images of this sub-batch
      const size_t r0 = (size_t)i0 * K;                    // first row
      const int Rs = Bs * K;
      cudaStream_t sst = ss[sb];
      ST* X = XH + ((size_t)(t & 1) * R + r0) * XW;
      ST* Xn = XH + ((size_t)((t + 1) & 1) * R + r0) * XW;
      float* HPs = HP + r0 * (A + D);
      const bool head = is_bf16 && attn_head_eligible(A, H, D, L, K);
      // attention of the rows whose hidden state is h (row stride h_ld); gated contexts to zg (row stride zg_ld)
      // phase 0 = head + context, 1 = head only, 2 = context only (concurrent with its predecessor: late_wait)
      auto attention = [&](const ST* h, long long h_ld, ST* zg, long long zg_ld, int step, int phase, cudaStream_t s_) -> int {
        if (!head && phase != 2) DIC_TRY(hproj<ST>(d, pk, h, h_ld, Rs, HPs, s_));
        AttnFwdArgs a;
        memset(&a, 0, sizeof(a));
        a.F = F + (size_t)i0 * L * D; a.att1 = att1 + (size_t)i0 * L * A; a.hp = HPs;
        a.w_full = pk.w_full(); a.b_full = pk.b_full();
        a.u = u ? u + ((size_t)step * R + r0) * L : nullptr;
        a.alpha_out = alphas_out ? alphas_out + ((size_t)step * R + r0) * L
                                 : reinterpret_cast<float*>(ws + lay.alpha) + r0 * L;
        a.alpha_stride = L;
        a.z_out = nullptr;
        a.zg_out = zg; a.zg_stride = zg_ld;
        a.L = L; a.D = D; a.A = A; a.mode = attn_mode; a.inv_temp = 1.f;
        if (head) {
          if (phase != 2) {
            HeadArgs hd;
            hd.h = reinterpret_cast<const bf16*>(h); hd.h_ld = h_ld;
            hd.Wdb = reinterpret_cast<const bf16*>(pk.Wdb(d)); hd.bias_db = pk.bias_db();
            hd.HP = HPs; hd.a = a; hd.rows = Rs;
            DIC_TRY(launch_attn_head(hd, K, s_));
          }
          a.skip_alpha = 1;
        }
        if (phase == 1) return 0;
        a.late_wait = phase == 2;
        return launch_attn_step<ST>(a, Bs, K, s_);
      };
      // (look-ahead: the contexts of step t > 0 were computed in the previous iteration)
      if (!(lookahead || lookahead_g) || t == 0) DIC_TRY(attention(X + E + D, XW, X + E, XW, t, 0, sst));

      float* gp = gate_part + r0 * 4 * H;
      // look-ahead beam: parent-order buffers of this step (h_t, c_t) and the previous one
      ST* ZHt = ZH + ((size_t)(t & 1) * R + r0) * ZW;
      float* cpt = c_par + ((size_t)(t & 1) * R + r0) * H;
      if (lookahead && t > 0) {
        const ST* ZHp = ZH + ((size_t)((t - 1) & 1) * R + r0) * ZW;
        int splits = 1;
        DIC_TRY(gates_gemm_from<ST>(d, ZHp, ZW, reinterpret_cast<const ST*>(pk.Wg()) + E, XW, (int)ZW, Rs, R, gp, &splits, sst));
        LstmBeamArgs lb;
        memset(&lb, 0, sizeof(lb));
        lb.gate_part = gp; lb.part_stride = (long long)R * 4 * H; lb.splits = splits;
        lb.etab = etab; lb.bias_g = pk.bias_g();
        lb.back = back_ws + (size_t)(t - 1) * R + r0; lb.tok = tok_ws + (size_t)(t - 1) * R + r0;
        lb.c_par = c_par + ((size_t)((t - 1) & 1) * R + r0) * H; lb.c_out = cpt;
        lb.h_out = ZHt + D; lb.h_stride = ZW;
        lb.rows = Rs; lb.H = H; lb.K = K; lb.trace = g_trace_host;
        {
          ProfScope prof(P_LSTM, sst);
          DIC_CUDA(launch_pdl(lstm_beam_kernel<ST>, dim3(cdiv(Rs * H, 256)), dim3(256), 0, sst, lb));
          DIC_LAUNCH_CHECK();
        }
      } else {
        LstmFwdArgs l;
        memset(&l, 0, sizeof(l));
        l.bias_g = pk.bias_g();
        l.c_in = c + r0 * H;
        l.c_out = lookahead ? cpt : (beam ? c_tmp + r0 * H : c + r0 * H);
        l.h_out = lookahead ? ZHt + D : (beam ? h_tmp + r0 * H : Xn + E + D);
        l.h_stride = lookahead ? ZW : (beam ? H : XW);
        l.rows = Rs; l.H = H;
        DIC_TRY(lstm_step<ST>(d, pk, X, XW, Rs, R, gp, l, sst));
      }

      float* lg = logits_out ? logits_out + ((size_t)t * R + r0) * V : logits_ws + r0 * V;
      if (fused_beam) {
        // vocabulary projection + log-sum-exp + per-row top-K in one tcgen05 kernel, merge + reorder in one more
        // (beam_fused.cuh)
        int32_t* back_t = back_ws + (size_t)t * R + r0;
        int32_t* tok_t = tok_ws + (size_t)t * R + r0;
        BeamMergeArgs mg;
        memset(&mg, 0, sizeof(mg));
        mg.scores = sc[t & 1] + r0; mg.finished = fin[t & 1] + r0;
        mg.new_scores = sc[(t + 1) & 1] + r0; mg.back = back_t; mg.tok = tok_t; mg.new_finished = fin[(t + 1) & 1] + r0;
        mg.h_tmp = h_tmp + r0 * H; mg.c_tmp = c_tmp + r0 * H; mg.emb = pk.Emb();
        mg.Xnext = Xn; mg.x_row = XW; mg.col_h = E + D; mg.c = c + r0 * H;
        mg.end_id = end_id; mg.E = E; mg.H = H;
        float* bstats = reinterpret_cast<float*>(ws + lay.bstats) + r0 * 2 * cdiv(V, kBfNB);
        if (lookahead) {
          mg.Xnext = nullptr;                    // nothing is reordered: the next LSTM kernel follows the backpointers
          const bool more = t + 1 < max_len;
          DIC_TRY(launch_beam_logits_stats(ZHt + D, ZW, pk.Wout(), pk.b_out(), lg, bstats, Rs, V, sst));
          if (more) DIC_TRY(attention(ZHt + D, ZW, ZHt, ZW, t + 1, 1, sst));
          DIC_TRY(launch_beam_select_reorder<ST>(lg, bstats, mg, Rs, V, K, sst));
          if (more) DIC_TRY(attention(ZHt + D, ZW, ZHt, ZW, t + 1, 2, sst));
        } else {
          DIC_TRY(launch_beam_fused<ST>(h_tmp + r0 * H, pk.Wout(), pk.b_out(), lg, bstats, mg, Rs, V, K, sst));
        }
        if (step_scores_out)
          DIC_CUDA(cudaMemcpyAsync(step_scores_out + (size_t)t * R + r0, sc[(t + 1) & 1] + r0, sizeof(float) * Rs,
                                   cudaMemcpyDeviceToDevice, sst));
        continue;
      }
      const ST* hsrc = beam ? h_tmp + r0 * H : Xn + E + D;
      GemmArgs g = gemm_args_nt(hsrc, is_bf16, beam ? H : XW, pk.Wout(), is_bf16, H, lg, 0, V, Rs, V, H, pk.b_out());
      g.tag = 5;
      g.b_static = 1;
      DIC_TRY(gemm(g, sst));

      if (!beam) {
        const bool more = lookahead_g && t + 1 < max_len;
        if (more) DIC_TRY(attention(Xn + E + D, XW, Xn + E, XW, t + 1, 1, sst));
        DIC_CUDA(launch_pdl(argmax_embed_kernel<ST>, dim3(Rs), dim3(256), 0, sst, (const float*)lg, V,
                            tokens + r0 * max_len + t, (long long)max_len,
                            reinterpret_cast<const ST*>(pk.Emb()), E, Xn, XW, g_trace_host));
        DIC_LAUNCH_CHECK();
        if (more) DIC_TRY(attention(Xn + E + D, XW, Xn + E, XW, t + 1, 2, sst));
      } else {
        float* lse_t = lse_out ? lse_out + (size_t)t * R + r0 : lse_ws + r0;
        int32_t* back_t = back_ws + (size_t)t * R + r0;
        int32_t* tok_t = tok_ws + (size_t)t * R + r0;
        // row log-sum-exp + per-row top-K in one kernel, then a per-image merge
        DIC_TRY(launch_beam_select(sc[t & 1] + r0, fin[t & 1] + r0, lg, nullptr, lse_t, Bs, K, V, end_id,
                                   ws + lay.cand + r0 * K * (sizeof(float) + sizeof(int)),
                                   sc[(t + 1) & 1] + r0, back_t, tok_t, fin[(t + 1) & 1] + r0, sst, is_bf16 != 0));
        DIC_CUDA(launch_pdl(beam_reorder_kernel<ST>, dim3(cdiv(Rs * (E + H), 256)), dim3(256), 0, sst,
                            (const ST*)(h_tmp + r0 * H), (const float*)(c_tmp + r0 * H), (const int32_t*)back_t,
                            (const int32_t*)tok_t, reinterpret_cast<const ST*>(pk.Emb()), Xn, XW, E + D,
                            c + r0 * H, Rs, K, E, H, g_trace_host));
        DIC_LAUNCH_CHECK();
        if (step_scores_out)
          DIC_CUDA(cudaMemcpyAsync(step_scores_out + (size_t)t * R + r0, sc[(t + 1) & 1] + r0, sizeof(float) * Rs,
                                   cudaMemcpyDeviceToDevice, sst));
      }
    }
  }
  DIC_TRY(sub_join(st, S));
  if (beam) {
    beam_backtrack_kernel<<<cdiv(B, 128), 128, 0, st>>>(back_ws, tok_ws, sc[max_len & 1], B, K, max_len,
                                                        end_id, tokens, lengths, scores_out);
    DIC_LAUNCH_CHECK();
    if (back_out)
      DIC_CUDA(cudaMemcpyAsync(back_out, back_ws, sizeof(int32_t) * (size_t)max_len * R,
                               cudaMemcpyDeviceToDevice, st));
    if (toks_out)
      DIC_CUDA(cudaMemcpyAsync(toks_out, tok_ws, sizeof(int32_t) * (size_t)max_len * R,
                               cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

}  // namespace dic

using namespace dic;

// =============================================================================================
// extern "C"
// =============================================================================================

// =============================================================================================
// depth CNN encoder (depth_encoder.cuh)
// =============================================================================================
static int enc_check(int B, int Hi, int Wi, int dtype) {
  if (dtype != DIC_F32 && dtype != DIC_BF16) DIC_FAIL(-2, "depth encoder: dtype must be DIC_F32 or DIC_BF16");
  if (B < 1 || Hi < 7 || Wi < 7) DIC_FAIL(-2, "depth encoder: bad input shape [%d, %d, %d]", B, Hi, Wi);
  const EncGeom g(B, Hi, Wi);
  if (g.P2h != 7 || g.P2w != 7)
    DIC_FAIL(-2, "depth encoder: the last feature map is %d x %d; AdaptiveAvgPool2d(14) is built as the exact 2 x 2 "
             "replication of a 7 x 7 map (224 x 224 depth images, depth_models.py:18-32)", g.P2h, g.P2w);
  return 0;
}

template <typename ST>
static int enc_bn_stage(const float* X, size_t M, int C, int slot, char* ws, const EncLayout& lay, int training,
                        float momentum, float eps, float* rmean, float* rvar, const float** mean_o, const float** invstd_o,
                        cudaStream_t st) {
  double* sum = reinterpret_cast<double*>(ws + lay.stats) + (size_t)slot * 2 * 2048;
  double* sumsq = sum + 2048;
  float* mean = reinterpret_cast<float*>(ws + lay.mean) + (size_t)slot * 2048;
  float* invstd = reinterpret_cast<float*>(ws + lay.invstd) + (size_t)slot * 2048;
  if (training) {
    DIC_CUDA(cudaMemsetAsync(sum, 0, sizeof(double) * 2 * 2048, st));
    const int rpb = 256;
    dim3 grid(cdiv(C, 256), (unsigned)((M + rpb - 1) / rpb));
    enc_stats_kernel<<<grid, 256, 0, st>>>(X, M, C, rpb, sum, sumsq);
    DIC_LAUNCH_CHECK();
  }
  enc_bn_finalize_kernel<<<cdiv(C, 256), 256, 0, st>>>(sum, sumsq, (double)M, C, training, momentum, eps, rmean, rvar,
                                                       mean, invstd);
  DIC_LAUNCH_CHECK();
  *mean_o = mean; *invstd_o = invstd;
  return 0;
}

template <typename ST>
static int depth_encoder_forward_impl(int training, const EncGeom& g, const float* imgs, const dic_enc_params& p,
                                      float momentum, float eps, void* feats, int feat_dtype, char* ws, cudaStream_t st) {
  const int dtype = sizeof(ST) == 2 ? DIC_BF16 : DIC_F32;
  const int is_bf16 = sizeof(ST) == 2;
  const EncLayout lay(g, dtype);
  ST* col1 = reinterpret_cast<ST*>(ws + lay.col1);
  float* out1 = reinterpret_cast<float*>(ws + lay.out1);
  ST* p1 = reinterpret_cast<ST*>(ws + lay.p1);
  ST* col2 = reinterpret_cast<ST*>(ws + lay.col2);
  float* out2 = reinterpret_cast<float*>(ws + lay.out2);
  ST* p2 = reinterpret_cast<ST*>(ws + lay.p2);
  float* out3 = reinterpret_cast<float*>(ws + lay.out3);
  ST* w1p = reinterpret_cast<ST*>(ws + lay.w1p);
  ST* w2p = reinterpret_cast<ST*>(ws + lay.w2p);
  ST* w3p = reinterpret_cast<ST*>(ws + lay.w3p);
  const int K2 = 9 * g.C1;
  enc_pack_w_kernel<ST><<<enc_grid((size_t)g.C1 * g.K1p), 256, 0, st>>>(p.conv1_w, w1p, g.C1, 1, 49, g.K1p);
  enc_pack_w_kernel<ST><<<enc_grid((size_t)g.C2 * K2), 256, 0, st>>>(p.conv2_w, w2p, g.C2, g.C1, 9, K2);
  enc_pack_w_kernel<ST><<<enc_grid((size_t)g.C3 * g.C2), 256, 0, st>>>(p.conv3_w, w3p, g.C3, g.C2, 1, g.C2);
  DIC_LAUNCH_CHECK();
  const float *mean, *invstd;
  // stage 1
  enc_im2col_kernel<float, ST><<<enc_grid(g.M1() * g.K1p), 256, 0, st>>>(imgs, col1, g.B, g.Hi, g.Wi, 1, 7, 3, g.H1, g.W1, g.K1p);
  DIC_LAUNCH_CHECK();
  DIC_TRY(gemm(gemm_args_nt(col1, is_bf16, g.K1p, w1p, is_bf16, g.K1p, out1, 0, g.C1, (int)g.M1(), g.C1, g.K1p, p.conv1_b), st));
  DIC_TRY(enc_bn_stage<ST>(out1, g.M1(), g.C1, 0, ws, lay, training, momentum, eps, p.bn1_mean, p.bn1_var, &mean, &invstd, st));
  enc_bn_relu_pool_kernel<ST><<<enc_grid((size_t)g.B * g.P1h * g.P1w * g.C1), 256, 0, st>>>(out1, mean, invstd, p.bn1_w, p.bn1_b, p1,
                                                                                          g.B, g.H1, g.W1, g.C1, 3, 1);
  DIC_LAUNCH_CHECK();
  // stage 2
  enc_im2col_kernel<ST, ST><<<enc_grid(g.M2() * K2), 256, 0, st>>>(p1, col2, g.B, g.P1h, g.P1w, g.C1, 3, 1, g.H2, g.W2, K2);
  DIC_LAUNCH_CHECK();
  DIC_TRY(gemm(gemm_args_nt(col2, is_bf16, K2, w2p, is_bf16, K2, out2, 0, g.C2, (int)g.M2(), g.C2, K2, p.conv2_b), st));
  DIC_TRY(enc_bn_stage<ST>(out2, g.M2(), g.C2, 1, ws, lay, training, momentum, eps, p.bn2_mean, p.bn2_var, &mean, &invstd, st));
  enc_bn_relu_pool_kernel<ST><<<enc_grid(g.M3() * g.C2), 256, 0, st>>>(out2, mean, invstd, p.bn2_w, p.bn2_b, p2, g.B, g.H2, g.W2,
                                                                       g.C2, 3, 1);
  DIC_LAUNCH_CHECK();
  // stage 3: 1 x 1 convolution, then BN + ReLU + 2 x 2 replication straight into the annotation tensor
  DIC_TRY(gemm(gemm_args_nt(p2, is_bf16, g.C2, w3p, is_bf16, g.C2, out3, 0, g.C3, (int)g.M3(), g.C3, g.C2, p.conv3_b), st));
  DIC_TRY(enc_bn_stage<ST>(out3, g.M3(), g.C3, 2, ws, lay, training, momentum, eps, p.bn3_mean, p.bn3_var, &mean, &invstd, st));
  if (feat_dtype == DIC_BF16)
    enc_bn_relu_pool_kernel<bf16><<<enc_grid(g.M3() * g.C3), 256, 0, st>>>(out3, mean, invstd, p.bn3_w, p.bn3_b,
                                                                           reinterpret_cast<bf16*>(feats), g.B, g.P2h, g.P2w,
                                                                           g.C3, 1, 2);
  else
    enc_bn_relu_pool_kernel<float><<<enc_grid(g.M3() * g.C3), 256, 0, st>>>(out3, mean, invstd, p.bn3_w, p.bn3_b,
                                                                            reinterpret_cast<float*>(feats), g.B, g.P2h, g.P2w,
                                                                            g.C3, 1, 2);
  DIC_LAUNCH_CHECK();
  return 0;
}

template <typename ST>
static int depth_encoder_backward_impl(const EncGeom& g, const dic_enc_params& p, const void* d_feats, int feat_dtype,
                                       const dic_enc_params& gr, char* ws, cudaStream_t st) {
  const int dtype = sizeof(ST) == 2 ? DIC_BF16 : DIC_F32;
  const int is_bf16 = sizeof(ST) == 2;
  const EncLayout lay(g, dtype);
  ST* col1 = reinterpret_cast<ST*>(ws + lay.col1);
  float* out1 = reinterpret_cast<float*>(ws + lay.out1);
  ST* col2 = reinterpret_cast<ST*>(ws + lay.col2);
  float* out2 = reinterpret_cast<float*>(ws + lay.out2);
  ST* p2 = reinterpret_cast<ST*>(ws + lay.p2);
  float* out3 = reinterpret_cast<float*>(ws + lay.out3);
  ST* w2p = reinterpret_cast<ST*>(ws + lay.w2p);
  ST* w3p = reinterpret_cast<ST*>(ws + lay.w3p);
  float* dy3 = reinterpret_cast<float*>(ws + lay.dy3);
  float* dy2 = reinterpret_cast<float*>(ws + lay.dy2);
  float* dy1 = reinterpret_cast<float*>(ws + lay.dy1);
  bf16* dx16 = reinterpret_cast<bf16*>(ws + lay.dx16);
  float* dp2 = reinterpret_cast<float*>(ws + lay.dp2);
  float* dcol2 = reinterpret_cast<float*>(ws + lay.dcol2);
  float* dp1 = reinterpret_cast<float*>(ws + lay.dp1);
  float* dwp = reinterpret_cast<float*>(ws + lay.dwp);
  double* stats = reinterpret_cast<double*>(ws + lay.stats);
  const float* mean = reinterpret_cast<const float*>(ws + lay.mean);
  const float* invstd = reinterpret_cast<const float*>(ws + lay.invstd);
  const int K2 = 9 * g.C1;
  DIC_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 3 * 2 * 2048, st));

  // one stage: route dY through replication / max-pool / ReLU, reduce, apply the BN backward
  auto stage = [&](auto gt_tag, const void* dY, const float* X, float* dy, int slot, int H, int W, int C, int P, int R,
                   const float* gamma, const float* beta, float* dgamma, float* dbeta) -> int {
    using GT = decltype(gt_tag);
    const size_t M = (size_t)g.B * H * W;
    if (H % P || W % P) DIC_CUDA(cudaMemsetAsync(dy, 0, sizeof(float) * M * C, st));
    double* s_dy = stats + (size_t)slot * 2 * 2048;
    double* s_dyx = s_dy + 2048;
    const size_t nwin = (size_t)g.B * (H / P) * (W / P);
    const int wpb = 64;
    dim3 grid(cdiv(C, 256), (unsigned)((nwin + wpb - 1) / wpb));
    enc_bwd_route_kernel<GT><<<grid, 256, 0, st>>>(reinterpret_cast<const GT*>(dY), X, mean + slot * 2048, invstd + slot * 2048,
                                                   gamma, beta, dy, g.B, H, W, C, P, R, wpb, s_dy, s_dyx);
    DIC_LAUNCH_CHECK();
    enc_bwd_apply_kernel<bf16><<<enc_grid(M * C), 256, 0, st>>>(dy, X, mean + slot * 2048, invstd + slot * 2048, gamma, s_dy, s_dyx,
                                                                (double)M, M, C, is_bf16 ? dx16 : nullptr, dgamma, dbeta);
    DIC_LAUNCH_CHECK();
    return 0;
  };
  // dW = dx^T . rows  (split-K), db = colsum(dx), through the mode's GEMM engine
  auto wgrad = [&](const float* dx, int C, const ST* rows, int Kp, size_t M, float* dW) -> int {
    const void* a = is_bf16 ? (const void*)dx16 : (const void*)dx;
    GemmArgs ga = gemm_args_nt(a, is_bf16, 0, rows, is_bf16, 0, dW, 0, Kp, C, Kp, (int)M, nullptr);
    ga.a_m = 1; ga.a_k = C; ga.b_n = 1; ga.b_k = Kp;
    return gemm_splitk(ga, st);
  };

  // stage 3
  if (feat_dtype == DIC_BF16)
    DIC_TRY(stage(bf16{}, d_feats, out3, dy3, 2, g.P2h, g.P2w, g.C3, 1, 2, p.bn3_w, p.bn3_b, gr.bn3_w, gr.bn3_b));
  else
    DIC_TRY(stage(float{}, d_feats, out3, dy3, 2, g.P2h, g.P2w, g.C3, 1, 2, p.bn3_w, p.bn3_b, gr.bn3_w, gr.bn3_b));
  DIC_TRY(launch_colsum(dy3, 0, (int)g.M3(), g.C3, g.C3, gr.conv3_b, st));
  DIC_TRY(wgrad(dy3, g.C3, p2, g.C2, g.M3(), gr.conv3_w));
  {
    const void* a = is_bf16 ? (const void*)dx16 : (const void*)dy3;
    GemmArgs ga = gemm_args_nt(a, is_bf16, g.C3, w3p, is_bf16, 0, dp2, 0, g.C2, (int)g.M3(), g.C2, g.C3, nullptr);
    ga.b_n = 1; ga.b_k = g.C2;
    DIC_TRY(gemm(ga, st));
  }
  // stage 2
  DIC_TRY(stage(float{}, dp2, out2, dy2, 1, g.H2, g.W2, g.C2, 3, 1, p.bn2_w, p.bn2_b, gr.bn2_w, gr.bn2_b));
  DIC_TRY(launch_colsum(dy2, 0, (int)g.M2(), g.C2, g.C2, gr.conv2_b, st));
  DIC_TRY(wgrad(dy2, g.C2, col2, K2, g.M2(), dwp));
  enc_unpack_dw_kernel<<<enc_grid((size_t)g.C2 * K2), 256, 0, st>>>(dwp, gr.conv2_w, g.C2, g.C1, 9, K2);
  DIC_LAUNCH_CHECK();
  {
    const void* a = is_bf16 ? (const void*)dx16 : (const void*)dy2;
    GemmArgs ga = gemm_args_nt(a, is_bf16, g.C2, w2p, is_bf16, 0, dcol2, 0, K2, (int)g.M2(), K2, g.C2, nullptr);
    ga.b_n = 1; ga.b_k = K2;
    DIC_TRY(gemm(ga, st));
  }
  enc_col2im3_kernel<<<enc_grid((size_t)g.B * g.P1h * g.P1w * g.C1), 256, 0, st>>>(dcol2, dp1, g.B, g.P1h, g.P1w, g.C1, g.H2, g.W2);
  DIC_LAUNCH_CHECK();
  // stage 1
  DIC_TRY(stage(float{}, dp1, out1, dy1, 0, g.H1, g.W1, g.C1, 3, 1, p.bn1_w, p.bn1_b, gr.bn1_w, gr.bn1_b));
  DIC_TRY(launch_colsum(dy1, 0, (int)g.M1(), g.C1, g.C1, gr.conv1_b, st));
  DIC_TRY(wgrad(dy1, g.C1, col1, g.K1p, g.M1(), dwp));
  enc_unpack_dw_kernel<<<enc_grid((size_t)g.C1 * 49), 256, 0, st>>>(dwp, gr.conv1_w, g.C1, 1, 49, g.K1p);
  DIC_LAUNCH_CHECK();
  return 0;
}


extern "C" {

int dic_version(void) { return DIC_VERSION; }

long long dic_launch_count(void) { return g_launches.load(); }
int dic_profile_classes(void) { return P_N; }
const char* dic_profile_class_name(int c) { return prof_class_name(c); }
void dic_profile_enable(int on) {
  g_prof.on = on != 0;
  for (int c = 0; c < P_N; ++c) { g_prof.used[c] = 0; g_prof.bytes[c] = 0; }
}
// Sums the CUDA-event durations recorded since dic_profile_enable(1) (synchronises on the
// recorded events) and resets the counters.  ms/launches/bytes: arrays of dic_profile_classes().
int dic_profile_read(float* ms, long long* launches, double* bytes) {
  for (int c = 0; c < P_N; ++c) {
    float tot = 0.f;
    for (size_t i = 0; i + 1 < g_prof.used[c]; i += 2) {
      float t = 0.f;
      DIC_CUDA(cudaEventSynchronize(g_prof.ev[c][i + 1]));
      DIC_CUDA(cudaEventElapsedTime(&t, g_prof.ev[c][i], g_prof.ev[c][i + 1]));
      tot += t;
    }
    ms[c] = tot;
    launches[c] = (long long)(g_prof.used[c] / 2);
    bytes[c] = g_prof.bytes[c];
    g_prof.used[c] = 0;
    g_prof.bytes[c] = 0;
  }
  return 0;
}
const char* dic_last_error(void) { return g_err; }

int dic_trace_start(void* buf, unsigned int capacity_records) {
  TraceRec* b = reinterpret_cast<TraceRec*>(buf);
  unsigned int zero = 0;
  DIC_CUDA(cudaDeviceSynchronize());
  DIC_CUDA(cudaMemcpyToSymbol(g_trace_cnt, &zero, sizeof(zero)));
  DIC_CUDA(cudaMemcpyToSymbol(g_trace_cap, &capacity_records, sizeof(capacity_records)));
  g_trace_host = b;
  return 0;
}
int dic_trace_stop(unsigned int* count) {
  DIC_CUDA(cudaDeviceSynchronize());
  g_trace_host = nullptr;
  if (count) DIC_CUDA(cudaMemcpyFromSymbol(count, g_trace_cnt, sizeof(unsigned int)));
  return 0;
}

void dic_set_grads_ready_event(void* event) { g_grads_ready_event.store(reinterpret_cast<cudaEvent_t>(event)); }
void dic_set_grads_ready_events(void* ev_linear, void* ev_middle, void* ev_all) {
  g_grads_lin_event.store(reinterpret_cast<cudaEvent_t>(ev_linear));
  g_grads_mid_event.store(reinterpret_cast<cudaEvent_t>(ev_middle));
  g_grads_ready_event.store(reinterpret_cast<cudaEvent_t>(ev_all));
}

size_t dic_depth_encoder_workspace_bytes(int B, int Hi, int Wi, int dtype) {
  if (enc_check(B, Hi, Wi, dtype)) return 0;
  return EncLayout(EncGeom(B, Hi, Wi), dtype).bytes;
}

int dic_depth_encoder_forward(int dtype, int training, int B, int Hi, int Wi, const float* depth_imgs,
                              const dic_enc_params* params, float momentum, float eps, void* feats, int feat_dtype,
                              void* workspace, size_t workspace_bytes, void* stream) {
  if (enc_check(B, Hi, Wi, dtype)) return -2;
  if (!depth_imgs || !params || !feats || !workspace) DIC_FAIL(-4, "depth encoder: null pointer");
  if (feat_dtype != DIC_F32 && feat_dtype != DIC_BF16) DIC_FAIL(-2, "depth encoder: bad feat_dtype");
  const EncGeom g(B, Hi, Wi);
  if (workspace_bytes < EncLayout(g, dtype).bytes) DIC_FAIL(-5, "depth encoder: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* ws = reinterpret_cast<char*>(workspace);
  if (dtype == DIC_BF16)
    return depth_encoder_forward_impl<bf16>(training, g, depth_imgs, *params, momentum, eps, feats, feat_dtype, ws, st);
  return depth_encoder_forward_impl<float>(training, g, depth_imgs, *params, momentum, eps, feats, feat_dtype, ws, st);
}

int dic_depth_encoder_backward(int dtype, int B, int Hi, int Wi, const dic_enc_params* params, const void* d_feats,
                               int feat_dtype, const dic_enc_params* grads, void* workspace, size_t workspace_bytes,
                               void* stream) {
  if (enc_check(B, Hi, Wi, dtype)) return -2;
  if (!params || !d_feats || !grads || !workspace) DIC_FAIL(-4, "depth encoder backward: null pointer");
  if (feat_dtype != DIC_F32 && feat_dtype != DIC_BF16) DIC_FAIL(-2, "depth encoder: bad feat_dtype");
  const EncGeom g(B, Hi, Wi);
  if (workspace_bytes < EncLayout(g, dtype).bytes) DIC_FAIL(-5, "depth encoder: workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* ws = reinterpret_cast<char*>(workspace);
  if (dtype == DIC_BF16) return depth_encoder_backward_impl<bf16>(g, *params, d_feats, feat_dtype, *grads, ws, st);
  return depth_encoder_backward_impl<float>(g, *params, d_feats, feat_dtype, *grads, ws, st);
}


size_t dic_dp_flag_bytes(void) { return kDpFlagWords * sizeof(uint32_t); }

int dic_dp_allreduce(int world, int rank, void* const* bufs, void* const* flags, void* multicast, long long n_floats,
                     float scale, unsigned int epoch, int blocks, void* stream) {
  if (!bufs || !flags) DIC_FAIL(-4, "dp_allreduce: null pointer table");
  if (world < 1 || world > kDpMaxRanks) DIC_FAIL(-4, "dp_allreduce: world size %d not in 1..%d", world, kDpMaxRanks);
  if (n_floats <= 0 || n_floats % (4LL * world)) DIC_FAIL(-4, "dp_allreduce: %lld floats are not a multiple of 4 x world", n_floats);
  if (epoch == 0) DIC_FAIL(-4, "dp_allreduce: epochs count from 1 (the flag words start at 0)");
  DpArgs p;
  memset(&p, 0, sizeof(p));
  for (int r = 0; r < world; ++r) {
    if (!bufs[r] || !flags[r]) DIC_FAIL(-4, "dp_allreduce: null buffer / flag pointer of rank %d", r);
    if (reinterpret_cast<uintptr_t>(bufs[r]) & 15) DIC_FAIL(-4, "dp_allreduce: buffer of rank %d is not 16-byte aligned", r);
    p.bufs[r] = reinterpret_cast<float*>(bufs[r]);
    p.flags[r] = reinterpret_cast<uint32_t*>(flags[r]);
  }
  p.mc = reinterpret_cast<float*>(multicast);
  p.rank = rank; p.world = world; p.n4 = n_floats / 4; p.scale = scale; p.epoch = epoch;
  return launch_dp_allreduce(p, blocks, reinterpret_cast<cudaStream_t>(stream));
}

void dic_set_substreams(int n) { g_sub_override.store(n < 0 ? 0 : (n > kMaxSub ? kMaxSub : n)); }

size_t dic_pack_bytes(const dic_dims* dims, int dtype) {
  if (check_dims(dims, dtype)) return 0;
  return PackLayout(*dims, dtype).bytes;
}

int dic_pack_weights(const dic_dims* dims, int dtype, const dic_params* p, void* pack, void* stream) {
  DIC_TRY(check_dims(dims, dtype));
  if (!p || !pack) DIC_FAIL(-1, "null argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const dic_dims& d = *dims;
  const PackLayout lay(d, dtype);
  char* base = reinterpret_cast<char*>(pack);
  const int bf = dtype == DIC_BF16;
  const size_t es = lay.es;
  const long long XW = (long long)d.E + d.D + d.H;
  PackJobs jobs;
  jobs.n = 0;
  auto add = [&](const float* src, long long src_ld, void* dst, long long dst_ld, int dst_bf16, int R, int C,
                 const float* src2 = nullptr) {
    PackJob& j = jobs.j[jobs.n++];
    j.src = src; j.src2 = src2; j.dst = dst; j.src_ld = src_ld; j.dst_ld = dst_ld; j.R = R; j.C = C;
    j.dst_bf16 = dst_bf16;
    j.gate_H = 0;
    j.gate_U = 16;
  };
  add(p->enc_att_w, d.D, base + lay.Wenc, d.D, bf, d.A, d.D);
  add(p->w_hh, d.H, base + lay.Whdb, d.H, bf, 4 * d.H, d.H);
  add(p->dec_att_w, d.H, base + lay.Whdb + (size_t)4 * d.H * d.H * es, d.H, bf, d.A, d.H);
  add(p->fbeta_w, d.H, base + lay.Whdb + (size_t)(4 * d.H + d.A) * d.H * es, d.H, bf, d.D, d.H);
  // Whdb0 = [W_hh ; 0 ; W_beta]
  add(p->w_hh, d.H, base + lay.Whdb0, d.H, bf, 4 * d.H, d.H);
  DIC_CUDA(cudaMemsetAsync(base + lay.Whdb0 + (size_t)4 * d.H * d.H * es, 0, (size_t)d.A * d.H * es, st));
  add(p->fbeta_w, d.H, base + lay.Whdb0 + (size_t)(4 * d.H + d.A) * d.H * es, d.H, bf, d.D, d.H);
  add(p->w_ih, d.E + d.D, base + lay.Wg, XW, bf, 4 * d.H, d.E + d.D);
  add(p->w_hh, d.H, base + lay.Wg + (size_t)(d.E + d.D) * es, XW, bf, 4 * d.H, d.H);
  if (bf && d.H % kGlUnits == 0 && gates_lstm_enabled()) {     // gate-interleaved copy for the cluster-fused LSTM step
    add(p->w_ih, d.E + d.D, base + lay.Wgp, XW, bf, 4 * d.H, d.E + d.D);
    jobs.j[jobs.n - 1].gate_H = d.H; jobs.j[jobs.n - 1].gate_U = kGlUnits;
    add(p->w_hh, d.H, base + lay.Wgp + (size_t)(d.E + d.D) * es, XW, bf, 4 * d.H, d.H);
    jobs.j[jobs.n - 1].gate_H = d.H; jobs.j[jobs.n - 1].gate_U = kGlUnits;
  }
  add(p->init_w, d.D, base + lay.Winit, d.D, bf, 2 * d.H, d.D);
  add(p->lin_w, d.H, base + lay.Wout, d.H, bf, d.V, d.H);
  add(p->embed_w, d.E, base + lay.Emb, d.E, bf, d.V, d.E);
  add(p->enc_att_b, d.A, base + lay.b_enc, d.A, 0, 1, d.A);
  add(p->dec_att_b, d.A, base + lay.bias_db, d.A, 0, 1, d.A);
  add(p->fbeta_b, d.D, base + lay.bias_db + sizeof(float) * d.A, d.D, 0, 1, d.D);
  add(p->b_ih, 4 * d.H, base + lay.bias_g, 4 * d.H, 0, 1, 4 * d.H, p->b_hh);     // b_ih + b_hh
  add(p->init_b, 2 * d.H, base + lay.b_init, 2 * d.H, 0, 1, 2 * d.H);
  add(p->lin_b, d.V, base + lay.b_out, d.V, 0, 1, d.V);
  add(p->full_att_w, d.A, base + lay.w_full, d.A, 0, 1, d.A);
  add(p->full_att_b, 1, base + lay.b_full, 1, 0, 1, 1);
  DIC_CUDA(launch_pdl(pack_jobs_kernel, dim3(296, jobs.n), dim3(256), 0, st, jobs));
  DIC_LAUNCH_CHECK();
  return 0;
}

int dic_adamw_step(int n, float* const* params, const float* const* grads, float* const* exp_avg,
                   float* const* exp_avg_sq, const long long* sizes, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int step, void* stream) {
  if (n < 0 || (n > 0 && (!params || !grads || !exp_avg || !exp_avg_sq || !sizes))) DIC_FAIL(-1, "null argument");
  if (step < 1) DIC_FAIL(-1, "step counts from 1");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  for (int base = 0; base < n; base += kMaxOptTensors) {
    AdamWJobs a;
    memset(&a, 0, sizeof(a));
    a.count = n - base < kMaxOptTensors ? n - base : kMaxOptTensors;
    long long biggest = 0;
    for (int i = 0; i < a.count; ++i) {
      a.p[i] = params[base + i]; a.g[i] = grads[base + i]; a.m[i] = exp_avg[base + i]; a.v[i] = exp_avg_sq[base + i];
      a.n[i] = sizes[base + i];
      if (!a.p[i] || !a.g[i] || !a.m[i] || !a.v[i] || a.n[i] < 0) DIC_FAIL(-1, "bad tensor %d", base + i);
      if (a.n[i] > biggest) biggest = a.n[i];
    }
    a.lr_wd = lr * weight_decay;
    a.one_m_b1 = 1.f - beta1; a.b2 = beta2; a.one_m_b2 = 1.f - beta2;
    a.step_size = (float)((double)lr / bc1);
    a.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    a.eps = eps;
    long long bx = (biggest / 4 + 255) / 256;
    if (bx > 148) bx = 148;
    if (bx < 1) bx = 1;
    DIC_CUDA(launch_pdl(adamw_kernel, dim3((unsigned)bx, a.count), dim3(256), 0, st, a));
    DIC_LAUNCH_CHECK();
  }
  return 0;
}

size_t dic_train_workspace_bytes(const dic_dims* dims, int dtype, int B, int T) {
  if (check_dims(dims, dtype) || B <= 0 || T <= 0) return 0;
  return TrainLayout(*dims, dtype, B, T).bytes;
}

int dic_decoder_forward_ex(const dic_dims* dims, int dtype, int attn_mode, const void* pack, const void* f_rgb,
                           const void* f_depth, int feat_dtype, const int64_t* captions, int cap_stride,
                           const int32_t* host_batch_sizes, int T, int B, const float* u, float temp,
                           const float* dropout_mask, void* logits, int logits_dtype, float* alphas,
                           void* workspace, size_t workspace_bytes, void* stream) {
  DIC_TRY(check_dims(dims, dtype));
  if (!pack || !f_rgb || !captions || !host_batch_sizes || !logits || !alphas || !workspace)
    DIC_FAIL(-1, "null argument");
  if (attn_mode < 0 || attn_mode > 2) DIC_FAIL(-1, "bad attn_mode %d", attn_mode);
  if (attn_mode != DIC_ATTN_SOFT && !u) DIC_FAIL(-1, "hard attention needs the uniform draws u");
  if (attn_mode == DIC_ATTN_GUMBEL_SOFTMAX && !(temp > 0.f)) DIC_FAIL(-1, "temp must be > 0");
  if (logits_dtype != DIC_F32 && !(logits_dtype == DIC_BF16 && dtype == DIC_BF16))
    DIC_FAIL(-1, "logits are float32, or bfloat16 in bf16 mode");
  StepSizes sizes;
  int total = 0;
  DIC_TRY(make_sizes(host_batch_sizes, T, B, &sizes, &total));
  if (sizes.n[0] != B) DIC_FAIL(-1, "batch_sizes[0]=%d must equal B=%d", sizes.n[0], B);
  if (T + 1 > cap_stride) DIC_FAIL(-1, "captions has %d columns, need >= T+1 = %d", cap_stride, T + 1);
  const size_t need = TrainLayout(*dims, dtype, B, T).bytes;
  if (workspace_bytes < need) DIC_FAIL(-1, "workspace too small: %zu < %zu", workspace_bytes, need);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* ws = reinterpret_cast<char*>(workspace);
  const int lb = logits_dtype == DIC_BF16;
  if (dtype == DIC_BF16)
    return decoder_forward_impl<bf16>(*dims, attn_mode, pack, f_rgb, f_depth, feat_dtype, captions, cap_stride,
                                      sizes, total, T, B, u, temp, dropout_mask, logits, lb, alphas, ws, st);
  return decoder_forward_impl<float>(*dims, attn_mode, pack, f_rgb, f_depth, feat_dtype, captions, cap_stride,
                                     sizes, total, T, B, u, temp, dropout_mask, logits, lb, alphas, ws, st);
}

int dic_decoder_forward(const dic_dims* dims, int dtype, int attn_mode, const void* pack, const void* f_rgb,
                        const void* f_depth, int feat_dtype, const int64_t* captions, int cap_stride,
                        const int32_t* host_batch_sizes, int T, int B, const float* u, float temp,
                        const float* dropout_mask, float* logits, float* alphas, void* workspace,
                        size_t workspace_bytes, void* stream) {
  return dic_decoder_forward_ex(dims, dtype, attn_mode, pack, f_rgb, f_depth, feat_dtype, captions, cap_stride,
                                host_batch_sizes, T, B, u, temp, dropout_mask, logits, DIC_F32, alphas, workspace,
                                workspace_bytes, stream);
}

int dic_decoder_backward_ex(const dic_dims* dims, int dtype, int attn_mode, const void* pack, const void* f_rgb,
                            const void* f_depth, int feat_dtype, const int64_t* captions, int cap_stride,
                            const int32_t* host_batch_sizes, int T, int B, const void* d_logits,
                            int d_logits_dtype, const float* d_alphas, const float* alphas, float temp,
                            const float* dropout_mask, const dic_params* grads, void* d_feats, void* workspace,
                            size_t workspace_bytes, void* stream) {
  DIC_TRY(check_dims(dims, dtype));
  // same aliasing rule as prologue(): no fused copy was made when there is no depth tensor and
  // the annotations already have the storage dtype
  const void* f_alias =
      (f_depth == nullptr && feat_dtype == dtype) ? f_rgb : nullptr;
  if (!pack || !f_rgb || !captions || !host_batch_sizes || !d_logits || !alphas || !grads || !workspace)
    DIC_FAIL(-1, "null argument");
  if (d_logits_dtype != DIC_F32 && d_logits_dtype != dtype)
    DIC_FAIL(-1, "d_logits must be float32 or the storage dtype of the mode");
  const int dl_is_st = d_logits_dtype == dtype;
  StepSizes sizes;
  int total = 0;
  DIC_TRY(make_sizes(host_batch_sizes, T, B, &sizes, &total));
  const size_t need = TrainLayout(*dims, dtype, B, T).bytes;
  if (workspace_bytes < need) DIC_FAIL(-1, "workspace too small: %zu < %zu", workspace_bytes, need);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* ws = reinterpret_cast<char*>(workspace);
  if (dtype == DIC_BF16)
    return decoder_backward_impl<bf16>(*dims, attn_mode, pack, captions, cap_stride, sizes, total, T, B,
                                       d_logits, dl_is_st, d_alphas, alphas, temp, dropout_mask, *grads, d_feats,
                                       feat_dtype == DIC_BF16, f_alias,
                                       ws, st);
  return decoder_backward_impl<float>(*dims, attn_mode, pack, captions, cap_stride, sizes, total, T, B, d_logits,
                                      dl_is_st, d_alphas, alphas, temp, dropout_mask, *grads, d_feats,
                                      feat_dtype == DIC_BF16, f_alias, ws, st);
}

int dic_decoder_backward(const dic_dims* dims, int dtype, int attn_mode, const void* pack, const void* f_rgb,
                         const void* f_depth, int feat_dtype, const int64_t* captions, int cap_stride,
                         const int32_t* host_batch_sizes, int T, int B, const float* d_logits,
                         const float* d_alphas, const float* alphas, float temp, const float* dropout_mask,
                         const dic_params* grads, void* d_feats, void* workspace, size_t workspace_bytes,
                         void* stream) {
  // fp32 d_logits: bf16 mode makes its tensor-core operand copy, fp32 mode uses them as they are
  return dic_decoder_backward_ex(dims, dtype, attn_mode, pack, f_rgb, f_depth, feat_dtype, captions, cap_stride,
                                 host_batch_sizes, T, B, d_logits, DIC_F32,
                                 d_alphas, alphas, temp, dropout_mask, grads, d_feats, workspace, workspace_bytes,
                                 stream);
}

size_t dic_caption_loss_workspace_bytes(int N, int B) {
  if (N <= 0 || B <= 0) return 0;
  return loss_workspace_bytes(N, B);
}

int dic_caption_loss(const dic_dims* dims, int dtype, const void* logits, int logits_dtype, const int64_t* captions,
                     int cap_stride, const int32_t* host_batch_sizes, int T, int B, int ignore_index,
                     const float* alphas, float lam, float* loss, void* d_logits, float* d_alphas, void* workspace,
                     size_t workspace_bytes, void* stream) {
  DIC_TRY(check_dims(dims, dtype));
  if (!logits || !captions || !host_batch_sizes || !loss || !d_logits || !workspace) DIC_FAIL(-1, "null argument");
  StepSizes sizes;
  int total = 0;
  DIC_TRY(make_sizes(host_batch_sizes, T, B, &sizes, &total));
  if (T + 1 > cap_stride) DIC_FAIL(-1, "captions has %d columns, need >= T+1 = %d", cap_stride, T + 1);
  if (workspace_bytes < loss_workspace_bytes(total, B)) DIC_FAIL(-1, "workspace too small");
  if (logits_dtype != DIC_F32 && !(logits_dtype == DIC_BF16 && dtype == DIC_BF16))
    DIC_FAIL(-1, "logits are float32, or bfloat16 in bf16 mode");
  if (logits == d_logits && logits_dtype != dtype) DIC_FAIL(-1, "in-place d_logits needs logits in the storage dtype");
  LossArgs p;
  memset(&p, 0, sizeof(p));
  p.logits = logits; p.logits_bf16 = logits_dtype == DIC_BF16; p.captions = captions; p.cap_stride = cap_stride; p.sizes = sizes;
  p.T = T; p.B = B; p.N = total; p.V = dims->V; p.L = dims->L; p.ignore_index = ignore_index;
  p.alphas = alphas; p.lam = lam; p.loss = loss; p.d_logits = d_logits; p.d_alphas = d_alphas;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == DIC_BF16) return launch_caption_loss<bf16>(p, workspace, st);
  return launch_caption_loss<float>(p, workspace, st);
}

int dic_scale_loss_grads(int dtype, const float* grad_loss, void* d_logits, size_t n_logits, float* d_alphas,
                         size_t n_alphas, void* stream) {
  if (!grad_loss || !d_logits) DIC_FAIL(-1, "null argument");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == DIC_BF16)
    loss_scale_kernel<bf16><<<592, 256, 0, st>>>(grad_loss, reinterpret_cast<bf16*>(d_logits), n_logits, d_alphas, n_alphas);
  else
    loss_scale_kernel<float><<<592, 256, 0, st>>>(grad_loss, reinterpret_cast<float*>(d_logits), n_logits, d_alphas, n_alphas);
  DIC_LAUNCH_CHECK();
  return 0;
}

size_t dic_decode_workspace_bytes(const dic_dims* dims, int dtype, int B, int beam) {
  if (check_dims(dims, dtype) || B <= 0 || beam <= 0) return 0;
  return DecodeLayout(*dims, dtype, B, beam).bytes;
}

int dic_decode_greedy(const dic_dims* dims, int dtype, int attn_mode, const void* pack, const void* f_rgb,
                      const void* f_depth, int feat_dtype, int B, int start_id, int max_len, const float* u,
                      int64_t* tokens, float* alphas_out, float* logits_out, void* workspace,
                      size_t workspace_bytes, void* stream) {
  DIC_TRY(check_dims(dims, dtype));
  if (!pack || !f_rgb || !tokens || !workspace) DIC_FAIL(-1, "null argument");
  if (attn_mode != DIC_ATTN_SOFT && attn_mode != DIC_ATTN_GUMBEL_MAX)
    DIC_FAIL(-1, "greedy decode supports soft and gumbel-max attention");
  if (attn_mode == DIC_ATTN_GUMBEL_MAX && !u) DIC_FAIL(-1, "gumbel-max needs the uniform draws u");
  if (B <= 0 || max_len <= 0 || max_len > DIC_MAX_STEPS) DIC_FAIL(-1, "bad B/max_len");
  if (start_id < 0 || start_id >= dims->V) DIC_FAIL(-1, "start_id out of range");
  const size_t need = DecodeLayout(*dims, dtype, B, 1).bytes;
  if (workspace_bytes < need) DIC_FAIL(-1, "workspace too small: %zu < %zu", workspace_bytes, need);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* ws = reinterpret_cast<char*>(workspace);
  if (dtype == DIC_BF16)
    return decode_impl<bf16>(*dims, attn_mode, pack, f_rgb, f_depth, feat_dtype, B, 1, false, start_id, -1,
                             max_len, u, tokens, nullptr, nullptr, alphas_out, logits_out, nullptr, nullptr,
                             nullptr, nullptr, ws, st);
  return decode_impl<float>(*dims, attn_mode, pack, f_rgb, f_depth, feat_dtype, B, 1, false, start_id, -1, max_len,
                            u, tokens, nullptr, nullptr, alphas_out, logits_out, nullptr, nullptr, nullptr,
                            nullptr, ws, st);
}

int dic_decode_beam(const dic_dims* dims, int dtype, const void* pack, const void* f_rgb, const void* f_depth,
                    int feat_dtype, int B, int beam, int start_id, int end_id, int max_len, int64_t* tokens,
                    int32_t* lengths, float* scores, int32_t* back, int32_t* toks, float* step_scores,
                    float* lse, float* logits_out, void* workspace, size_t workspace_bytes, void* stream) {
  DIC_TRY(check_dims(dims, dtype));
  if (!pack || !f_rgb || !tokens || !lengths || !scores || !workspace) DIC_FAIL(-1, "null argument");
  if (beam < 1 || beam > DIC_MAX_BEAM) DIC_FAIL(-1, "beam must be in 1..%d", DIC_MAX_BEAM);
  if (beam > dims->V) DIC_FAIL(-1, "beam larger than vocabulary");
  if (B <= 0 || max_len <= 0 || max_len > DIC_MAX_STEPS) DIC_FAIL(-1, "bad B/max_len");
  if (start_id < 0 || start_id >= dims->V || end_id < 0 || end_id >= dims->V)
    DIC_FAIL(-1, "start/end id out of range");
  const size_t need = DecodeLayout(*dims, dtype, B, beam).bytes;
  if (workspace_bytes < need) DIC_FAIL(-1, "workspace too small: %zu < %zu", workspace_bytes, need);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* ws = reinterpret_cast<char*>(workspace);
  if (dtype == DIC_BF16)
    return decode_impl<bf16>(*dims, DIC_ATTN_SOFT, pack, f_rgb, f_depth, feat_dtype, B, beam, true, start_id,
                             end_id, max_len, nullptr, tokens, lengths, scores, nullptr, logits_out, back, toks,
                             step_scores, lse, ws, st);
  return decode_impl<float>(*dims, DIC_ATTN_SOFT, pack, f_rgb, f_depth, feat_dtype, B, beam, true, start_id,
                            end_id, max_len, nullptr, tokens, lengths, scores, nullptr, logits_out, back, toks,
                            step_scores, lse, ws, st);
}

// ---- standalone attention module -------------------------------------------------------------
namespace {
struct AttnWsLayout {
  size_t Fst, att1, hp, Wenc, Wdec, hst, zg, bias, bytes;
  int es;
  AttnWsLayout(const dic_dims& d, int dtype, int B) {
    es = dtype == DIC_BF16 ? 2 : 4;
    Carver c;
    Fst = c.take((size_t)B * d.L * d.D * es);
    att1 = c.take((size_t)B * d.L * d.A * es);
    hp = c.take(sizeof(float) * B * (d.A + d.D));
    Wenc = c.take((size_t)d.A * d.D * es);
    Wdec = c.take((size_t)d.A * d.H * es);
    hst = c.take((size_t)B * d.H * es);
    zg = c.take((size_t)B * d.D * es);
    bias = c.take(sizeof(float) * 4);
    bytes = c.off;
  }
};
__global__ void fill_kernel(float* p, float v, size_t n) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) p[i] = v;
}
__global__ void cast_out_kernel(const void* src, int src_bf16, float* dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256)
    dst[i] = ld_as_float(src, i, src_bf16);
}
}  // namespace

size_t dic_attention_workspace_bytes(const dic_dims* dims, int dtype, int B) {
  if (check_dims(dims, dtype) || B <= 0) return 0;
  return AttnWsLayout(*dims, dtype, B).bytes;
}

}  // extern "C" (templates need C++ linkage)

template <typename ST>
static int attention_forward_impl(const dic_dims& d, int attn_mode, const float* enc_w, const float* enc_b,
                                  const float* dec_w, const float* dec_b, const float* full_w,
                                  const float* full_b, const float* feats, const float* h, int B,
                                  const float* u, float temp, float* context, float* alpha, char* ws,
                                  cudaStream_t st) {
  const int dtype = sizeof(ST) == 2 ? DIC_BF16 : DIC_F32;
  const int bf = sizeof(ST) == 2;
  const AttnWsLayout lay(d, dtype, B);
  ST* Fst = reinterpret_cast<ST*>(ws + lay.Fst);
  ST* att1 = reinterpret_cast<ST*>(ws + lay.att1);
  float* hp = reinterpret_cast<float*>(ws + lay.hp);
  ST* Wenc = reinterpret_cast<ST*>(ws + lay.Wenc);
  ST* Wdec = reinterpret_cast<ST*>(ws + lay.Wdec);
  ST* hst = reinterpret_cast<ST*>(ws + lay.hst);
  ST* zg = reinterpret_cast<ST*>(ws + lay.zg);
  const ST* F = reinterpret_cast<const ST*>(feats);
  if (bf) {
    DIC_TRY(launch_copy2d(feats, d.D, Fst, d.D, 1, B * d.L, d.D, st));
    F = Fst;
  }
  DIC_TRY(launch_copy2d(enc_w, d.D, Wenc, d.D, bf, d.A, d.D, st));
  DIC_TRY(launch_copy2d(dec_w, d.H, Wdec, d.H, bf, d.A, d.H, st));
  DIC_TRY(launch_copy2d(h, d.H, hst, d.H, bf, B, d.H, st));
  GemmArgs g = gemm_args_nt(F, bf, d.D, Wenc, bf, d.D, att1, bf, d.A, B * d.L, d.A, d.D, enc_b);
  DIC_TRY(gemm(g, st));
  // hp = [att2 | beta := 1]  (the standalone module has no f_beta gate)
  fill_kernel<<<148, 256, 0, st>>>(hp, 1.f, (size_t)B * (d.A + d.D));
  DIC_LAUNCH_CHECK();
  g = gemm_args_nt(hst, bf, d.H, Wdec, bf, d.H, hp, 0, d.A + d.D, B, d.A, d.H, dec_b);
  DIC_TRY(gemm_generic(g, st));
  AttnFwdArgs a;
  memset(&a, 0, sizeof(a));
  a.F = F; a.att1 = att1; a.hp = hp; a.w_full = full_w; a.b_full = full_b; a.u = u;
  a.alpha_out = alpha; a.alpha_stride = d.L;
  a.z_out = context;
  a.zg_out = zg; a.zg_stride = d.D;
  a.L = d.L; a.D = d.D; a.A = d.A; a.mode = attn_mode;
  a.inv_temp = attn_mode == DIC_ATTN_GUMBEL_SOFTMAX ? 1.f / temp : 1.f;
  return launch_attn_step<ST>(a, B, 1, st);
}

extern "C" {

int dic_attention_forward(const dic_dims* dims, int dtype, int attn_mode, const float* enc_w, const float* enc_b,
                          const float* dec_w, const float* dec_b, const float* full_w, const float* full_b,
                          const float* feats, const float* h, int B, const float* u, float temp, float* context,
                          float* alpha, void* workspace, size_t workspace_bytes, void* stream) {
  DIC_TRY(check_dims(dims, dtype));
  if (!enc_w || !enc_b || !dec_w || !dec_b || !full_w || !full_b || !feats || !h || !context || !alpha ||
      !workspace)
    DIC_FAIL(-1, "null argument");
  if (attn_mode != DIC_ATTN_SOFT && !u) DIC_FAIL(-1, "hard attention needs the uniform draws u");
  if (attn_mode == DIC_ATTN_GUMBEL_SOFTMAX && !(temp > 0.f)) DIC_FAIL(-1, "temp must be > 0");
  const size_t need = AttnWsLayout(*dims, dtype, B).bytes;
  if (workspace_bytes < need) DIC_FAIL(-1, "workspace too small: %zu < %zu", workspace_bytes, need);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* ws = reinterpret_cast<char*>(workspace);
  if (dtype == DIC_BF16)
    return attention_forward_impl<bf16>(*dims, attn_mode, enc_w, enc_b, dec_w, dec_b, full_w, full_b, feats, h, B,
                                        u, temp, context, alpha, ws, st);
  return attention_forward_impl<float>(*dims, attn_mode, enc_w, enc_b, dec_w, dec_b, full_w, full_b, feats, h, B, u,
                                       temp, context, alpha, ws, st);
}

size_t dic_beam_select_workspace_bytes(int B, int K) {
  if (B <= 0 || K <= 0) return 0;
  return beam_select_workspace_bytes(B, K);
}

int dic_beam_select(const float* scores, const uint8_t* finished, const float* logits, const float* lse, int B,
                    int K, int V, int end_id, float* new_scores, int32_t* back, int32_t* tok,
                    uint8_t* new_finished, void* workspace, size_t workspace_bytes, void* stream) {
  if (!scores || !finished || !logits || !lse || !new_scores || !back || !tok || !new_finished || !workspace)
    DIC_FAIL(-1, "null argument");
  if (B <= 0 || K <= 0 || K > DIC_MAX_BEAM) DIC_FAIL(-1, "bad B/K");
  if (workspace_bytes < beam_select_workspace_bytes(B, K)) DIC_FAIL(-1, "workspace too small");
  return launch_beam_select(scores, finished, logits, lse, nullptr, B, K, V, end_id, workspace, new_scores, back,
                            tok, new_finished, reinterpret_cast<cudaStream_t>(stream));
}

int dic_row_lse(const float* logits, int R, int V, float* lse, void* stream) {
  if (!logits || !lse || R <= 0 || V <= 0) DIC_FAIL(-1, "bad argument");
  row_lse_kernel<<<R, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(logits, V, lse);
  DIC_LAUNCH_CHECK();
  return 0;
}

int dic_dfeat_gemm(const void* datt1, const void* w_enc, const void* alpha16, int Lp, const void* dz,
                   const float* dmeanF, void* dF, int B, int L, int D, int A, int T, void* stream) {
  if (!datt1 || !w_enc || !alpha16 || !dz || !dmeanF || !dF) DIC_FAIL(-1, "null argument");
  if (!dfeat_tc_eligible(A, D, Lp)) DIC_FAIL(-5, "shape not eligible for the fused dL/dF GEMM");
  return launch_dfeat_tc(reinterpret_cast<const bf16*>(datt1), reinterpret_cast<const bf16*>(w_enc),
                         reinterpret_cast<const bf16*>(alpha16), Lp, reinterpret_cast<const bf16*>(dz), dmeanF,
                         reinterpret_cast<bf16*>(dF), B, L, D, A, T, reinterpret_cast<cudaStream_t>(stream));
}

int dic_gemm_nt(int engine, int M, int N, int K, const void* A, int a_dtype, const void* B, int b_dtype,
                const float* bias, float* C, void* stream) {
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0) DIC_FAIL(-1, "bad argument");
  GemmArgs g = gemm_args_nt(A, a_dtype == DIC_BF16, K, B, b_dtype == DIC_BF16, K, C, 0, N, M, N, K, bias);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (engine == 1) {
    if (!tc_gemm_eligible(g)) DIC_FAIL(-5, "shape/dtype not eligible for the tcgen05 engine");
    return tc_gemm(g, st);
  }
  return gemm_generic(g, st);
}

int dic_gemm_ex(int engine, int M, int N, int K, const void* A, int a_dtype, long long a_m, long long a_k,
                const void* B, int b_dtype, long long b_n, long long b_k, const float* bias, float* C,
                long long ldc, int splits, void* stream) {
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0 || splits < 1) DIC_FAIL(-1, "bad argument");
  GemmArgs g = gemm_args_nt(A, a_dtype == DIC_BF16, a_m, B, b_dtype == DIC_BF16, b_n, C, 0, ldc, M, N, K, bias);
  g.a_k = a_k;
  g.b_k = b_k;
  g.splits = splits;
  g.split_mode = 0;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (engine == 1) {
    if (!tc_gemm_eligible(g)) DIC_FAIL(-5, "shape/dtype not eligible for the tcgen05 engine");
    return tc_gemm(g, st);
  }
  return gemm_generic(g, st);
}

int dic_gemm_nt_bf16(int engine, int M, int N, int K, const void* A, const void* B, const float* bias, void* C,
                     long long ldc, void* stream) {
  if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0 || ldc < N) DIC_FAIL(-1, "bad argument");
  GemmArgs g = gemm_args_nt(A, 1, K, B, 1, K, C, 1, ldc, M, N, K, bias);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (engine == 1) {
    if (!tc_gemm_eligible(g)) DIC_FAIL(-5, "shape/dtype not eligible for the tcgen05 engine");
    return tc_gemm(g, st);
  }
  return gemm_generic(g, st);
}

}  // extern "C"
