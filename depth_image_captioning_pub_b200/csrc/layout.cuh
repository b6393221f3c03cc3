// Host-side memory plans: the packed-weight block and the training / decode workspaces.
// Everything lives in caller-provided device memory; these structs only compute offsets.
#pragma once
#include "common.cuh"

namespace dic {

struct Carver {
  size_t off = 0;
  size_t take(size_t bytes) {
    const size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  }
};

// Compute-layout weights.  ST = storage dtype of the mode (fp32 or bf16).
//   Wenc  [A,D]            attention.encoder_att.weight
//   Whdb  [4H+A+D, H]      rows: W_hh | W_dec | W_beta  (backward: dh = G . Whdb)
//         Wdb = Whdb + 4H*H  -> [A+D, H] forward h-projection (att2 | beta)
//   Whdb0 [4H+A+D, H]      rows: W_hh | 0 | W_beta: the part of dh that does not wait for the attention backward
//   Wg    [4H, E+D+H]      [W_ih | W_hh]: gates = [emb|zg|h] . Wg^T
//   Wgp   [4H, E+D+H]      Wg with rows interleaved per tile of U = 32 units: row (u/U)*4U + gate*U + u%U = Wg row gate*H + u
//   Winit [2H,D], Wout [V,H], Emb [V,E]
//   fp32: b_enc[A], bias_db[A+D] = b_dec|b_beta, bias_g[4H] = b_ih+b_hh, b_init[2H], b_out[V],
//         w_full[A], b_full[1]
struct PackLayout {
  size_t Wenc, Whdb, Whdb0, Wg, Wgp, Winit, Wout, Emb;
  size_t b_enc, bias_db, bias_g, b_init, b_out, w_full, b_full;
  size_t bytes;
  int es;  // element size of ST
  PackLayout(const dic_dims& d, int dtype) {
    es = dtype == DIC_BF16 ? 2 : 4;
    Carver c;
    const size_t XW = (size_t)d.E + d.D + d.H;
    Wenc = c.take((size_t)d.A * d.D * es);
    Whdb = c.take((size_t)(4 * d.H + d.A + d.D) * d.H * es);
    Whdb0 = c.take((size_t)(4 * d.H + d.A + d.D) * d.H * es);   // Whdb with the W_dec rows zeroed (off-chain dh GEMM)
    Wg = c.take((size_t)4 * d.H * XW * es);
    Wgp = c.take((size_t)4 * d.H * XW * es);   // gate-interleaved copy of Wg (gates_lstm.cuh)
    Winit = c.take((size_t)2 * d.H * d.D * es);
    Wout = c.take((size_t)d.V * d.H * es);
    Emb = c.take((size_t)d.V * d.E * es);
    b_enc = c.take(sizeof(float) * d.A);
    bias_db = c.take(sizeof(float) * (d.A + d.D));
    bias_g = c.take(sizeof(float) * 4 * d.H);
    b_init = c.take(sizeof(float) * 2 * d.H);
    b_out = c.take(sizeof(float) * d.V);
    w_full = c.take(sizeof(float) * d.A);
    b_full = c.take(sizeof(float) * 4);
    bytes = c.off;
  }
};

// Views into a packed-weight block.
struct Pack {
  const char* base;
  PackLayout lay;
  Pack(const void* p, const dic_dims& d, int dtype) : base(reinterpret_cast<const char*>(p)), lay(d, dtype) {}
  const void* Wenc() const { return base + lay.Wenc; }
  const void* Whdb() const { return base + lay.Whdb; }
  const void* Whdb0() const { return base + lay.Whdb0; }
  const void* Wdec(const dic_dims& d) const { return base + lay.Whdb + (size_t)4 * d.H * d.H * lay.es; }
  const void* Wdb(const dic_dims& d) const { return base + lay.Whdb + (size_t)4 * d.H * d.H * lay.es; }
  const void* Wg() const { return base + lay.Wg; }
  const void* Wgp() const { return base + lay.Wgp; }
  const void* Winit() const { return base + lay.Winit; }
  const void* Wout() const { return base + lay.Wout; }
  const void* Emb() const { return base + lay.Emb; }
  const float* b_enc() const { return reinterpret_cast<const float*>(base + lay.b_enc); }
  const float* bias_db() const { return reinterpret_cast<const float*>(base + lay.bias_db); }
  const float* bias_g() const { return reinterpret_cast<const float*>(base + lay.bias_g); }
  const float* b_init() const { return reinterpret_cast<const float*>(base + lay.b_init); }
  const float* b_out() const { return reinterpret_cast<const float*>(base + lay.b_out); }
  const float* w_full() const { return reinterpret_cast<const float*>(base + lay.w_full); }
  const float* b_full() const { return reinterpret_cast<const float*>(base + lay.b_full); }
};

constexpr int kGateSplitsMax = 16;
constexpr int kDzgSplits = 2;        // K halves of the per-step dzg GEMM (atomic accumulation into a zeroed buffer)
constexpr int kDhSplitsMax = 10;     // split-K partial buffers of the per-step dh GEMM (<= 16)

// Training workspace (forward state saved for backward + backward scratch).
struct TrainLayout {
  size_t Fsum, meanF, att1, XH, HP, Z, acts, c_all, gate_part, Hdrop;
  size_t G, DZ, de, dzg, dh, dc, dHout, dwfull_part, dbfull_part, datt1, dXemb, dmeanF, tmpvec, dlogits16, dal_part, h0, dF32, dh_part;
  size_t alpha16, meanF16, hc0, dhc16, ready, done;
  int Lp;
  size_t bytes;
  size_t XW, GW;
  int es;
  TrainLayout(const dic_dims& d, int dtype, int B, int T) {
    es = dtype == DIC_BF16 ? 2 : 4;
    XW = (size_t)d.E + d.D + d.H;
    GW = (size_t)4 * d.H + d.A + d.D;
    const size_t TB = (size_t)T * B;
    Carver c;
    Fsum = c.take((size_t)B * d.L * d.D * es);
    meanF = c.take(sizeof(float) * B * d.D);
    att1 = c.take((size_t)B * d.L * d.A * es);
    XH = c.take((TB + B) * XW * es);
    HP = c.take(sizeof(float) * TB * (d.A + d.D));
    Z = c.take(sizeof(float) * TB * d.D);
    acts = c.take(sizeof(float) * TB * 4 * d.H);
    c_all = c.take(sizeof(float) * (TB + B) * d.H);
    gate_part = c.take(sizeof(float) * kGateSplitsMax * B * 4 * d.H);
    Hdrop = c.take(TB * d.H * es);
    G = c.take(TB * GW * es);
    DZ = c.take(TB * d.D * es);
    de = c.take(sizeof(float) * TB * d.L);
    dzg = c.take(sizeof(float) * B * d.D);
    dh = c.take(sizeof(float) * B * d.H);
    dc = c.take(sizeof(float) * B * d.H);
    dHout = c.take(sizeof(float) * TB * d.H);
    dwfull_part = c.take(sizeof(float) * TB * d.A);
    dbfull_part = c.take(sizeof(float) * TB);
    datt1 = c.take((size_t)B * d.L * d.A * es);
    dXemb = c.take(sizeof(float) * TB * d.E);
    dmeanF = c.take(sizeof(float) * B * d.D);
    tmpvec = c.take(sizeof(float) * (GW + 16));
    dlogits16 = c.take(dtype == DIC_BF16 ? TB * d.V * 2 : 16);   // bf16 copy of d_logits (GEMM operand)
    dal_part = c.take(sizeof(float) * (size_t)((d.D + 255) / 256) * B * d.L);   // per-chunk dalpha partials
    h0 = c.take(sizeof(float) * B * d.H);
    dh_part = c.take(sizeof(float) * (kDhSplitsMax + 1) * B * d.H);   // + the datt2 . W_dec slot of the small kernel
    dF32 = c.take(sizeof(float) * (size_t)B * d.L * d.D);   // fp32 dL/dF accumulator when annotations are bf16
    Lp = (d.L + 7) & ~7;
    alpha16 = c.take(TB * Lp * 2);     // bf16 alpha [B,T,Lp]: A operand of the fused dL/dF GEMM
    ready = c.take(sizeof(unsigned int) * B);   // per-image alpha -> context hand-off flags (epoch = step + 1)
    done = c.take(sizeof(unsigned int) * B);    // per-image streaming-backward -> small-backward arrival counters
    meanF16 = c.take((size_t)B * d.D * 2);
    hc0 = c.take(sizeof(float) * B * 2 * d.H);
    dhc16 = c.take((size_t)B * 2 * d.H * 2);
    bytes = c.off;
  }
};

// Decode workspace: R = B*beam rows.
struct DecodeLayout {
  size_t Fsum, meanF, att1, XH, HP, c, c_tmp, h_tmp, h0, c0, gate_part, logits, lse;
  size_t scores, scores2, fin, fin2, back, tok, step_scores, alpha, cand, meanF16, hc0, bstats, ZH, c_par, etab;
  size_t bytes;
  size_t XW;
  int es;
  DecodeLayout(const dic_dims& d, int dtype, int B, int beam, int max_len = DIC_MAX_STEPS) {
    es = dtype == DIC_BF16 ? 2 : 4;
    XW = (size_t)d.E + d.D + d.H;
    const size_t R = (size_t)B * beam;
    Carver c_;
    Fsum = c_.take((size_t)B * d.L * d.D * es);
    meanF = c_.take(sizeof(float) * B * d.D);
    att1 = c_.take((size_t)B * d.L * d.A * es);
    XH = c_.take(2 * R * XW * es);
    HP = c_.take(sizeof(float) * R * (d.A + d.D));
    c = c_.take(sizeof(float) * R * d.H);
    c_tmp = c_.take(sizeof(float) * R * d.H);
    h_tmp = c_.take(R * d.H * es);
    h0 = c_.take(sizeof(float) * B * d.H);
    c0 = c_.take(sizeof(float) * B * d.H);
    gate_part = c_.take(sizeof(float) * kGateSplitsMax * R * 4 * d.H);
    logits = c_.take(sizeof(float) * R * d.V);
    lse = c_.take(sizeof(float) * R);
    scores = c_.take(sizeof(float) * R);
    scores2 = c_.take(sizeof(float) * R);
    fin = c_.take(R);
    fin2 = c_.take(R);
    back = c_.take(sizeof(int32_t) * R * max_len);
    tok = c_.take(sizeof(int32_t) * R * max_len);
    step_scores = c_.take(sizeof(float) * R * max_len);
    alpha = c_.take(sizeof(float) * R * d.L);
    cand = c_.take(R * beam * (sizeof(float) + sizeof(int)));     // per-row top-K candidates
    meanF16 = c_.take((size_t)B * d.D * 2);
    hc0 = c_.take(sizeof(float) * B * 2 * d.H);
    bstats = c_.take(sizeof(float) * 2 * R * ((d.V + 79) / 80));   // per-slice (max, sum-exp) pairs of the fused beam step
    // look-ahead beam step (dic_api.cu decode_impl): rows in PARENT order, double buffered over the steps
    ZH = c_.take(2 * R * ((size_t)d.D + d.H) * es);                  // [beta.z of step t+1 | h_t]: A operand of the gate GEMM
    c_par = c_.take(2 * sizeof(float) * R * d.H);                    // c_t
    etab = c_.take(sizeof(float) * (size_t)d.V * 4 * d.H);           // Emb . W_ih[:, :E]^T: a token's share of the gates
    bytes = c_.off;
  }
};

}  // namespace dic
