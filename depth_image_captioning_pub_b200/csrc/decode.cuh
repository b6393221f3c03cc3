// Decode-side kernels: greedy argmax + next-token embedding, row log-sum-exp, beam top-k with
// backpointers, beam state reorder and final backtrack.  No host synchronisation anywhere
// (the reference copies the argmax to the host every step, depth_models.py:298-299).
#pragma once
#include "common.cuh"

namespace dic {

// ---- initial state: replicate (h0,c0) over the rows of each image, embed <start> ---------------
template <typename ST>
__global__ void __launch_bounds__(256) decode_init_kernel(const float* __restrict__ h0,
                                                          const float* __restrict__ c0, int hc_stride,
                                                          const ST* __restrict__ emb, int start_id,
                                                          ST* __restrict__ X, long long x_row, int col_h,
                                                          float* __restrict__ c, int rows, int KB, int E,
                                                          int H) {
  const int W = E + H;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < rows * W; i += gridDim.x * 256) {
    const int r = i / W, q = i - r * W;
    const int b = r / KB;
    if (q < E) {
      X[(size_t)r * x_row + q] = emb[(size_t)start_id * E + q];
    } else {
      const int j = q - E;
      X[(size_t)r * x_row + col_h + j] = from_f<ST>(h0[(size_t)b * hc_stride + j]);
      c[(size_t)r * H + j] = c0[(size_t)b * hc_stride + j];
    }
  }
}

// ---- greedy: token = argmax_v logits[r,v] (ties -> lowest id), then embed it ---------------------
// The reference takes argmax(softmax(logits)) (depth_models.py:296-297); softmax is monotone,
// so the argmax of the logits is the same token except on fp32 rounding ties.
template <typename ST>
__global__ void __launch_bounds__(256) argmax_embed_kernel(const float* __restrict__ logits, int V,
                                                           int64_t* __restrict__ tokens,
                                                           long long tok_stride, const ST* __restrict__ emb,
                                                           int E, ST* __restrict__ Xnext, long long x_row, TraceRec* trace_buf) {
  __shared__ float sv[8];
  __shared__ int si[8];
  __shared__ int s_tok;
  Trace trace(trace_buf);
  pdl_wait();
  pdl_trigger();
  trace.mark();
  const int r = blockIdx.x;
  const float* lg = logits + (size_t)r * V;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int v = threadIdx.x; v < V; v += 256) {
    const float x = lg[v];
    if (x > best) { best = x; bi = v; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sv[warp] = best; si[warp] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (sv[w] > best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
    if (bi == 0x7fffffff) bi = 0;
    s_tok = bi;
    tokens[(size_t)r * tok_stride] = bi;
  }
  __syncthreads();
  const int tok = s_tok;
  for (int e = threadIdx.x; e < E; e += 256) Xnext[(size_t)r * x_row + e] = emb[(size_t)tok * E + e];
  trace.end(TK_ARGMAX);
}

// ---- lse[r] = log sum_v exp(logits[r,v]) -------------------------------------------------------
__global__ void __launch_bounds__(256) row_lse_kernel(const float* __restrict__ logits, int V,
                                                      float* __restrict__ lse) {
  __shared__ float scratch[64];
  pdl_wait();
  pdl_trigger();
  const int r = blockIdx.x;
  const float* lg = logits + (size_t)r * V;
  float m = -INFINITY;
  for (int v = threadIdx.x; v < V; v += 256) m = fmaxf(m, lg[v]);
  m = block_max(m, scratch);
  float s = 0.f;
  for (int v = threadIdx.x; v < V; v += 256) s += expf(lg[v] - m);
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) lse[r] = m + logf(s);
}

// ---- beam selection ----------------------------------------------------------------------------
// cand[j*V+v] = scores[j] + (logits[j,v] - lse[j]); finished rows keep only <end> at cost 0.
// Stable top-K: order by (value desc, flat index asc) -- bit-exact w.r.t. the oracle's
// beam_select given identical inputs (plain fp32 add/sub, round-to-nearest, no contraction).
//
// Two kernels.  (1) beam_row_topk_kernel: one CTA per beam ROW computes the row's log-sum-exp
// (unless given) and its own top-K candidates; (2) beam_merge_kernel: one warp per image merges
// the K*K row candidates.  The top-K of the union of per-row top-Ks under one total order IS the
// global top-K, so the result is identical to a single pass over all K*V candidates -- which is
// what the first version did with one CTA per image and took 118 us per step (ncu launch list,
// profiles/): 128 CTAs cannot scan 6.4 M candidates quickly.
__device__ __forceinline__ bool cand_better(float v, int i, float bv, int bi) {
  return v > bv || (v == bv && i < bi);
}

template <bool FAST>
__device__ __forceinline__ float exp_sel(float x) { return FAST ? __expf(x) : expf(x); }

// FAST (bf16 mode): ex2.approx-based exponentials in the row log-sum-exp; the fp32 parity mode keeps expf.
template <int K, bool FAST>
__global__ void __launch_bounds__(256) beam_row_topk_kernel(const float* __restrict__ scores,
                                                            const uint8_t* __restrict__ finished,
                                                            const float* __restrict__ logits,
                                                            const float* __restrict__ lse_in,
                                                            float* __restrict__ lse_out, int V, int end_id,
                                                            float* __restrict__ cand_v, int* __restrict__ cand_i,
                                                            int staged, TraceRec* trace_buf) {
  __shared__ float scratch[64];
  __shared__ float wv[8];
  __shared__ int wi[8];
  __shared__ int s_win;
  Trace trace(trace_buf);
  pdl_wait();
  pdl_trigger();
  trace.mark();
  const int r = blockIdx.x;            // global row = b*K + j
  const int j = r % K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* lg = logits + (size_t)r * V;
  // Stage the logits row in shared memory with 16-byte loads, four in flight per thread, and take the
  // row's log-sum-exp on the way in (online max / sum per thread, one block reduction of the pairs);
  // rows too long for the 47 KB window are read from global memory.
  extern __shared__ __align__(16) float row_s[];
  float ls;
  if (staged && !lse_in) {
    float m = -INFINITY, ssum = 0.f;
    // online max / sum in batches: one rescale per batch of up to 16 values, no per-element branch
    auto push4 = [&](const float4* x, int n) {
      float bm = m;
      for (int u = 0; u < n; ++u) bm = fmaxf(fmaxf(bm, fmaxf(x[u].x, x[u].y)), fmaxf(x[u].z, x[u].w));
      if (bm > m) { ssum *= exp_sel<FAST>(m - bm); m = bm; }          // exp(-inf) = 0 on the first batch
      for (int u = 0; u < n; ++u)
        ssum += (exp_sel<FAST>(x[u].x - m) + exp_sel<FAST>(x[u].y - m)) + (exp_sel<FAST>(x[u].z - m) + exp_sel<FAST>(x[u].w - m));
    };
    if ((V & 3) == 0 && (reinterpret_cast<uintptr_t>(lg) & 15) == 0) {
      const float4* src4 = reinterpret_cast<const float4*>(lg);
      float4* dst4 = reinterpret_cast<float4*>(row_s);
      const int n4 = V / 4;
      int i = tid;
      for (; i + 3 * 256 < n4; i += 4 * 256) {
        float4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = __ldg(src4 + i + u * 256);
#pragma unroll
        for (int u = 0; u < 4; ++u) dst4[i + u * 256] = x[u];
        push4(x, 4);
      }
      for (; i < n4; i += 256) {
        const float4 x = __ldg(src4 + i);
        dst4[i] = x;
        push4(&x, 1);
      }
    } else {
      for (int v = tid; v < V; v += 256) {
        const float x = lg[v];
        row_s[v] = x;
        if (x > m) { ssum = ssum * exp_sel<FAST>(m - x) + 1.f; m = x; }
        else ssum += exp_sel<FAST>(x - m);
      }
    }
    // combine (m, s) pairs: warp shuffles, then the 8 warp results
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, o);
      const float os = __shfl_xor_sync(0xffffffffu, ssum, o);
      const float nm = fmaxf(m, om);
      ssum = (nm == -INFINITY) ? 0.f : ssum * exp_sel<FAST>(m - nm) + os * exp_sel<FAST>(om - nm);
      m = nm;
    }
    if (lane == 0) { scratch[warp] = m; scratch[8 + warp] = ssum; }
    __syncthreads();                       // also publishes row_s
    float gm = scratch[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) gm = fmaxf(gm, scratch[w]);
    float gs = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) gs += (scratch[w] == -INFINITY) ? 0.f : scratch[8 + w] * exp_sel<FAST>(scratch[w] - gm);
    ls = gm + logf(gs);
    lg = row_s;
  } else {
    if (staged) {
      for (int v = tid; v < V; v += 256) row_s[v] = lg[v];
      __syncthreads();
      lg = row_s;
    }
    if (lse_in) {
      ls = lse_in[r];
    } else {
      float m = -INFINITY;
      for (int v = tid; v < V; v += 256) m = fmaxf(m, lg[v]);
      m = block_max(m, scratch);
      float s2 = 0.f;
      for (int v = tid; v < V; v += 256) s2 += expf(lg[v] - m);
      s2 = block_sum(s2, scratch);
      ls = m + logf(s2);
    }
  }
  if (tid == 0 && lse_out) lse_out[r] = ls;
  const float sc = scores[r];
  const bool fin = finished[r] != 0;

  // candidate values, in place in shared memory when the row is staged; the same pass finds each
  // thread's best candidate (the first "rescan")
  float tv = INFINITY;     // last candidate taken from this thread's slice
  int ti = -1;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  float b2v = -INFINITY;   // runner-up of the thread's slice: the first win needs no rescan
  int b2i = 0x7fffffff;
  if (staged) {
    for (int v = tid; v < V; v += 256) {
      float c;
      if (fin) c = (v == end_id) ? sc : -INFINITY;
      else c = __fadd_rn(sc, __fsub_rn(row_s[v], ls));
      row_s[v] = c;                         // a thread only ever re-reads its own slice: no barrier needed
      if (bi == 0x7fffffff || c > bv) { b2v = bv; b2i = bi; bv = c; bi = v; }
      else if (b2i == 0x7fffffff || c > b2v) { b2v = c; b2i = v; }
    }
  }
  auto cand = [&](int v) -> float {
    if (staged) return row_s[v];
    if (fin) return (v == end_id) ? sc : -INFINITY;
    return __fadd_rn(sc, __fsub_rn(lg[v], ls));
  };
  // Selection by K rounds of block arg-best.  Every thread caches the best not-yet-taken candidate
  // of its own strided slice; only the thread that wins a round rescans its slice (the first
  // version kept a sorted K-list per thread and was instruction bound on the insertions).
  // Order: value descending, index ascending; a thread scans v ascending, so strict '>' keeps the
  // lowest index among equal values.
  auto rescan = [&]() {
    bv = -INFINITY;
    bi = 0x7fffffff;
    for (int v = tid; v < V; v += 256) {
      const float c = cand(v);
      const bool elig = (c < tv) || (c == tv && v > ti);
      if (elig && (bi == 0x7fffffff || c > bv)) { bv = c; bi = v; }
    }
  };
  if (!staged) rescan();
  for (int rd = 0; rd < K; ++rd) {
    float wvv = bv;
    int wii = bi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, wvv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, wii, o);
      if (cand_better(ov, oi, wvv, wii)) { wvv = ov; wii = oi; }
    }
    if (lane == 0) { wv[warp] = wvv; wi[warp] = wii; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < 8; ++w)
        if (cand_better(wv[w], wi[w], wvv, wii)) { wvv = wv[w]; wii = wi[w]; }
      s_win = wii;
      cand_v[(size_t)r * K + rd] = wvv;
      cand_i[(size_t)r * K + rd] = j * V + wii;
    }
    __syncthreads();
    if (bi == s_win && bi != 0x7fffffff) {   // this thread's candidate was taken: find its next one
      tv = bv;
      ti = bi;
      if (b2i != 0x7fffffff) { bv = b2v; bi = b2i; b2i = 0x7fffffff; }    // cached runner-up
      else rescan();
    }
    __syncthreads();
  }
  trace.end(TK_BEAM_TOPK);
}

// one warp per image: merge the K sorted row lists (K*K <= 64 candidates, two per lane)
template <int K>
__global__ void __launch_bounds__(128) beam_merge_kernel(const float* __restrict__ cand_v,
                                                         const int* __restrict__ cand_i,
                                                         const uint8_t* __restrict__ finished, int B, int V,
                                                         int end_id, float* __restrict__ new_scores,
                                                         int32_t* __restrict__ back, int32_t* __restrict__ tok,
                                                         uint8_t* __restrict__ new_finished, TraceRec* trace_buf) {
  Trace trace(trace_buf);
  pdl_wait();
  pdl_trigger();
  trace.mark();
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) { trace.end(TK_BEAM_MERGE); return; }
  constexpr int N = K * K;
  float v0 = -INFINITY, v1 = -INFINITY;
  int i0 = 0x7fffffff, i1 = 0x7fffffff;
  if (lane < N) { v0 = cand_v[(size_t)b * N + lane]; i0 = cand_i[(size_t)b * N + lane]; }
  if (lane + 32 < N) { v1 = cand_v[(size_t)b * N + lane + 32]; i1 = cand_i[(size_t)b * N + lane + 32]; }
  for (int rd = 0; rd < K; ++rd) {
    float bv = v0;
    int bi = i0;
    if (cand_better(v1, i1, bv, bi)) { bv = v1; bi = i1; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) {
      const int jb = bi / V, tk = bi - jb * V;
      new_scores[b * K + rd] = bv;
      back[b * K + rd] = jb;
      tok[b * K + rd] = tk;
      new_finished[b * K + rd] = (uint8_t)((finished[b * K + jb] != 0) || (tk == end_id));
    }
    if (i0 == bi && bi != 0x7fffffff) { v0 = -INFINITY; i0 = 0x7fffffff; }
    if (i1 == bi && bi != 0x7fffffff) { v1 = -INFINITY; i1 = 0x7fffffff; }
  }
  trace.end(TK_BEAM_MERGE);
}

inline size_t beam_select_workspace_bytes(int B, int K) { return (size_t)B * K * K * (sizeof(float) + sizeof(int)); }

// lse_in != null: use the given row log-sum-exps; else compute them (and write lse_out if given)
inline int launch_beam_select(const float* scores, const uint8_t* finished, const float* logits,
                              const float* lse_in, float* lse_out, int B, int K, int V, int end_id,
                              void* workspace, float* new_scores, int32_t* back, int32_t* tok,
                              uint8_t* new_finished, cudaStream_t st, bool fast = false) {
  if (B <= 0) return 0;
  if (K > V) DIC_FAIL(-4, "beam %d larger than vocabulary %d", K, V);
  float* cv = reinterpret_cast<float*>(workspace);
  int* ci = reinterpret_cast<int*>(cv + (size_t)B * K * K);
  const int staged = (sizeof(float) * (size_t)V <= 47 * 1024) ? 1 : 0;
  ProfScope prof(P_BEAM_SELECT, st, (double)B * K * V * sizeof(float));
#define DIC_TOPK_CASE(KK)                                                                                  \
  case KK:                                                                                                 \
    if (fast)                                                                                              \
      DIC_CUDA(launch_pdl(beam_row_topk_kernel<KK, true>, dim3(B * KK), dim3(256), staged ? sizeof(float) * V : 0, \
                          st, scores, finished, logits, lse_in, lse_out, V, end_id, cv, ci, staged, g_trace_host)); \
    else                                                                                                   \
      DIC_CUDA(launch_pdl(beam_row_topk_kernel<KK, false>, dim3(B * KK), dim3(256), staged ? sizeof(float) * V : 0, \
                          st, scores, finished, logits, lse_in, lse_out, V, end_id, cv, ci, staged, g_trace_host));         \
    DIC_LAUNCH_CHECK();                                                                                    \
    DIC_CUDA(launch_pdl(beam_merge_kernel<KK>, dim3(cdiv(B, 4)), dim3(128), 0, st, (const float*)cv,       \
                        (const int*)ci, finished, B, V, end_id, new_scores, back, tok, new_finished,     \
                        g_trace_host));     \
    break;
  switch (K) {
    DIC_TOPK_CASE(1) DIC_TOPK_CASE(2) DIC_TOPK_CASE(3) DIC_TOPK_CASE(4)
    DIC_TOPK_CASE(5) DIC_TOPK_CASE(6) DIC_TOPK_CASE(7) DIC_TOPK_CASE(8)
    default: DIC_FAIL(-4, "beam size %d not in 1..%d", K, DIC_MAX_BEAM);
  }
#undef DIC_TOPK_CASE
  DIC_LAUNCH_CHECK();
  return 0;
}

// ---- beam reorder: gather (h,c) by backpointer, embed the chosen tokens ------------------------------
template <typename ST>
__global__ void __launch_bounds__(256) beam_reorder_kernel(const ST* __restrict__ h_tmp,
                                                           const float* __restrict__ c_tmp,
                                                           const int32_t* __restrict__ back,
                                                           const int32_t* __restrict__ tok,
                                                           const ST* __restrict__ emb,
                                                           ST* __restrict__ Xnext, long long x_row, int col_h,
                                                           float* __restrict__ c, int rows, int K, int E,
                                                           int H, TraceRec* trace_buf) {
  Trace trace(trace_buf);
  pdl_wait();
  pdl_trigger();
  trace.mark();
  const int W = E + H;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < rows * W; i += gridDim.x * 256) {
    const int r = i / W, q = i - r * W;
    if (q < E) {
      Xnext[(size_t)r * x_row + q] = emb[(size_t)tok[r] * E + q];
    } else {
      const int j = q - E;
      const int src = (r / K) * K + back[r];
      Xnext[(size_t)r * x_row + col_h + j] = h_tmp[(size_t)src * H + j];
      c[(size_t)r * H + j] = c_tmp[(size_t)src * H + j];
    }
  }
  trace.end(TK_BEAM_REORDER);
}

// ---- final backtrack from row 0 (top-k output is sorted, so row 0 is the best hypothesis) --------
__global__ void __launch_bounds__(128) beam_backtrack_kernel(const int32_t* __restrict__ back,
                                                             const int32_t* __restrict__ tok,
                                                             const float* __restrict__ final_scores, int B,
                                                             int K, int T, int end_id,
                                                             int64_t* __restrict__ tokens,
                                                             int32_t* __restrict__ lengths,
                                                             float* __restrict__ scores) {
  const int b = blockIdx.x * 128 + threadIdx.x;
  if (b >= B) return;
  int row = 0;
  for (int t = T - 1; t >= 0; --t) {
    const size_t o = ((size_t)t * B + b) * K + row;
    tokens[(size_t)b * T + t] = tok[o];
    row = back[o];
  }
  int len = T;
  for (int t = 0; t < T; ++t)
    if (tokens[(size_t)b * T + t] == end_id) { len = t + 1; break; }
  lengths[b] = len;
  scores[b] = final_scores[b * K];
}

__global__ void __launch_bounds__(256) beam_state_init_kernel(float* scores, uint8_t* finished, int B, int K) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= B * K) return;
  scores[i] = (i % K == 0) ? 0.f : -INFINITY;
  finished[i] = 0;
}

// LSTM cell of the look-ahead beam step (dic_api.cu decode_impl).  The gate GEMM ran on the rows of step t-1 in
// PARENT order ([beta.z | h] . [W_z | W_hh]^T, split-K partials); row r of step t continues parent back[r] of its
// image with token tok[r]:
//     gates[r] = sum_s part[s][parent] + etab[tok[r]] + (b_ih + b_hh),   c_prev = c_par[parent]
// so the beam reorder is three index loads here instead of a copy of every row.  h' goes to the h columns of the next
// parent-order operand, c' to the next parent-order cell buffer.
struct LstmBeamArgs {
  const float* gate_part;   // [splits][rows_alloc][4H]
  long long part_stride;
  int splits;
  const float* etab;        // [V, 4H]
  const float* bias_g;      // [4H]
  const int32_t* back;      // [rows] parent inside the image
  const int32_t* tok;       // [rows]
  const float* c_par;       // [rows, H] parent order
  float* c_out;             // [rows, H]
  void* h_out;              // ST, row r at h_out + r*h_stride
  long long h_stride;
  int rows, H, K;
  TraceRec* trace;
};

template <typename ST>
__global__ void __launch_bounds__(256) lstm_beam_kernel(const LstmBeamArgs p) {
  Trace trace(p.trace);
  const int idx = blockIdx.x * 256 + threadIdx.x;
  const bool live = idx < p.rows * p.H;
  const int H = p.H;
  const int r = live ? idx / H : 0, j = live ? idx - r * H : 0;
  // backpointers, tokens and the parents' cell states are three launches old (selection -> context -> gate GEMM):
  // complete before this grid may start, loaded before the wait
  int parent = 0;
  float c_prev = 0.f, g4[4] = {0.f, 0.f, 0.f, 0.f};
  if (live) {
    parent = (r / p.K) * p.K + p.back[r];
    const int tk = p.tok[r];
    c_prev = p.c_par[(size_t)parent * H + j];
#pragma unroll
    for (int q = 0; q < 4; ++q) g4[q] = p.bias_g[q * H + j] + p.etab[(size_t)tk * 4 * H + q * H + j];
  }
  pdl_wait();
  pdl_trigger();
  trace.mark();
  if (!live) { trace.end(TK_LSTM_FWD); return; }
  constexpr int kMaxSplits = 16;
  float part[4][kMaxSplits];
#pragma unroll
  for (int sp = 0; sp < kMaxSplits; ++sp) {
#pragma unroll
    for (int q = 0; q < 4; ++q)
      part[q][sp] = sp < p.splits ? p.gate_part[(size_t)sp * p.part_stride + (size_t)parent * 4 * H + q * H + j] : 0.f;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float s = g4[q];
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
  reinterpret_cast<ST*>(p.h_out)[(size_t)r * p.h_stride + j] = from_f<ST>(h);
  trace.end(TK_LSTM_FWD);
}

}  // namespace dic
