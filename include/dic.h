/*
 * dic.h -- C ABI of the B200-native (sm_100a) depth-image-captioning decoder path.
 *
 * The reference (Kyo-suke-S/Depth_image_captioning_pub) is pure Python/PyTorch and has
 * no FFI layer; the drop-in boundary is its nn.Module surface.  This
 * header is the layer BELOW that surface: the host-side Python modules in
 * depth_image_captioning_pub_b200/ keep the reference's class names, constructor
 * arguments, state_dict keys and forward/sample/batch_sample signatures and call these
 * entry points through ctypes.  Each entry point cites the reference code it replaces
 * (paths relative to the reference checkout, Captioning_models/...).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer borrowed from the caller (torch tensor.data_ptr())
 *     unless its name starts with host_;
 *   - the caller allocates all outputs and the workspace (dic_*_workspace_bytes);
 *     the library never allocates device memory, never frees, never synchronises;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - return value 0 = ok, negative = error; dic_last_error() gives a thread-local message;
 *   - there is NO CPU fallback: without a CUDA device every compute call fails.
 *
 * Storage modes (dic_dtype): DIC_F32 keeps annotations, projections and GEMM operands in
 * fp32 (CUDA-core FMA GEMMs, the parity mode: logits within 1e-4, alpha within 1e-5 of the
 * reference); DIC_BF16 stores annotations / att1 / GEMM operands in bf16 with fp32
 * accumulation and fp32 recurrent state (tcgen05 tensor-core GEMMs; logits within 2e-2).
 */
#ifndef DIC_H_
#define DIC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIC_VERSION 100

enum dic_dtype { DIC_F32 = 0, DIC_BF16 = 1 };

/* attention variants (attention.py) */
enum dic_attn_mode {
  DIC_ATTN_SOFT = 0,           /* Soft_Attention.forward            attention.py:81-95   */
  DIC_ATTN_GUMBEL_SOFTMAX = 1, /* Hard_Attention.forward            attention.py:132-148 */
  DIC_ATTN_GUMBEL_MAX = 2      /* Hard_Attention.Hard_sample        attention.py:150-167 */
};

#define DIC_MAX_STEPS 128 /* longest decoder sequence (reference: max_length 30) */
#define DIC_MAX_BEAM 8

/* Problem dimensions (config.py:11-15 defaults: L=196 D=2048 A=E=H=128).
 * Constraints: D % 8 == 0, A % 4 == 0, E % 4 == 0, H % 4 == 0, A <= 1024, L <= 1024. */
typedef struct dic_dims {
  int32_t L; /* annotation vectors per image (14*14)        */
  int32_t D; /* annotation channels        dim_encoder      */
  int32_t A; /* attention dim              dim_attention    */
  int32_t E; /* embedding dim              dim_embedding    */
  int32_t H; /* LSTM hidden dim            dim_decoder      */
  int32_t V; /* vocabulary size                              */
} dic_dims;

/* The 17 fp32 parameter tensors of a reference decoder, in state_dict layout
 * (depth_models.py:106-135; identical for soft/hard and base/depth decoders). */
typedef struct dic_params {
  float* enc_att_w;  /* attention.encoder_att.weight [A,D]   */
  float* enc_att_b;  /* attention.encoder_att.bias   [A]     */
  float* dec_att_w;  /* attention.decoder_att.weight [A,H]   */
  float* dec_att_b;  /* attention.decoder_att.bias   [A]     */
  float* full_att_w; /* attention.full_att.weight    [1,A]   */
  float* full_att_b; /* attention.full_att.bias      [1]     */
  float* embed_w;    /* embed.weight                 [V,E]   */
  float* w_ih;       /* decode_step.weight_ih        [4H,E+D] gate order i,f,g,o */
  float* w_hh;       /* decode_step.weight_hh        [4H,H]  */
  float* b_ih;       /* decode_step.bias_ih          [4H]    */
  float* b_hh;       /* decode_step.bias_hh          [4H]    */
  float* init_w;     /* init_linear.weight           [2H,D]  */
  float* init_b;     /* init_linear.bias             [2H]    */
  float* fbeta_w;    /* f_beta.weight                [D,H]   */
  float* fbeta_b;    /* f_beta.bias                  [D]     */
  float* lin_w;      /* linear.weight                [V,H]   */
  float* lin_b;      /* linear.bias                  [V]     */
} dic_params;

int dic_version(void);
const char* dic_last_error(void);

/* ---- weight pack -------------------------------------------------------------------
 * Compute-layout copy of the parameters (storage dtype, fused/concatenated operands:
 * [W_dec;W_beta], [W_ih|W_hh], b_ih+b_hh, ...).  Rebuild after every optimizer step. */
size_t dic_pack_bytes(const dic_dims* dims, int dtype);
int dic_pack_weights(const dic_dims* dims, int dtype, const dic_params* params, void* pack,
                     void* stream);

/* ---- teacher-forced training forward / backward ---------------------------------------
 * Replaces CD_RNNDecoderWithSoftAttention.forward (depth_models.py:153-207),
 * RNNDecoderWithSoftAttention.forward (base_caption_models.py:105-156),
 * CD_/RNNDecoderWithHardAttention.forward (depth_models.py:580-634) and
 * .eval_forward (depth_models.py:637-689).
 *
 *   f_rgb, f_depth : [B,L,D] annotations, feat_dtype (DIC_F32|DIC_BF16); f_depth may be NULL
 *                    (base decoders).  RGB + depth are added once (depth_models.py:163).
 *   captions       : [B,cap_stride] int64, column 0 = <start> (depth_models.py:160)
 *   host_batch_sizes[T] : bs_valid per step, non-increasing (depth_models.py:182); the
 *                    batch must be sorted by length, descending (util.py:95)
 *   u              : [sum(bs), L] fp32 uniform draws for the hard variants, packed
 *                    time-major like the reference's per-step torch.rand (attention.py:17,40)
 *   dropout_mask   : [sum(bs), H] fp32, already scaled by 1/(1-p), or NULL (eval mode)
 *   logits         : out [sum(bs), V] fp32 = PackedSequence.data (depth_models.py:204)
 *   alphas         : out [B,T,L] fp32; the CALLER zero-fills it (depth_models.py:176).
 *                    Required (it is also the state saved for backward).
 *   workspace      : dic_train_workspace_bytes(); holds the saved state dic_decoder_backward reads.
 */
size_t dic_train_workspace_bytes(const dic_dims* dims, int dtype, int B, int T);

int dic_decoder_forward(const dic_dims* dims, int dtype, int attn_mode, const void* pack,
                        const void* f_rgb, const void* f_depth, int feat_dtype,
                        const int64_t* captions, int cap_stride, const int32_t* host_batch_sizes,
                        int T, int B, const float* u, float temp, const float* dropout_mask,
                        float* logits, float* alphas, void* workspace, size_t workspace_bytes,
                        void* stream);

/* Backward of dic_decoder_forward (implicit autograd in the reference).
 *   d_logits : [sum(bs), V] fp32 gradient of PackedSequence.data
 *   d_alphas : [B,T,L] fp32 gradient of alphas, or NULL
 *   grads    : 17 fp32 buffers in dic_params layout, OVERWRITTEN with the gradients
 *   d_feats  : out [B,L,D] in feat_dtype = dL/d(f_rgb + f_depth) (same tensor for both inputs), or NULL
 *   f_rgb, f_depth, feat_dtype : the same annotation tensors the forward call received
 *              (when no fused copy was needed the forward kept reading the caller's tensor)
 */
int dic_decoder_backward(const dic_dims* dims, int dtype, int attn_mode, const void* pack,
                         const void* f_rgb, const void* f_depth, int feat_dtype,
                         const int64_t* captions, int cap_stride, const int32_t* host_batch_sizes,
                         int T, int B, const float* d_logits, const float* d_alphas,
                         const float* alphas, float temp, const float* dropout_mask,
                         const dic_params* grads, void* d_feats, void* workspace,
                         size_t workspace_bytes, void* stream);

/* Same as dic_decoder_forward, with the logits written as float32 or, in bf16 mode, bfloat16
 * (logits_dtype): the fused training step keeps them in bf16 and the loss head overwrites them
 * with d_logits in place, which halves the traffic of the [sum(bs), V] block. */
int dic_decoder_forward_ex(const dic_dims* dims, int dtype, int attn_mode, const void* pack,
                           const void* f_rgb, const void* f_depth, int feat_dtype,
                           const int64_t* captions, int cap_stride, const int32_t* host_batch_sizes,
                           int T, int B, const float* u, float temp, const float* dropout_mask,
                           void* logits, int logits_dtype, float* alphas, void* workspace,
                           size_t workspace_bytes, void* stream);

/* Same as dic_decoder_backward, with d_logits in either float32 or the storage dtype of the mode
 * (d_logits_dtype = DIC_F32 / DIC_BF16): the fused loss head below writes bf16 d_logits directly,
 * which saves the fp32 -> bf16 operand copy of the tensor-core GEMMs. */
int dic_decoder_backward_ex(const dic_dims* dims, int dtype, int attn_mode, const void* pack,
                            const void* f_rgb, const void* f_depth, int feat_dtype,
                            const int64_t* captions, int cap_stride, const int32_t* host_batch_sizes,
                            int T, int B, const void* d_logits, int d_logits_dtype,
                            const float* d_alphas, const float* alphas, float temp,
                            const float* dropout_mask, const dic_params* grads, void* d_feats,
                            void* workspace, size_t workspace_bytes, void* stream);

/* One-shot hook for data-parallel training: the NEXT dic_decoder_backward(_ex) call of this process
 * (any host thread: PyTorch runs backward on an autograd worker thread) records `event` (a cudaEvent_t) on its stream once all 17 parameter gradients are enqueued, i.e.
 * before the dL/dF GEMM, so the caller's gradient all-reduce can overlap that last kernel.  NULL clears it. */
void dic_set_grads_ready_event(void* event);
/* Bucketed form of the same hook (any of the three may be NULL).  The gradients live in one flat buffer in
 * dic_params order; the next backward records
 *   ev_linear  once lin_w / lin_b (the LAST two tensors) are final -- before the backward time loop starts,
 *   ev_middle  once everything but enc_att_w / enc_att_b (the FIRST two tensors) is final,
 *   ev_all     once all 17 are final (before the dL/dF GEMM, which then leaves some SMs to the all-reduce). */
void dic_set_grads_ready_events(void* ev_linear, void* ev_middle, void* ev_all);

/* ---- depth CNN encoder -----------------------------------------------------------------------
 * Replaces Depth_CNN_endoder.forward (depth_models.py:12-56) and its autograd backward: conv 1->128 k7 s3, BN,
 * ReLU, maxpool 3, conv 128->512 k3, BN, ReLU, maxpool 3, conv 512->2048 k1, BN, ReLU, AdaptiveAvgPool2d(14),
 * permute -> annotations [B, 196, 2048] written directly in the decoder's layout.  depth_imgs: [B, Hi, Wi] fp32
 * (one channel; the last feature map must be 7 x 7, i.e. Hi = Wi = 224..232).  All 18 tensors fp32 in the module's
 * state_dict layout; running_mean / running_var are updated in place when training != 0 (BatchNorm2d semantics:
 * batch statistics, momentum, unbiased running variance) and used for normalisation otherwise.
 * feats: out [B, 196, 2048] float32 (feat_dtype = DIC_F32) or bfloat16.  The workspace keeps the activations the
 * backward needs; dic_depth_encoder_backward takes d_feats [B,196,2048] (dL/dF from dic_decoder_backward) and
 * writes the 12 parameter gradients (grads: same struct, running_* ignored). */
typedef struct dic_enc_params {
  float* conv1_w; float* conv1_b; float* bn1_w; float* bn1_b; float* bn1_mean; float* bn1_var;   /* [128,1,7,7] ... */
  float* conv2_w; float* conv2_b; float* bn2_w; float* bn2_b; float* bn2_mean; float* bn2_var;   /* [512,128,3,3] ... */
  float* conv3_w; float* conv3_b; float* bn3_w; float* bn3_b; float* bn3_mean; float* bn3_var;   /* [2048,512,1,1] ... */
} dic_enc_params;
size_t dic_depth_encoder_workspace_bytes(int B, int Hi, int Wi, int dtype);
int dic_depth_encoder_forward(int dtype, int training, int B, int Hi, int Wi, const float* depth_imgs,
                              const dic_enc_params* params, float momentum, float eps, void* feats, int feat_dtype,
                              void* workspace, size_t workspace_bytes, void* stream);
int dic_depth_encoder_backward(int dtype, int B, int Hi, int Wi, const dic_enc_params* params, const void* d_feats,
                               int feat_dtype, const dic_enc_params* grads, void* workspace, size_t workspace_bytes,
                               void* stream);

/* ---- data-parallel gradient all-reduce over NVLink peer memory ------------------------------
 * The one exchange step of the path.  The reference has no distributed code (config.py:68 pins 'cuda:0'); this
 * replaces the NCCL all-reduce a DistributedDataParallel wrapper would issue after depth_train.py:219 (loss.backward()).
 * Every rank passes the SAME tables: bufs[r] / flags[r] = rank r's flat fp32 gradient buffer / flag block
 * (dic_dp_flag_bytes() bytes, zero before the first call) as mapped into THIS process (peer-visible memory, e.g.
 * torch.distributed._symmetric_memory buffer_ptrs); multicast = NVLS multicast mapping of the buffers or NULL.
 * In place: buf <- scale * sum_r buf_r on every rank, summed in rank order (bit-identical on all ranks).
 * n_floats: a multiple of 4*world.  epoch: 1, 2, 3, ... the same on every rank for the same call.  blocks: CTAs
 * to use (<= 148; 0 = default).  One kernel on `stream`; no host synchronisation. */
size_t dic_dp_flag_bytes(void);
int dic_dp_allreduce(int world, int rank, void* const* bufs, void* const* flags, void* multicast,
                     long long n_floats, float scale, unsigned int epoch, int blocks, void* stream);

/* ---- fused caption-loss head ----------------------------------------------
 * Replaces the caller-side loss of the training loop (depth_train.py:210-216 / :530-532):
 *   loss = cross_entropy(packed logits, packed targets, ignore_index, mean over non-ignored)
 *        + lam * mean_{b,l} (1 - sum_t alphas[b,t,l])^2       (lam = 0 / alphas NULL: CE only)
 * and returns its gradients for an upstream gradient of 1: d_logits [N,V] in the storage dtype
 * of the mode (may alias `logits` when those have the same dtype), d_alphas [B,T,L] fp32 (may
 * be NULL).  logits: float32, or (bf16 mode, dic_decoder_forward_ex) bfloat16.
 * Targets are read from `captions` (target of packed row (t,b) = captions[b,t+1]).
 * loss: [1] fp32 on the device.  No host synchronisation. */
size_t dic_caption_loss_workspace_bytes(int N, int B);
int dic_caption_loss(const dic_dims* dims, int dtype, const void* logits, int logits_dtype,
                     const int64_t* captions, int cap_stride, const int32_t* host_batch_sizes, int T,
                     int B, int ignore_index, const float* alphas, float lam, float* loss,
                     void* d_logits, float* d_alphas, void* workspace, size_t workspace_bytes,
                     void* stream);
/* d_logits, d_alphas *= grad_loss[0] (device scalar); a no-op kernel when it is exactly 1. */
int dic_scale_loss_grads(int dtype, const float* grad_loss, void* d_logits, size_t n_logits,
                         float* d_alphas, size_t n_alphas, void* stream);

/* ---- decoding -----------------------------------------------------------------------
 * dic_decode_greedy replaces .sample / .batch_sample (depth_models.py:216-305, 698-789):
 * fixed max_len steps, no early stop, next token = argmax (ties -> lowest id); no host
 * sync per step (the reference copies to the host every step, depth_models.py:298-299).
 *   tokens : out [B,max_len] int64;  alphas_out : out [max_len,B,L] fp32 or NULL;
 *   logits_out : out [max_len,B,V] fp32 or NULL;  u : [max_len*B, L] for DIC_ATTN_GUMBEL_MAX.
 *
 * dic_decode_beam is NOT in the reference (greedy only); semantics are this build's own
 * specification (oracle/decoder_oracle.py:beam_search).
 *   tokens [B,max_len] int64 (best row, <end>-padded), lengths [B] int32, scores [B] fp32;
 *   optional traces: back/toks [max_len,B,beam] int32, step_scores/lse [max_len,B,beam] fp32,
 *   logits_out [max_len,B*beam,V] fp32.
 *
 * Kernel order (bf16 mode, the shapes of the reference model; DESIGN.md section 4 "look-ahead attention"): the
 * attention of step t+1 is computed from h_t before the token / beam selection of step t has finished, the context
 * pass running next to the selection kernel; beam rows are never physically reordered (the LSTM kernel follows the
 * backpointers).  Results are those of the serial order (DIC_BEAM_LOOKAHEAD=0) up to fp32 summation order.
 */
size_t dic_decode_workspace_bytes(const dic_dims* dims, int dtype, int B, int beam);

int dic_decode_greedy(const dic_dims* dims, int dtype, int attn_mode, const void* pack,
                      const void* f_rgb, const void* f_depth, int feat_dtype, int B,
                      int start_id, int max_len, const float* u, int64_t* tokens,
                      float* alphas_out, float* logits_out, void* workspace,
                      size_t workspace_bytes, void* stream);

int dic_decode_beam(const dic_dims* dims, int dtype, const void* pack, const void* f_rgb,
                    const void* f_depth, int feat_dtype, int B, int beam, int start_id,
                    int end_id, int max_len, int64_t* tokens, int32_t* lengths, float* scores,
                    int32_t* back, int32_t* toks, float* step_scores, float* lse,
                    float* logits_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- step-level operators ------------------------------------------------------------
 * dic_attention_forward replaces Soft_Attention.forward / Hard_Attention.forward /
 * Hard_Attention.Hard_sample as standalone modules (attention.py:81-95,132-167):
 * feats [B,L,D] fp32, h [B,H] fp32 -> context [B,D] fp32, alpha [B,L] fp32.
 * Weights are raw fp32 (enc_w [A,D], enc_b [A], dec_w [A,H], dec_b [A], full_w [A], full_b [1]).
 */
size_t dic_attention_workspace_bytes(const dic_dims* dims, int dtype, int B);
int dic_attention_forward(const dic_dims* dims, int dtype, int attn_mode, const float* enc_w,
                          const float* enc_b, const float* dec_w, const float* dec_b,
                          const float* full_w, const float* full_b, const float* feats,
                          const float* h, int B, const float* u, float temp, float* context,
                          float* alpha, void* workspace, size_t workspace_bytes, void* stream);

/* One beam selection (the integer part of beam search, bit-exact vs the oracle's
 * beam_select given identical inputs): scores [B,K] fp32, finished [B,K] uint8,
 * logits [B*K,V] fp32, lse [B*K] fp32 -> new_scores [B,K], back [B,K] int32, tok [B,K] int32,
 * new_finished [B,K] uint8.  cand = scores + (logits - lse), stable top-K, ties -> lowest index. */
size_t dic_beam_select_workspace_bytes(int B, int K);
int dic_beam_select(const float* scores, const uint8_t* finished, const float* logits,
                    const float* lse, int B, int K, int V, int end_id, float* new_scores,
                    int32_t* back, int32_t* tok, uint8_t* new_finished, void* workspace,
                    size_t workspace_bytes, void* stream);

/* Row-wise log-sum-exp of logits [R,V] fp32 -> lse [R] fp32. */
int dic_row_lse(const float* logits, int R, int V, float* lse, void* stream);

/* ---- instrumentation (bench.py) ---------------------------------------------------------
 * dic_launch_count: kernels launched by this library since load.  dic_profile_*: optional
 * CUDA-event timing per kernel class on the launching stream (off by default). */
long long dic_launch_count(void);
int dic_profile_classes(void);
const char* dic_profile_class_name(int cls);
void dic_profile_enable(int on);
int dic_profile_read(float* ms, long long* launches, double* bytes);

/* Number of image sub-batches (each on its own library-owned stream, forked from and joined back
 * into the caller's stream with events) the time loops of forward / backward / decode run over.
 * 0 (default) = chosen from the batch size; 1 = everything on the caller's stream. */
void dic_set_substreams(int n);

/* Debug timeline: while a trace is active, thread 0 of every CTA of the step kernels appends a
 * 32-byte record {u64 t_entry, t_after_dependency_wait, t_exit (globaltimer ns); i32 kernel id,
 * linear block id} to buf (device memory, capacity_records records).  Both calls synchronise the
 * device.  dic_trace_stop returns the number of records produced (may exceed the capacity). */
int dic_trace_start(void* buf, unsigned int capacity_records);
int dic_trace_stop(unsigned int* count);

/* Fused multi-tensor AdamW step (replaces torch.optim.AdamW.step of
 * depth_train.py:136-137,221 for fp32 CUDA tensors): decoupled weight decay, bias-corrected
 * moments, no amsgrad.  n tensors; params / grads / exp_avg / exp_avg_sq are HOST arrays of n
 * device pointers, sizes[i] = element count; step counts from 1. */
int dic_adamw_step(int n, float* const* params, const float* const* grads, float* const* exp_avg,
                   float* const* exp_avg_sq, const long long* sizes, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, void* stream);

/* Test hook of the fused dL/dF GEMM (bf16 mode, csrc/dfeat_tc.cuh):
 *   dF[b] (L x D, bf16) = datt1[b] (L x A) . w_enc (A x D) + alpha16[b]^T (L x T) . dz[b] (T x D) + dmeanF[b] / L
 * datt1 [B*L, A] bf16; w_enc [A, D] bf16; alpha16 [B*T, Lp] bf16 (Lp % 8 == 0, columns >= L zero);
 * dz [T, B, D] bf16; dmeanF [B, D] fp32. */
int dic_dfeat_gemm(const void* datt1, const void* w_enc, const void* alpha16, int Lp, const void* dz,
                   const float* dmeanF, void* dF, int B, int L, int D, int A, int T, void* stream);

/* GEMM test hook: C[M,N] (fp32) = A[M,K] . B[N,K]^T (+bias[N]); a_dtype/b_dtype storage.
 * engine 0 = CUDA-core FMA path, 1 = tcgen05/TMA path (bf16 operands, K % 64 == 0).
 * workspace: dic_gemm_workspace_bytes (tensor maps / split-K partials). */
int dic_gemm_nt(int engine, int M, int N, int K, const void* A, int a_dtype, const void* B,
                int b_dtype, const float* bias, float* C, void* stream);

/* General form: A(m,k) = A[m*a_m + k*a_k], B(n,k) = B[n*b_n + k*b_k] (element strides), so
 * K-major and MN-major operands of both engines can be exercised; splits > 1 = atomic split-K
 * into a zero-filled C. */
int dic_gemm_ex(int engine, int M, int N, int K, const void* A, int a_dtype, long long a_m,
                long long a_k, const void* B, int b_dtype, long long b_n, long long b_k,
                const float* bias, float* C, long long ldc, int splits, void* stream);

/* Same product with a bf16 C[M,N] (row stride ldc elements) -- the output mode of the logits / att1 GEMMs of the
 * bf16 training step (depth_models.py:166,203 in bf16 storage).  engine 1 with ldc % 8 == 0 and a 16-byte aligned C
 * takes the shared-memory + bulk-tensor-store epilogue (csrc/gemm_tc.cuh). */
int dic_gemm_nt_bf16(int engine, int M, int N, int K, const void* A, const void* B, const float* bias,
                     void* C, long long ldc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIC_H_ */
