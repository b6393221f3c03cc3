"""CPU oracle for the Show-Attend-Tell + depth-fusion decoder hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product path
(``depth_image_captioning_pub_b200``) never routes through this file and fails
loudly when its CUDA extension is missing.

What it is: a plain PyTorch (CPU, fp32 or fp64) restatement of the algorithm
the reference decoders run, written as explicit tensor math on a flat
``state_dict``-keyed weight dictionary (no ``nn.Module``s, no ``nn.LSTMCell``).
Every function cites the reference file:line it restates (paths relative to
the reference checkout).

Parity pin: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so the pin is manufactured: ``oracle/make_golden.py``
imports the UNMODIFIED reference modules in the build container, runs them on
seeded inputs and commits the input/output vectors under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every function below against those
vectors.  Beam search has no reference implementation at all (SURVEY.md
section 0, row 2): ``beam_search`` here is the build's own specification and
is "parity unpinned" with respect to the reference.

The per-step op order deliberately follows the reference AS WRITTEN (the
loop-invariant annotation projection is recomputed every timestep and the
[B,L,D] product temporary is materialised), so that timing this file is a fair
"port" CPU baseline of the reference path.  ``hoist=True`` gives the
algebraically identical hoisted form for faster large-size checks.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

Weights = Dict[str, torch.Tensor]

# state_dict keys shared by all four reference decoders
# (depth_models.py:106-135, 532-562; base_caption_models.py:59-91)
KEYS = (
    "attention.encoder_att.weight", "attention.encoder_att.bias",
    "attention.decoder_att.weight", "attention.decoder_att.bias",
    "attention.full_att.weight", "attention.full_att.bias",
    "embed.weight",
    "decode_step.weight_ih", "decode_step.weight_hh",
    "decode_step.bias_ih", "decode_step.bias_hh",
    "init_linear.weight", "init_linear.bias",
    "f_beta.weight", "f_beta.bias",
    "linear.weight", "linear.bias",
)


def make_weights(A: int, E: int, D: int, H: int, V: int, seed: int = 1234,
                 dtype=torch.float32) -> Weights:
    """Random weights with the reference's shapes and init ranges.

    nn.Linear / nn.LSTMCell default init is U(-1/sqrt(fan_in), 1/sqrt(fan_in))
    (LSTMCell: 1/sqrt(hidden)); embed and linear.weight are re-drawn from
    U(-0.1, 0.1) and linear.bias zeroed (depth_models.py:140-143).  The draw
    ORDER differs from constructing the reference module, so these are not the
    same numbers as ``torch.manual_seed(seed); CD_RNNDecoder...()`` -- golden
    fixtures store the reference module's own weights instead.
    """
    g = torch.Generator().manual_seed(seed)

    def u(shape, bound):
        return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)

    w: Weights = {}
    w["attention.encoder_att.weight"] = u((A, D), D ** -0.5)
    w["attention.encoder_att.bias"] = u((A,), D ** -0.5)
    w["attention.decoder_att.weight"] = u((A, H), H ** -0.5)
    w["attention.decoder_att.bias"] = u((A,), H ** -0.5)
    w["attention.full_att.weight"] = u((1, A), A ** -0.5)
    w["attention.full_att.bias"] = u((1,), A ** -0.5)
    w["embed.weight"] = u((V, E), 0.1)
    w["decode_step.weight_ih"] = u((4 * H, E + D), H ** -0.5)
    w["decode_step.weight_hh"] = u((4 * H, H), H ** -0.5)
    w["decode_step.bias_ih"] = u((4 * H,), H ** -0.5)
    w["decode_step.bias_hh"] = u((4 * H,), H ** -0.5)
    w["init_linear.weight"] = u((2 * H, D), D ** -0.5)
    w["init_linear.bias"] = u((2 * H,), D ** -0.5)
    w["f_beta.weight"] = u((D, H), H ** -0.5)
    w["f_beta.bias"] = u((D,), H ** -0.5)
    w["linear.weight"] = u((V, H), 0.1)
    w["linear.bias"] = torch.zeros(V, dtype=dtype)
    return w


def synthetic_vocab(V: int) -> Dict[str, int]:
    """Special tokens are the LAST four ids, in the order the reference's
    vocabulary builder appends them (dataset/vocabulary_dict.ipynb cell 1)."""
    return {"<start>": V - 4, "<end>": V - 3, "<unk>": V - 2, "<null>": V - 1}


# --------------------------------------------------------------------------
# attention (attention.py)
# --------------------------------------------------------------------------
def attention_energy(w: Weights, feats: torch.Tensor, h: torch.Tensor,
                     att1: Optional[torch.Tensor] = None) -> torch.Tensor:
    """e[b,l] = relu(att1 + att2) . w_full + b_full   (attention.py:84-87).

    The nonlinearity is ReLU, not tanh (attention.py:73)."""
    if att1 is None:
        att1 = feats @ w["attention.encoder_att.weight"].t() + w["attention.encoder_att.bias"]
    att2 = h @ w["attention.decoder_att.weight"].t() + w["attention.decoder_att.bias"]
    s = torch.relu(att1 + att2.unsqueeze(1))
    e = s @ w["attention.full_att.weight"].t() + w["attention.full_att.bias"]
    return e.squeeze(2)


def soft_attention(w: Weights, feats: torch.Tensor, h: torch.Tensor,
                   att1: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Soft_Attention.forward (attention.py:81-95) -> (context [B,D], alpha [B,L])."""
    e = attention_energy(w, feats, h, att1)
    alpha = e.softmax(dim=1)
    ctx = (feats * alpha.unsqueeze(2)).sum(dim=1)
    return ctx, alpha


def gumbel_noise(u: torch.Tensor) -> torch.Tensor:
    """g = -log(-log u)  (attention.py:18, 41)."""
    return -torch.log(-torch.log(u))


def gumbel_softmax_attention(w: Weights, feats, h, u: torch.Tensor, temp,
                             att1=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Hard_Attention.forward (attention.py:132-148) with the uniform draw
    ``u`` (attention.py:17, made there by torch.rand on the CPU generator) as
    an explicit input: alpha = softmax((e + g) / temp)."""
    e = attention_energy(w, feats, h, att1)
    alpha = ((e + gumbel_noise(u).to(e.dtype)) / temp).softmax(dim=1)
    ctx = (feats * alpha.unsqueeze(2)).sum(dim=1)
    return ctx, alpha


def gumbel_max_attention(w: Weights, feats, h, u: torch.Tensor,
                         att1=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Hard_Attention.Hard_sample (attention.py:150-167): one-hot int64 alpha
    at argmax(e + g); the context is still written as the dense weighted sum."""
    e = attention_energy(w, feats, h, att1)
    pos = torch.argmax(e + gumbel_noise(u).to(e.dtype), dim=1)
    alpha = torch.nn.functional.one_hot(pos, num_classes=e.shape[1])
    ctx = (feats * alpha.unsqueeze(2)).sum(dim=1)
    return ctx, alpha


# --------------------------------------------------------------------------
# one decoder timestep
# --------------------------------------------------------------------------
def init_state(w: Weights, feats: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """h0, c0 = chunk(init_linear(mean_l F), 2)   (depth_models.py:166-168)."""
    s = feats.mean(dim=1) @ w["init_linear.weight"].t() + w["init_linear.bias"]
    h, c = s.chunk(2, dim=1)
    return h, c


def lstm_cell(w: Weights, x, h, c):
    """torch.nn.LSTMCell semantics, gate order i,f,g,o (depth_models.py:122,193)."""
    g = (x @ w["decode_step.weight_ih"].t() + w["decode_step.bias_ih"]
         + h @ w["decode_step.weight_hh"].t() + w["decode_step.bias_hh"])
    i, f, gg, o = g.chunk(4, dim=1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return h2, c2


def decoder_step(w: Weights, feats, h, c, emb, attn: str = "soft", u=None, temp=None, att1=None):
    """One iteration of the loop body depth_models.py:185-194 (soft),
    :613-622 (gumbel-softmax), :662-671 (gumbel-max)."""
    if attn == "soft":
        ctx, alpha = soft_attention(w, feats, h, att1)
    elif attn == "gumbel_softmax":
        ctx, alpha = gumbel_softmax_attention(w, feats, h, u, temp, att1)
    elif attn == "gumbel_max":
        ctx, alpha = gumbel_max_attention(w, feats, h, u, att1)
    else:
        raise ValueError(attn)
    beta = torch.sigmoid(h @ w["f_beta.weight"].t() + w["f_beta.bias"])
    x = torch.cat((emb, beta * ctx), dim=1)
    h2, c2 = lstm_cell(w, x, h, c)
    return h2, c2, alpha


def _fuse(F_rgb, F_depth):
    """features.add(depth_features) (depth_models.py:163); base decoders have
    no second tensor (base_caption_models.py:105)."""
    return F_rgb if F_depth is None else F_rgb.add(F_depth)


# --------------------------------------------------------------------------
# teacher-forced forward
# --------------------------------------------------------------------------
def packed_batch_sizes(dec_lengths: Sequence[int]) -> List[int]:
    """bs_valid per step for lengths sorted descending (depth_models.py:182)."""
    return [sum(1 for l in dec_lengths if l > t) for t in range(max(dec_lengths))]


def decoder_forward(w: Weights, F_rgb, F_depth, captions, lengths: Sequence[int],
                    attn: str = "soft", u_steps: Optional[List[torch.Tensor]] = None,
                    temp=None, dropout_masks: Optional[List[torch.Tensor]] = None,
                    hoist: bool = False):
    """Teacher-forced forward (depth_models.py:153-207 soft, :580-634 hard,
    :637-689 eval_forward).

    Returns (packed_logits [sum(len-1), V] time-major, batch_sizes list,
    alphas [B, Tmax, L] zero padded).  ``u_steps[t]`` is the [bs_valid_t, L]
    uniform draw of step t for the hard variants.  ``dropout_masks[t]``
    ([bs_valid_t, H], already scaled by 1/(1-p)) stands in for nn.Dropout
    (depth_models.py:197); None means eval mode.
    """
    feats = _fuse(F_rgb, F_depth)
    B, L, _ = feats.shape
    emb_all = w["embed.weight"][captions]            # depth_models.py:160
    h, c = init_state(w, feats)
    dec_lengths = [l - 1 for l in lengths]
    bsz = packed_batch_sizes(dec_lengths)
    Tmax = len(bsz)
    alphas = feats.new_zeros(B, Tmax, L)
    att1 = None
    if hoist:
        att1 = feats @ w["attention.encoder_att.weight"].t() + w["attention.encoder_att.bias"]
    outs = []
    for t in range(Tmax):
        n = bsz[t]
        h, c, alpha = decoder_step(
            w, feats[:n], h[:n], c[:n], emb_all[:n, t], attn,
            None if u_steps is None else u_steps[t], temp,
            None if att1 is None else att1[:n])
        hd = h if dropout_masks is None else h * dropout_masks[t]
        outs.append(hd @ w["linear.weight"].t() + w["linear.bias"])
        alphas[:n, t] = alpha.to(alphas.dtype)
    return torch.cat(outs, dim=0), bsz, alphas


def pack_targets(captions, lengths: Sequence[int]) -> torch.Tensor:
    """pack_padded_sequence(captions[:,1:], lengths-1).data (depth_train.py:210-213)."""
    dec_lengths = [l - 1 for l in lengths]
    bsz = packed_batch_sizes(dec_lengths)
    tg = captions[:, 1:]
    return torch.cat([tg[:n, t] for t, n in enumerate(bsz)], dim=0)


def caption_loss(packed_logits, packed_targets, null_id: int, alphas=None, lam: float = 0.7):
    """CE ignoring <null> + doubly-stochastic regulariser (depth_train.py:132-133,
    214-216); the hard variants use the CE term only (:530-532)."""
    loss = torch.nn.functional.cross_entropy(packed_logits, packed_targets, ignore_index=null_id)
    if alphas is not None:
        loss = loss + lam * ((1.0 - alphas.sum(dim=1)) ** 2).mean()
    return loss


# --------------------------------------------------------------------------
# decoding
# --------------------------------------------------------------------------
@torch.no_grad()
def greedy_decode(w: Weights, F_rgb, F_depth, start_id: int, max_length: int,
                  attn: str = "soft", u_steps=None, hoist: bool = False,
                  use_softmax: bool = True):
    """batch_sample / sample (depth_models.py:216-305, 698-789): fixed
    max_length steps, no early stop.  Returns (tokens [B,max_length] int64,
    alphas list of [B,L], logits list of [B,V])."""
    feats = _fuse(F_rgb, F_depth)
    B = feats.shape[0]
    h, c = init_state(w, feats)
    prev = torch.full((B,), start_id, dtype=torch.int64)
    att1 = None
    if hoist:
        att1 = feats @ w["attention.encoder_att.weight"].t() + w["attention.encoder_att.bias"]
    toks, alphas, logits_all = [], [], []
    for t in range(max_length):
        emb = w["embed.weight"][prev]
        h, c, alpha = decoder_step(w, feats, h, c, emb, attn,
                                   None if u_steps is None else u_steps[t], None, att1)
        logits = h @ w["linear.weight"].t() + w["linear.bias"]
        # the reference takes argmax(softmax(logits)) (depth_models.py:296-297)
        prev = (logits.softmax(dim=1) if use_softmax else logits).argmax(dim=1)
        toks.append(prev)
        alphas.append(alpha)
        logits_all.append(logits)
    return torch.stack(toks, dim=1), alphas, logits_all


def stable_topk_desc(x: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-k along the last dim, descending, ties broken by LOWEST index."""
    idx = torch.sort(x, dim=-1, descending=True, stable=True).indices[..., :k]
    return torch.gather(x, -1, idx), idx


def beam_select(scores: torch.Tensor, finished: torch.Tensor, logits: torch.Tensor,
                lse: torch.Tensor, end_id: int):
    """One beam-search selection, the arithmetic the CUDA top-k kernel must
    reproduce BIT-EXACTLY given identical (scores, logits, lse):

        cand[b, j*V + v] = scores[b,j] + (logits[b,j,v] - lse[b,j])      (fp32, in this order)
        finished rows:     cand = scores[b,j] for v == <end>, -inf otherwise
        (new_scores, flat) = stable top-k (descending, ties -> lowest flat index)
        back = flat // V ; tok = flat % V
    """
    B, K, V = logits.shape
    logp = logits - lse.unsqueeze(2)
    cand = scores.unsqueeze(2) + logp
    frozen = torch.full_like(cand, float("-inf"))
    frozen[:, :, end_id] = scores
    cand = torch.where(finished.unsqueeze(2), frozen, cand)
    new_scores, flat = stable_topk_desc(cand.reshape(B, K * V), K)
    back = flat // V
    tok = flat % V
    new_finished = torch.gather(finished, 1, back) | (tok == end_id)
    return new_scores, back, tok, new_finished


@torch.no_grad()
def beam_search(w: Weights, F_rgb, F_depth, start_id: int, end_id: int, beam: int,
                max_length: int, hoist: bool = True, lse_fn=None):
    """Beam search over the soft-attention decoder.  NOT IN THE REFERENCE
    (it only has greedy decoding); this is the build's specification
    (SURVEY.md section 8a row 9), "parity unpinned".

    Fixed-shape formulation: every image keeps ``beam`` rows that share the
    image's annotations.  Step 0 expands row 0 only (the other rows start at
    score -inf).  A row that emits <end> is frozen: its score is kept and its
    only continuation is <end> at cost 0.  Runs exactly ``max_length`` steps.
    Returns dict(tokens [B,max_length] of the best row, padded with <end>
    after the first <end>; lengths [B] (tokens up to and including the first
    <end>, else max_length); scores [B]; back [T,B,K] int32; toks [T,B,K];
    all_scores [T,B,K]).
    """
    feats = _fuse(F_rgb, F_depth)
    B, L, D = feats.shape
    K = beam
    h, c = init_state(w, feats)
    h = h.unsqueeze(1).expand(B, K, -1).reshape(B * K, -1)
    c = c.unsqueeze(1).expand(B, K, -1).reshape(B * K, -1)
    featsK = feats.unsqueeze(1).expand(B, K, L, D).reshape(B * K, L, D)
    att1 = None
    if hoist:
        att1 = feats @ w["attention.encoder_att.weight"].t() + w["attention.encoder_att.bias"]
        att1 = att1.unsqueeze(1).expand(B, K, L, -1).reshape(B * K, L, -1)
    scores = feats.new_full((B, K), float("-inf"))
    scores[:, 0] = 0.0
    finished = torch.zeros(B, K, dtype=torch.bool)
    prev = torch.full((B * K,), start_id, dtype=torch.int64)
    backs, toks, all_scores, all_lse, all_logits = [], [], [], [], []
    for t in range(max_length):
        emb = w["embed.weight"][prev]
        h, c, _ = decoder_step(w, featsK, h, c, emb, "soft", None, None, att1)
        logits = (h @ w["linear.weight"].t() + w["linear.bias"]).reshape(B, K, -1)
        lse = torch.logsumexp(logits, dim=2) if lse_fn is None else lse_fn(t, logits)
        scores, back, tok, finished = beam_select(scores, finished, logits, lse, end_id)
        gidx = (back + torch.arange(B).unsqueeze(1) * K).reshape(-1)
        h, c = h[gidx], c[gidx]
        prev = tok.reshape(-1)
        backs.append(back.to(torch.int32)); toks.append(tok); all_scores.append(scores.clone())
        all_lse.append(lse); all_logits.append(logits)
    # backtrack from the best final row (row 0: top-k output is sorted descending)
    T = max_length
    out = torch.full((B, T), end_id, dtype=torch.int64)
    row = torch.zeros(B, dtype=torch.int64)
    for t in range(T - 1, -1, -1):
        out[:, t] = toks[t][torch.arange(B), row]
        row = backs[t][torch.arange(B), row].to(torch.int64)
    is_end = out == end_id
    first_end = torch.where(is_end.any(dim=1), is_end.to(torch.int64).argmax(dim=1) + 1,
                            torch.full((B,), T, dtype=torch.int64))
    return dict(tokens=out, lengths=first_end, scores=scores[:, 0].clone(),
                back=torch.stack(backs), toks=torch.stack(toks),
                all_scores=torch.stack(all_scores), lse=torch.stack(all_lse),
                logits=all_logits)


def beam_search_lookahead(w: Weights, F_rgb, F_depth, start_id: int, end_id: int, beam: int,
                          max_length: int):
    """The SAME search as ``beam_search`` in the order the CUDA path runs it (csrc/dic_api.cu decode_impl,
    "look-ahead attention"; test infrastructure like the rest of this file).

    The attention of step t+1 needs h_t of a row's PARENT only, and the selection of step t merely permutes /
    duplicates the rows of an image.  So nothing is reordered: every per-row tensor stays in parent order, the
    gated context of step t+1 is computed from the un-reordered h_t (before the selection of step t is known), the
    gate pre-activations are P = [beta*z | h] . [W_z | W_hh]^T on the parents, and a child row r adds its token's
    share from a table:  gates[r] = P[parent(r)] + (Emb . W_e^T)[token(r)] + b_ih + b_hh,  c_prev = c[parent(r)].
    tests/test_oracle_golden.py pins this to beam_search (same tokens, backpointers and scores).
    """
    feats = _fuse(F_rgb, F_depth)
    B, L, D = feats.shape
    K = beam
    E = w["embed.weight"].shape[1]
    W_ih, W_hh = w["decode_step.weight_ih"], w["decode_step.weight_hh"]
    bias_g = w["decode_step.bias_ih"] + w["decode_step.bias_hh"]
    etab = w["embed.weight"] @ W_ih[:, :E].t()                      # [V, 4H]: a token's share of the gates
    W_zh = torch.cat((W_ih[:, E:], W_hh), dim=1)                    # [4H, D+H]
    featsK = feats.unsqueeze(1).expand(B, K, L, D).reshape(B * K, L, D)
    att1 = feats @ w["attention.encoder_att.weight"].t() + w["attention.encoder_att.bias"]
    att1 = att1.unsqueeze(1).expand(B, K, L, -1).reshape(B * K, L, -1)

    def gated_context(h):
        ctx, _ = soft_attention(w, featsK, h, att1)
        return torch.sigmoid(h @ w["f_beta.weight"].t() + w["f_beta.bias"]) * ctx

    def cell(gates, c_prev):
        i, f, gg, o = gates.chunk(4, dim=1)
        c2 = torch.sigmoid(f) * c_prev + torch.sigmoid(i) * torch.tanh(gg)
        return torch.sigmoid(o) * torch.tanh(c2), c2

    h0, c0 = init_state(w, feats)
    h = h0.unsqueeze(1).expand(B, K, -1).reshape(B * K, -1)         # parent order from here on
    c = c0.unsqueeze(1).expand(B, K, -1).reshape(B * K, -1)
    parent = torch.arange(B * K)                                     # step 0: every row is its own parent
    tok_prev = torch.full((B * K,), start_id, dtype=torch.int64)
    zg = gated_context(h)                                            # context of step 0 (in order)
    scores = feats.new_full((B, K), float("-inf"))
    scores[:, 0] = 0.0
    finished = torch.zeros(B, K, dtype=torch.bool)
    backs, toks = [], []
    for t in range(max_length):
        P = torch.cat((zg, h), dim=1) @ W_zh.t()                     # gate GEMM on the parents
        h, c = cell(P[parent] + etab[tok_prev] + bias_g, c[parent])  # the LSTM kernel follows the backpointers
        logits = (h @ w["linear.weight"].t() + w["linear.bias"]).reshape(B, K, -1)
        if t + 1 < max_length:
            zg = gated_context(h)                                    # look-ahead: before the selection below
        lse = torch.logsumexp(logits, dim=2)
        scores, back, tok, finished = beam_select(scores, finished, logits, lse, end_id)
        parent = (back + torch.arange(B).unsqueeze(1) * K).reshape(-1)
        tok_prev = tok.reshape(-1)
        backs.append(back.to(torch.int32)); toks.append(tok)
    T = max_length
    out = torch.full((B, T), end_id, dtype=torch.int64)
    row = torch.zeros(B, dtype=torch.int64)
    for t in range(T - 1, -1, -1):
        out[:, t] = toks[t][torch.arange(B), row]
        row = backs[t][torch.arange(B), row].to(torch.int64)
    return dict(tokens=out, scores=scores[:, 0].clone(), back=torch.stack(backs), toks=torch.stack(toks))
