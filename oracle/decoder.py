"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement, in plain torch functional ops, of the autoregressive front half of the reference's `infer()`
(/root/reference/HelloSippyTTSRT/HelloSippyRTPipe.py:195-229): per decoder step

    prenet(output_sequence, speaker)[:, -1:]      transformers modeling_speecht5.py:648-697  (SpeechT5SpeechDecoderPrenet; scaled positions :400-422)
    wrapped_decoder(.., past_key_values)          :1449-1607 -> six SpeechT5DecoderLayer :1070-1160 (post-LN; attention :837-986, FFN :989-1010)
    feat_out / prob_out                           :740-750 (SpeechT5SpeechDecoderPostnet), call sites HelloSippyRTPipe.py:213,223
    stop bookkeeping                              HelloSippyRTPipe.py:225-228

[third-party, transformers 5.5.0 in this image; the reference pins only ">=4.0.0", /root/reference/requirements.txt:3]

The prenet's dropout is ALWAYS on (:689-691, `_consistent_dropout`: one Bernoulli(1 - p) mask per (position, unit), shared by the batch);
here the masks are explicit inputs so that the restatement, the live module and the CUDA path can be driven with the same ones.
Pinned by tests/test_oracle_decoder.py, which runs the LIVE transformers modules through the reference's own call sequence on the same
weights and masks, and by tests/golden/decoder_golden.npz (oracle/make_golden_decoder.py).  Only tests/, smoke() and bench.py's CPU legs
may import this module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

HIDDEN, LAYERS, HEADS, HEAD_DIM = 768, 6, 12, 64
P = "speecht5.decoder.prenet."
D = "speecht5.decoder.wrapped_decoder.layers."


def position_table(max_len: int, dim: int = HIDDEN) -> torch.Tensor:
    """SpeechT5ScaledPositionalEncoding.pe (modeling_speecht5.py:405-412), same expressions, fp32."""
    pe = torch.zeros(max_len, dim)
    position = torch.arange(0, max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dim, 2, dtype=torch.int64).float() * -(math.log(10000.0) / dim))
    pe[:, 0::2] = torch.sin(position.float() * div_term)
    pe[:, 1::2] = torch.cos(position.float() * div_term)
    return pe


def prenet_last(sd: Dict[str, torch.Tensor], frame: torch.Tensor, pos: torch.Tensor, speaker: torch.Tensor,
                masks: Optional[torch.Tensor], p: float = 0.5, pe: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The last position of decoder.prenet(output_sequence, speaker) (:676-697): frame (B, 80) is output_sequence[:, -1], pos (B,) its
    index in the sequence, masks (2, 256) the dropout keep-masks (0/1) of that position or None for no dropout."""
    x = frame
    for i in range(2):
        x = F.relu(F.linear(x, sd[P + f"layers.{i}.weight"], sd[P + f"layers.{i}.bias"]))      # :689
        if masks is not None:
            x = torch.where(masks[i] == 1, x, torch.zeros((), dtype=x.dtype)) * 1 / (1 - p)       # :671-674
    x = F.linear(x, sd[P + "final_layer.weight"], sd[P + "final_layer.bias"])                  # :692
    pe = position_table(int(pos.max()) + 1) if pe is None else pe
    x = x + sd[P + "encode_positions.alpha"] * pe[pos]                                         # :420
    s = F.normalize(speaker)                                                                    # :696
    x = torch.cat([x, s], dim=-1)
    return F.relu(F.linear(x, sd[P + "speaker_embeds_layer.weight"], sd[P + "speaker_embeds_layer.bias"]))   # :699


def _heads(x):                                                                                  # (B, T, 768) -> (B, 12, T, 64)
    B, T, _ = x.shape
    return x.view(B, T, HEADS, HEAD_DIM).transpose(1, 2)


def _attend(sd, pre, h, K, V, mask):
    """SpeechT5Attention.forward for one query position (:872-986): h (B, 768); K, V (B, T, 768) already projected; mask (B, T) bool or None."""
    q = F.linear(h, sd[pre + "q_proj.weight"], sd[pre + "q_proj.bias"]) * (HEAD_DIM ** -0.5)      # :889
    q = _heads(q[:, None])                                                                      # (B, 12, 1, 64)
    w = torch.matmul(q, _heads(K).transpose(2, 3))                                              # :925
    if mask is not None:
        w = w.masked_fill(~mask[:, None, None, :], torch.finfo(w.dtype).min)                    # :943-949 (additive min == excluded)
    w = F.softmax(w, dim=-1)                                                                    # :951
    o = torch.matmul(w, _heads(V)).transpose(1, 2).reshape(h.size(0), HIDDEN)                   # :965-978
    return F.linear(o, sd[pre + "out_proj.weight"], sd[pre + "out_proj.bias"])                  # :980


def cross_kv(sd, enc: torch.Tensor) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Cross-attention keys / values of the six layers, computed once per sentence (:915-923 on the first step, reused after)."""
    out = []
    for i in range(LAYERS):
        pre = D + f"{i}.encoder_attn."
        out.append((F.linear(enc, sd[pre + "k_proj.weight"], sd[pre + "k_proj.bias"]), F.linear(enc, sd[pre + "v_proj.weight"], sd[pre + "v_proj.bias"])))
    return out


def decoder_step(sd, h: torch.Tensor, self_kv: List[Optional[Tuple[torch.Tensor, torch.Tensor]]], xkv, enc_mask: torch.Tensor) -> torch.Tensor:
    """wrapped_decoder on one new position with a KV cache (:1095-1160 per layer).  h (B, 768); self_kv is updated in place."""
    for i in range(LAYERS):
        L = D + f"{i}."
        pre = L + "self_attn."
        k = F.linear(h, sd[pre + "k_proj.weight"], sd[pre + "k_proj.bias"])[:, None]
        v = F.linear(h, sd[pre + "v_proj.weight"], sd[pre + "v_proj.bias"])[:, None]
        K, V = (k, v) if self_kv[i] is None else (torch.cat([self_kv[i][0], k], 1), torch.cat([self_kv[i][1], v], 1))
        self_kv[i] = (K, V)
        h = F.layer_norm(h + _attend(sd, pre, h, K, V, None), (HIDDEN,), sd[L + "self_attn_layer_norm.weight"], sd[L + "self_attn_layer_norm.bias"], 1e-5)
        h = F.layer_norm(h + _attend(sd, L + "encoder_attn.", h, xkv[i][0], xkv[i][1], enc_mask.bool()), (HIDDEN,),
                         sd[L + "encoder_attn_layer_norm.weight"], sd[L + "encoder_attn_layer_norm.bias"], 1e-5)
        f = F.gelu(F.linear(h, sd[L + "feed_forward.intermediate_dense.weight"], sd[L + "feed_forward.intermediate_dense.bias"]))       # :1004-1005 (erf gelu)
        f = F.linear(f, sd[L + "feed_forward.output_dense.weight"], sd[L + "feed_forward.output_dense.bias"])
        h = F.layer_norm(h + f, (HIDDEN,), sd[L + "final_layer_norm.weight"], sd[L + "final_layer_norm.bias"], 1e-5)                      # :1150-1151
    return h


class DecoderState:
    """What the reference keeps in HelloSippyPipeStateBatched for the front half (:81-118): output_sequence's last frame and length,
    past_key_values, encoder states + mask, speaker."""

    def __init__(self, sd, enc: torch.Tensor, enc_mask: torch.Tensor, speaker: torch.Tensor):
        B = enc.size(0)
        self.enc_mask, self.speaker = enc_mask, speaker
        self.xkv = cross_kv(sd, enc)
        self.self_kv: List[Optional[Tuple[torch.Tensor, torch.Tensor]]] = [None] * LAYERS
        self.last = torch.zeros(B, 80, dtype=enc.dtype)                # output_sequence starts as one all-zero frame (:118)
        self.pos = torch.zeros(B, dtype=torch.long)


def step(sd, st: DecoderState, masks: Optional[torch.Tensor], pe: Optional[torch.Tensor] = None):
    """One trip of the reference's while loop (:195-223): -> spectrum (B, 2, 80) and stop probabilities (B, 2)."""
    h = prenet_last(sd, st.last, st.pos, st.speaker, masks, pe=pe)
    h = decoder_step(sd, h, st.self_kv, st.xkv, st.enc_mask)
    spectrum = F.linear(h, sd["speech_decoder_postnet.feat_out.weight"], sd["speech_decoder_postnet.feat_out.bias"]).view(-1, 2, 80)   # :213-215
    prob = torch.sigmoid(F.linear(h, sd["speech_decoder_postnet.prob_out.weight"], sd["speech_decoder_postnet.prob_out.bias"]))         # :223
    st.last = spectrum[:, -1]                                                                                                           # :219-220
    st.pos = st.pos + 1
    return spectrum, prob
