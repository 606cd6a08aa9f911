"""ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement, in plain torch functional ops, of the floating-point half of the
reference's TTS tail.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference leg may import this module.

What it follows (reference file:line):
  * window builder + re-assembly   /root/reference/HelloSippyTTSRT/HelloSippyRTPipe.py:231-240
  * HiFiGAN vocoder                 transformers/models/speecht5/modeling_speecht5.py:2954-2962 (ResBlock),
                                    :3055-3085 (SpeechT5HifiGan.forward)   [third-party, transformers 5.5.0,
                                    reference pins only ">=4.0.0": /root/reference/requirements.txt:3]
  * chunker (AmendmentNetwork1)     /root/reference/HelloSippyTTSRT/HelloSippyRT.py:182-198, 219-237
  * resampler                       torchaudio/functional/functional.py:1340-1402 (kernel), :1416-1428 (apply)
                                    [third-party, torchaudio 2.11.0; requirements.txt:7]
  * unbatch slicing                 /root/reference/HelloSippyTTSRT/HelloSippyRTPipe.py:242-259
  * decoder post-net                transformers/models/speecht5/modeling_speecht5.py:700-737 (BatchNormConvLayer), :758-762
                                    (SpeechT5SpeechDecoderPostnet.postnet), called at HelloSippyRTPipe.py:230

Pinned by tests/golden/*.npz, which were produced by the REAL reference modules in the
build container (oracle/make_golden.py); tests/test_oracle_golden.py checks this file against
them on every CPU run.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

LRELU_VOC = 0.1      # SpeechT5HifiGanConfig.leaky_relu_slope (configuration_speecht5.py:267-276)
LRELU_DEFAULT = 0.01  # torch default, used before conv_post (modeling_speecht5.py:3074) and by the chunker
RES_KERNELS = (3, 7, 11)
RES_DILATIONS = (1, 3, 5)


# --------------------------------------------------------------------------- HiFiGAN
def hifigan_forward(sd: Dict[str, torch.Tensor], mel: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """mel (W, T, 80) -> waveform (W, 256*T).  modeling_speecht5.py:3055-3085."""
    x = (mel - sd["mean"]) / sd["scale"]                                   # :3055-3056
    h = x.transpose(2, 1)                                                   # :3062
    h = F.conv1d(h, sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)  # :3064
    if taps is not None: taps["conv_pre"] = h
    for i in range(4):
        h = F.leaky_relu(h, LRELU_VOC)                                      # :3066
        h = F.conv_transpose1d(h, sd[f"upsampler.{i}.weight"], sd[f"upsampler.{i}.bias"],
                               stride=4, padding=2)                         # :3067
        if taps is not None: taps[f"up{i}"] = h
        acc = None
        for j, k in enumerate(RES_KERNELS):
            n = i * 3 + j
            r = h
            for di, d in enumerate(RES_DILATIONS):                          # :2954-2962
                res = r
                r = F.leaky_relu(r, LRELU_VOC)
                r = F.conv1d(r, sd[f"resblocks.{n}.convs1.{di}.weight"], sd[f"resblocks.{n}.convs1.{di}.bias"],
                             dilation=d, padding=(k * d - d) // 2)
                r = F.leaky_relu(r, LRELU_VOC)
                r = F.conv1d(r, sd[f"resblocks.{n}.convs2.{di}.weight"], sd[f"resblocks.{n}.convs2.{di}.bias"],
                             padding=(k - 1) // 2)
                r = r + res
            acc = r if acc is None else acc + r                             # :3069-3071
        h = acc / 3                                                         # :3072
        if taps is not None: taps[f"stage{i}"] = h
    h = F.leaky_relu(h)                                                     # :3074 (default slope 0.01)
    h = F.conv1d(h, sd["conv_post.weight"], sd["conv_post.bias"], padding=3)
    return torch.tanh(h).squeeze(1)                                         # :3076-3083


# --------------------------------------------------------------------------- decoder post-net
def postnet_forward(sd: Dict[str, torch.Tensor], mel: torch.Tensor) -> torch.Tensor:
    """mel (B, T, 80) -> (B, T, 80).  modeling_speecht5.py:758-762 over five layers of :731-737 in eval mode
    (dropout is the identity; BatchNorm1d uses its running statistics, eps 1e-5)."""
    h = mel.transpose(1, 2)                                                 # :759
    for i in range(5):
        h = F.conv1d(h, sd[f"layers.{i}.conv.weight"], None, padding=2)     # :714-721, :732  (kernel 5, no bias)
        h = F.batch_norm(h, sd[f"layers.{i}.batch_norm.running_mean"], sd[f"layers.{i}.batch_norm.running_var"],
                         sd[f"layers.{i}.batch_norm.weight"], sd[f"layers.{i}.batch_norm.bias"], training=False, eps=1e-5)   # :733
        if i < 4:
            h = torch.tanh(h)                                               # :724-727, :734-735
    return mel + h.transpose(1, 2)                                          # :762


# --------------------------------------------------------------------------- chunker
def chunker_forward(sd: Dict[str, torch.Tensor], mel: torch.Tensor, audio: torch.Tensor) -> torch.Tensor:
    """mel (W,12,80) + audio (W,3072) -> (W,2048).  HelloSippyRT.py:219-237."""
    W = audio.size(0)
    T = mel.size(-1)                                    # :221  (80: a raw reinterpretation, not a transpose)
    a = audio.contiguous().view(W, 256, -1)             # :223
    m = mel.contiguous().view(W, T, -1)                 # :224
    xm = F.conv1d(m, sd["conv_pre_m.weight"], sd["conv_pre_m.bias"], padding=1)
    xa = F.conv1d(a, sd["conv_pre_a.weight"], sd["conv_pre_a.bias"], padding=1)
    z = torch.cat((xm, xa), dim=1)                      # :228
    for i in range(2):                                  # :229-231
        z = F.leaky_relu(z, LRELU_DEFAULT)
        z = F.conv_transpose1d(z, sd[f"upsampler.{i}.weight"], sd[f"upsampler.{i}.bias"], stride=4, padding=2)
    res = z                                             # SimpleResidualBlock :190-198
    z = F.leaky_relu(z, LRELU_DEFAULT)
    z = F.conv1d(z, sd["resblock.conv1.weight"], sd["resblock.conv1.bias"], padding=1)
    z = F.leaky_relu(z, LRELU_DEFAULT)
    z = F.conv1d(z, sd["resblock.conv2.weight"], sd["resblock.conv2.bias"], padding=3, dilation=3)
    z = z + res
    z = F.leaky_relu(z, LRELU_DEFAULT)                  # :233
    z = F.conv1d(z, sd["post_conv.weight"], sd["post_conv.bias"], stride=24)
    g = F.leaky_relu(z, LRELU_DEFAULT).reshape(W, -1)   # :235
    return torch.tanh(audio[:, 512:-512] * g)           # :236-237


# --------------------------------------------------------------------------- resampler
def resample_kernel(orig: int, new: int, dtype=torch.float32) -> Tuple[torch.Tensor, int]:
    """torchaudio _get_sinc_resample_kernel, sinc_interp_hann, width 6, rolloff 0.99
    (functional.py:1340-1402).  Returns (kernel (new, 1, K), width)."""
    g = math.gcd(orig, new)
    orig, new = orig // g, new // g
    lowpass_filter_width, rolloff = 6, 0.99
    base_freq = min(orig, new) * rolloff
    width = math.ceil(lowpass_filter_width * orig / base_freq)
    idx = torch.arange(-width, width + orig, dtype=torch.float64)[None, None] / orig
    t = torch.arange(0, -new, -1, dtype=torch.float64)[:, None, None] / new + idx
    t = t * base_freq
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    scale = base_freq / orig
    kernels = torch.where(t == 0, torch.tensor(1.0, dtype=torch.float64), t.sin() / t)
    kernels = kernels * window * scale
    return kernels.to(dtype), width


def resample(x: torch.Tensor, orig: int, new: int) -> torch.Tensor:
    """(..., L) -> (..., ceil(L*new/orig)).  functional.py:1405-1432; zero padding per call."""
    g = math.gcd(orig, new)
    o, n = orig // g, new // g
    kernel, width = resample_kernel(orig, new, x.dtype)
    shape = x.shape
    w = x.reshape(-1, shape[-1])
    L = w.shape[-1]
    w = F.pad(w, (width, width + o))
    y = F.conv1d(w[:, None], kernel, stride=o)
    y = y.transpose(1, 2).reshape(w.shape[0], -1)
    target = int(math.ceil(n * L / o))
    return y[..., :target].reshape(shape[:-1] + (target,))


# --------------------------------------------------------------------------- windows / tail
def build_windows(pre_frames: torch.Tensor, mel: torch.Tensor, chunk: int = 8, eframes: int = 4):
    """HelloSippyRTPipe.py:231-235.  pre_frames (B,4,80), mel (B,n,80), n % chunk == 0.
    Returns windows (nchunks*B, chunk+eframes, 80) stacked chunk-major on dim 0, and new pre_frames."""
    spec = torch.cat((pre_frames, mel), dim=1)
    new_pre = spec[:, -eframes:, :]
    nchunks = spec.size(1) // chunk
    win = torch.cat([spec[:, i * chunk:(i + 1) * chunk + eframes, :] for i in range(nchunks)], dim=0)
    return win, new_pre


def tts_tail(voc_sd, chk_sd, pre_frames, mel, output_sr: int = 8000, model_sr: int = 16000,
             use_chunker: bool = True):
    """One reference `infer()` tail: lines 231-240.  Returns (audio (B, n*256*output_sr/model_sr), new pre_frames)."""
    B = mel.size(0)
    win, new_pre = build_windows(pre_frames, mel)
    audio = hifigan_forward(voc_sd, win)                # :236
    if use_chunker:
        audio = chunker_forward(chk_sd, win, audio)     # :237
    else:
        audio = audio[:, 512:-512]
    slices = audio.split(B, dim=0)                      # :238
    audio = torch.cat(slices, dim=1)                    # :239
    if output_sr != model_sr:
        audio = resample(audio, model_sr, output_sr)    # :240
    return audio, new_pre


def unbatch_slices(asize: int, idx: int, starts_at: List[int], ends_at: List[int], live: List[bool],
                   sr_rr: int = 2) -> Tuple[List[Optional[Tuple[int, int]]], List[bool], bool]:
    """Index arithmetic of unbatch_and_dispatch (HelloSippyRTPipe.py:242-259).
    Returns per-session (startoff, endoff) or None when nothing is emitted, per-session `finished now`
    flags, and the method's return value (True = keep going)."""
    end_idx = idx - 1
    stepsize = 256 * 2 // sr_rr
    out, fin = [], []
    for i in range(len(starts_at)):
        if not live[i]:
            out.append(None); fin.append(False); continue
        startoff = max(0, asize - ((idx - starts_at[i]) * stepsize))
        e = ends_at[i]
        endoff = min(asize, asize - (((idx - e) * stepsize) if e >= 0 else 0))
        assert startoff <= endoff
        out.append((startoff, endoff) if startoff != endoff else None)
        fin.append(e >= 0 and e <= end_idx)
    more = any((e < 0) or (e > end_idx) for e in ends_at)
    return out, fin, more


# --------------------------------------------------------------------------- bf16-operand emulation
def _bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def hifigan_forward_bf16emu(sd: Dict[str, torch.Tensor], mel: torch.Tensor, split_weights: bool = False) -> torch.Tensor:
    """What the product's tensor-core mode computes, restated on CPU: every conv of the four
    upsample stages takes bf16-rounded operands (activations after leaky-ReLU, weights) and
    accumulates in fp32; the residual stream, conv_pre and conv_post stay fp32.  Used to predict
    the SNR of that mode against the fp32 reference and to test the tcgen05 kernels tightly."""
    def wq(w):
        if split_weights:
            hi = _bf(w)
            return hi + _bf(w - hi)
        return _bf(w)
    x = (mel - sd["mean"]) / sd["scale"]
    h = F.conv1d(x.transpose(2, 1), sd["conv_pre.weight"], sd["conv_pre.bias"], padding=3)
    for i in range(4):
        a = _bf(F.leaky_relu(h, LRELU_VOC))
        h = F.conv_transpose1d(a, wq(sd[f"upsampler.{i}.weight"]), sd[f"upsampler.{i}.bias"], stride=4, padding=2)
        acc = None
        for j, k in enumerate(RES_KERNELS):
            n = i * 3 + j
            r = h
            for di, d in enumerate(RES_DILATIONS):
                a1 = _bf(F.leaky_relu(r, LRELU_VOC))
                y = F.conv1d(a1, wq(sd[f"resblocks.{n}.convs1.{di}.weight"]), sd[f"resblocks.{n}.convs1.{di}.bias"],
                             dilation=d, padding=(k * d - d) // 2)
                a2 = _bf(F.leaky_relu(y, LRELU_VOC))
                r = r + F.conv1d(a2, wq(sd[f"resblocks.{n}.convs2.{di}.weight"]), sd[f"resblocks.{n}.convs2.{di}.bias"],
                                 padding=(k - 1) // 2)
            acc = r if acc is None else acc + r
        h = acc / 3
    h = F.leaky_relu(h)
    h = F.conv1d(h, sd["conv_post.weight"], sd["conv_post.bias"], padding=3)
    return torch.tanh(h).squeeze(1)


def snr_db(ref: torch.Tensor, x: torch.Tensor) -> float:
    ref = ref.double(); x = x.double()
    return float(10.0 * torch.log10(ref.pow(2).sum() / (ref - x).pow(2).sum().clamp_min(1e-300)))
