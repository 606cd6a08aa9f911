"""ORACLE (test infrastructure): freezes golden vectors under tests/golden/ by running the REAL
reference modules where they lie (/root/reference + the third-party modules it calls).

Run in the build container only:   python -m oracle.make_golden
The fixtures travel with the repo; /root/reference does not exist on the GPU box.

What is frozen, and which reference code produced it:
  g711_tables.npz      audioop.lin2ulaw/ulaw2lin/lin2alaw/alaw2lin over all inputs  (G711.py:7-19 builds its
                       LUTs from exactly these calls)
  g711_golden.json     sha256 of the tables and of byte streams from the real Core.Codecs.G711.G711Codec
                       (vectors G1, G2, G4 of SURVEY.md App. A.4) + the real codec's small-case outputs
  resample_taps.npz    torchaudio.transforms.Resample(16000,8000).kernel and (8000,16000).kernel
  hifigan_golden.npz   real transformers SpeechT5HifiGan on the seeded synthetic weights/mel
  chunker_golden.npz   real AmendmentNetwork1 (HelloSippyRT.py:200-237) on seeded weights
  postnet_golden.npz   real transformers SpeechT5SpeechDecoderPostnet.postnet (modeling_speecht5.py:758-762, the call at
                       HelloSippyRTPipe.py:230) on seeded weights: one 32-frame call of 3 sessions and a 10-frame one
  infer_golden.npz     the real HelloSippyRTPipe.infer() + unbatch_and_dispatch() (HelloSippyRTPipe.py:191-259)
                       driven end to end in fp32 with a scripted front half (the AR decoder is out of scope,
                       so feat_out/prob_out/postnet are scripted to emit a fixed mel plan), and the real
                       G711Codec.encode over every dispatched chunk
"""
from __future__ import annotations

import hashlib
import json
import os
import types
import warnings

import numpy as np
import torch

from infernos_b200 import synth
from oracle import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def sha(a) -> str:
    if isinstance(a, torch.Tensor):
        a = a.contiguous().numpy()
    if isinstance(a, np.ndarray):
        a = a.tobytes()
    return hashlib.sha256(a).hexdigest()


def g711(G711Codec):
    warnings.simplefilter("ignore", DeprecationWarning)
    import audioop
    pcm = np.arange(-32768, 32768, dtype=np.int16)
    b = np.arange(256, dtype=np.uint8)
    t = dict(
        ulaw_enc=np.frombuffer(audioop.lin2ulaw(pcm.tobytes(), 2), dtype=np.uint8),
        alaw_enc=np.frombuffer(audioop.lin2alaw(pcm.tobytes(), 2), dtype=np.uint8),
        ulaw_dec=np.frombuffer(audioop.ulaw2lin(b.tobytes(), 2), dtype=np.int16),
        alaw_dec=np.frombuffer(audioop.alaw2lin(b.tobytes(), 2), dtype=np.int16),
    )
    np.savez_compressed(os.path.join(OUT, "g711_tables.npz"), **t)
    codec = G711Codec()
    g1_in = torch.linspace(-1.25, 1.25, 48001)
    g = torch.Generator().manual_seed(20240101)
    g2_in = (torch.rand(64, 8192, generator=g) * 2 - 1) * 0.9
    assert torch.equal(g2_in, synth.synth_audio(64, 8192))
    edge = torch.tensor([0.0, -0.0, 0.5, -0.5, 0.99999, -0.99999, 1.0, -1.0, 1.5, -1.5, 1e-5, -1e-5,
                         3.0518e-05, -3.0518e-05, 1.2207e-04, -1.2207e-04, 0.999969, 32766.5 / 32767.0])
    gold = {
        "sha256": {k: sha(v) for k, v in t.items()},
        "G1_encode_linspace": sha(codec.encode(g1_in)),
        "G2_encode_rand": sha(codec.encode(g2_in)),
        "G4_decode_all": sha(codec.decode(bytes(range(256)), resample=False).audio),
        "edge_in": [float(x) for x in edge],
        "edge_ulaw": list(codec.encode(edge)),
        "silence8": list(codec.silence(8)),
        "rtpmap": G711Codec.rtpmap(),
        "e2d_160_16000": codec.e2d_frames(160, 16000),
        "d2e_320_16000": codec.d2e_frames(320, 16000),
    }
    # decode with the reference's own 8k->16k path (G711.py:44-46 -> AudioChunk.resample)
    dec16 = codec.decode(bytes(range(256)) * 2, resample=True, sample_rate=16000).audio
    np.savez_compressed(os.path.join(OUT, "g711_decode16k.npz"), inp=np.frombuffer(bytes(range(256)) * 2, dtype=np.uint8),
                        out=dec16.numpy())
    with open(os.path.join(OUT, "g711_golden.json"), "w") as f:
        json.dump(gold, f, indent=1)


def taps():
    import torchaudio.transforms as T
    down = T.Resample(16000, 8000).kernel.reshape(28).numpy()
    up = T.Resample(8000, 16000).kernel.reshape(2, 15).numpy()
    x = synth.synth_audio(3, 1000)
    np.savez_compressed(os.path.join(OUT, "resample_taps.npz"), down=down, up=up,
                        x=x.numpy(), y_down=T.Resample(16000, 8000)(x).numpy(),
                        x_odd=x[:, :333].numpy(), y_down_odd=T.Resample(16000, 8000)(x[:, :333].contiguous()).numpy(),
                        y_up=T.Resample(8000, 16000)(x[:, :200].contiguous()).numpy())


def hifigan_and_chunker():
    vsd, csd = synth.hifigan_state_dict(), synth.chunker_state_dict()
    voc, chk = ref_shim.real_hifigan(vsd), ref_shim.real_chunker(csd)
    mel = synth.synth_mel(3, 12)
    mel_long = synth.synth_mel(1, 20, seed=11)
    with torch.no_grad():
        a = voc(mel)
        a_long = voc(mel_long)
        c = chk(mel, a)
    np.savez_compressed(os.path.join(OUT, "hifigan_golden.npz"), mel=mel.numpy(), audio=a.numpy(),
                        mel_long=mel_long.numpy(), audio_long=a_long.numpy(),
                        weights_sha=np.frombuffer(sha(torch.cat([v.flatten() for v in vsd.values()])).encode(), dtype=np.uint8))
    np.savez_compressed(os.path.join(OUT, "chunker_golden.npz"), mel=mel.numpy(), audio=a.numpy(), out=c.numpy(),
                        weights_sha=np.frombuffer(sha(torch.cat([v.flatten() for v in csd.values()])).encode(), dtype=np.uint8))


def postnet():
    psd = synth.postnet_state_dict()
    pn = ref_shim.real_postnet(psd)
    mel = synth.synth_mel(3, 32, seed=21)
    mel_short = synth.synth_mel(2, 10, seed=22)
    with torch.no_grad():
        out, out_short = pn.postnet(mel), pn.postnet(mel_short)
    np.savez_compressed(os.path.join(OUT, "postnet_golden.npz"), mel=mel.numpy(), out=out.numpy(), mel_short=mel_short.numpy(),
                        out_short=out_short.numpy(),
                        weights_sha=np.frombuffer(sha(torch.cat([v.flatten() for v in psd.values()])).encode(), dtype=np.uint8))
    print("postnet golden: |out - mel| rms", float((out - mel).pow(2).mean().sqrt()), "max", float((out - mel).abs().max()))


class _ScriptedFront:
    """Stands in for SpeechT5ForTextToSpeech inside the REAL infer(): emits a fixed mel plan,
    two frames per decoder step, and scripted stop logits.  postnet is the identity."""

    def __init__(self, plan: torch.Tensor, stop_step):
        B = plan.size(0)
        self.plan, self.stop_step, self.step = plan, stop_step, 0
        self.config = types.SimpleNamespace(num_mel_bins=80, reduction_factor=2)
        self.device = torch.device("cpu")
        me = self

        def prenet(output_sequence, speaker_embeddings):
            return torch.zeros(B, output_sequence.size(1), 4)

        def wrapped_decoder(**kw):
            return types.SimpleNamespace(last_hidden_state=torch.zeros(B, 1, 4), past_key_values=None)

        def feat_out(last):
            s = me.step
            return me.plan[:, 2 * s:2 * s + 2, :].reshape(B, 160)

        def prob_out(last):
            s = me.step
            me.step += 1
            logit = torch.full((B, 2), -20.0)
            for i, st in enumerate(me.stop_step):
                if s >= st:
                    logit[i] = 20.0
            return logit

        self.speecht5 = types.SimpleNamespace(decoder=types.SimpleNamespace(prenet=prenet, wrapped_decoder=wrapped_decoder))
        self.speech_decoder_postnet = types.SimpleNamespace(feat_out=feat_out, prob_out=prob_out, postnet=lambda x: x)


def infer_e2e(P, G711Codec):
    import threading
    import torchaudio.transforms as T
    vsd, csd = synth.hifigan_state_dict(), synth.chunker_state_dict()
    B, ncalls = 3, 4
    stop_step = [20, 37, 1000]            # decoder step at which each session's stop logit fires
    plan = synth.synth_mel(B, 32 * ncalls, seed=99)
    P.maybe_half = lambda x: x            # fp32 run (the product default is bf16, HelloSippyRTPipe.py:57)
    pp = P.HelloSippyRTPipe.__new__(P.HelloSippyRTPipe)
    pp.cuda_lock = threading.Lock()
    pp.model = _ScriptedFront(plan, stop_step)
    pp.vocoder = ref_shim.real_hifigan(vsd)
    RT = ref_shim.load()[0]
    pp.c_conf = RT.AmendmentNetwork1Config()
    pp.chunker = ref_shim.real_chunker(csd)
    pp.resampler = T.Resample(orig_freq=16000, new_freq=8000)
    pp.output_sr = 8000
    got = [[] for _ in range(B)]
    ended = [0] * B

    def mk(i):
        def cb(chunk):
            if chunk is None:
                ended[i] += 1
            else:
                got[i].append(chunk.clone())
        return cb

    st = P.HelloSippyPipeStateBatched.__new__(P.HelloSippyPipeStateBatched)
    st.dispatch = [mk(i) for i in range(B)]
    st.speaker_embeddings = torch.zeros(B, 512)
    st.encoder_last_hidden_state = torch.zeros(B, 5, 4)
    st.encoder_attention_mask = torch.ones(B, 5, dtype=torch.int)
    st.output_sequence = torch.zeros(B, 1, 80)
    st.past_key_values = None
    st.pre_frames = torch.zeros(B, 4, 80)
    st.starts_at = torch.tensor([1] * B)
    st.ends_at = torch.tensor([-1] * B)
    st.minlen, st.maxlen, st.idx = 0, 60, 0     # maxlen 60 ends session 2 by the length rule
    audios, ends_hist, idx_hist, emitted = [], [], [], []
    codec = G711Codec()
    with torch.no_grad():
        for c in range(ncalls):
            n_before = [len(g) for g in got]
            pp.infer(st)
            more = pp.unbatch_and_dispatch(st)
            audios.append(st.audio.clone())
            ends_hist.append(st.ends_at.clone())
            idx_hist.append(st.idx)
            emitted.append([sum(int(x.numel()) for x in got[i][n_before[i]:]) for i in range(B)])
            if not more:
                break
    out = dict(plan=plan.numpy(), stop_step=np.array(stop_step), audio=torch.stack(audios).numpy(),
               ends_at=torch.stack(ends_hist).numpy(), idx=np.array(idx_hist), emitted=np.array(emitted),
               ended=np.array(ended), ncalls_run=np.array(len(audios)), more_last=np.array(bool(more)))
    for i in range(B):
        full = torch.cat(got[i]) if got[i] else torch.zeros(0)
        out[f"session{i}_audio"] = full.numpy()
        out[f"session{i}_ulaw"] = np.frombuffer(b"".join(codec.encode(ch) for ch in got[i]), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "infer_golden.npz"), **out)
    print("infer golden: calls", len(audios), "emitted", emitted, "ended", ended, "ends_at", ends_hist[-1].tolist())


def main():
    os.makedirs(OUT, exist_ok=True)
    import sys
    if "--postnet-only" in sys.argv:          # adds postnet_golden.npz without regenerating the other fixtures
        postnet()
        return
    RT, P, G711Codec = ref_shim.load()
    g711(G711Codec)
    taps()
    hifigan_and_chunker()
    postnet()
    infer_e2e(P, G711Codec)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
