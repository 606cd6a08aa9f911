"""Freezes tests/golden/decoder_golden.npz from the LIVE transformers SpeechT5 decoder modules, driven exactly like the reference's while
loop (/root/reference/HelloSippyTTSRT/HelloSippyRTPipe.py:195-223), on infernos_b200.synth.decoder_state_dict() and seeded prenet dropout
masks.  Run in the build container:  python -m oracle.make_golden_decoder

ORACLE / test infrastructure: nothing under infernos_b200/ imports this."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def live_run(sd, enc, enc_mask, speaker, masks):
    import transformers as tr
    from oracle import decoder as odec
    """The reference loop on the real modules; `_consistent_dropout` is given the test's masks for the last position (the only one used)."""
    cfg = tr.SpeechT5Config()
    with torch.device("meta"):
        model = tr.SpeechT5ForTextToSpeech(cfg)
    dec, post = model.speecht5.decoder.to_empty(device="cpu"), model.speech_decoder_postnet.to_empty(device="cpu")
    dec.load_state_dict({k[len("speecht5.decoder."):]: v for k, v in sd.items() if k.startswith("speecht5.decoder.")}, strict=False)
    post.load_state_dict({k[len("speech_decoder_postnet."):]: v for k, v in sd.items() if k.startswith("speech_decoder_postnet.")}, strict=False)
    dec.prenet.encode_positions.pe = odec.position_table(cfg.max_speech_positions)[None]       # non-persistent buffer: rebuilt after to_empty
    dec.eval(); post.eval()
    calls = {"step": 0, "layer": 0}

    def consistent_dropout(inputs_embeds, p):
        m = torch.ones_like(inputs_embeds[0])
        m[-1] = masks[calls["step"], calls["layer"]]
        calls["layer"] += 1
        return torch.where(m.unsqueeze(0).repeat(inputs_embeds.size(0), 1, 1) == 1, inputs_embeds, 0) * 1 / (1 - p)
    dec.prenet._consistent_dropout = consistent_dropout
    B = enc.size(0)
    output_sequence = enc.new_zeros(B, 1, 80)
    past, specs, probs = None, [], []
    with torch.no_grad():
        for s in range(masks.size(0)):
            calls["step"], calls["layer"] = s, 0
            hs = dec.prenet(output_sequence, speaker)[:, -1:]
            out = dec.wrapped_decoder(hidden_states=hs, attention_mask=None, encoder_hidden_states=enc, encoder_attention_mask=enc_mask,
                                      past_key_values=past, use_cache=True, output_attentions=False, return_dict=True)
            last = out.last_hidden_state[:, -1, :]
            past = out.past_key_values
            spectrum = post.feat_out(last).view(B, 2, 80)
            output_sequence = torch.cat((output_sequence, spectrum[:, -1:, :]), dim=1)
            specs.append(spectrum)
            probs.append(post.prob_out(last).sigmoid())
    return torch.cat(specs, 1), torch.stack(probs, 1)




def main():
    from infernos_b200 import synth
    B, L, steps = 3, 11, 16
    g = torch.Generator().manual_seed(0)
    enc = synth.synth_encoder_states(B, L, seed=1)
    lens = torch.tensor([L, L - 4, L - 7])
    enc_mask = (torch.arange(L)[None] < lens[:, None]).to(torch.int)
    speaker = synth.synth_speakers(B, seed=2)
    masks = (torch.rand(steps, 2, 256, generator=g) < 0.5).float()
    mel, prob = live_run(synth.decoder_state_dict(), enc, enc_mask, speaker, masks)
    out = os.path.join(ROOT, "tests", "golden", "decoder_golden.npz")
    np.savez_compressed(out, enc=enc.numpy(), enc_mask=enc_mask.numpy(), speaker=speaker.numpy(), masks=masks.numpy(), mel=mel.numpy(), prob=prob.numpy())
    print(out, mel.shape, float(mel.std()))


if __name__ == "__main__":
    main()
