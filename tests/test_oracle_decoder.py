"""Pins oracle/decoder.py (the CPU restatement of the AR front half) to the LIVE transformers modules driven through the
reference's own call sequence (/root/reference/HelloSippyTTSRT/HelloSippyRTPipe.py:195-223) on the same synthetic weights and the same
prenet dropout masks, and to the committed golden (tests/golden/decoder_golden.npz, made by oracle/make_golden_decoder.py)."""
import os

import numpy as np
import pytest
import torch

from infernos_b200 import synth
from oracle import decoder as odec

tr = pytest.importorskip("transformers")
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


from oracle.make_golden_decoder import live_run  # noqa: E402  (the reference loop on the real modules)


def _case(B=3, L=11, steps=5, seed=0):
    g = torch.Generator().manual_seed(seed)
    enc = synth.synth_encoder_states(B, L, seed=seed + 1)
    lens = torch.tensor([L, L - 4, L - 7][:B])
    enc_mask = (torch.arange(L)[None] < lens[:, None]).to(torch.int)
    speaker = synth.synth_speakers(B, seed=seed + 2)
    masks = (torch.rand(steps, 2, 256, generator=g) < 0.5).float()
    return enc, enc_mask, speaker, masks


def oracle_run(sd, enc, enc_mask, speaker, masks):
    st = odec.DecoderState(sd, enc, enc_mask, speaker)
    specs, probs = [], []
    with torch.no_grad():
        for s in range(masks.size(0)):
            sp, pr = odec.step(sd, st, masks[s])
            specs.append(sp); probs.append(pr)
    return torch.cat(specs, 1), torch.stack(probs, 1)


def test_oracle_matches_the_live_transformers_decoder():
    sd = synth.decoder_state_dict()
    enc, enc_mask, speaker, masks = _case()
    ref_s, ref_p = live_run(sd, enc, enc_mask, speaker, masks)
    got_s, got_p = oracle_run(sd, enc, enc_mask, speaker, masks)
    assert ref_s.shape == got_s.shape == (3, 10, 80)
    print("oracle vs live transformers decoder: max |d mel| %.2e, max |d prob| %.2e; mel rms %.2f" %
          (float((ref_s - got_s).abs().max()), float((ref_p - got_p).abs().max()), float(ref_s.std())))
    assert float((ref_s - got_s).abs().max()) < 2e-4 and float((ref_p - got_p).abs().max()) < 1e-5


def test_oracle_matches_the_committed_golden():
    p = os.path.join(G, "decoder_golden.npz")
    if not os.path.exists(p):
        pytest.skip("decoder_golden.npz has not been generated")
    d = np.load(p)
    sd = synth.decoder_state_dict()
    got_s, got_p = oracle_run(sd, torch.from_numpy(d["enc"]), torch.from_numpy(d["enc_mask"]), torch.from_numpy(d["speaker"]), torch.from_numpy(d["masks"]))
    assert np.abs(got_s.numpy() - d["mel"]).max() < 2e-4 and np.abs(got_p.numpy() - d["prob"]).max() < 1e-5
