"""CPU test of SpeechT5Frontend (the default front half the engine builds around a transformers SpeechT5ForTextToSpeech): the
reference casts every module AND every speaker vector to bf16 (maybe_half, /root/reference/HelloSippyTTSRT/HelloSippyRTPipe.py:57,75,166-168),
while x-vectors and the engine's default speaker arrive as fp32 — start() has to bring them to the model's dtype or the first decoder
step dies in speaker_embeds_layer."""
import types

import pytest
import torch

tr = pytest.importorskip("transformers")


def _tiny(dtype):
    cfg = tr.SpeechT5Config(vocab_size=40, hidden_size=32, encoder_layers=1, encoder_attention_heads=2, encoder_ffn_dim=64,
                            decoder_layers=1, decoder_attention_heads=2, decoder_ffn_dim=64, speech_decoder_prenet_units=16,
                            speech_decoder_prenet_layers=2, speaker_embedding_dim=16, speech_decoder_postnet_units=16,
                            speech_decoder_postnet_layers=2, max_text_positions=64, max_speech_positions=256, num_mel_bins=80, reduction_factor=2)
    torch.manual_seed(0)
    return tr.SpeechT5ForTextToSpeech(cfg).to(dtype).eval()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_start_and_steps_on_a_tiny_random_speecht5(dtype):
    from infernos_b200.HelloSippyTTSRT.HelloSippyRTPipe import SpeechT5Frontend
    fe = SpeechT5Frontend(_tiny(dtype), processor=None)
    assert fe.reduction_factor == 2 and fe.num_mel_bins == 80
    states = []
    for n in (5, 9):
        ids = torch.randint(4, 40, (1, n))
        states.append(types.SimpleNamespace(inputs=ids, encoder_attention_mask=torch.ones_like(ids, dtype=torch.int),
                                            speaker_embeddings=torch.randn(1, 16)))          # fp32, like get_voice() / torch.zeros(1, 512)
    st = types.SimpleNamespace()
    fe.start(st, states)
    assert st.speaker_embeddings.dtype == dtype and st.encoder_last_hidden_state.shape[:2] == (2, 9)
    lo, hi = fe.length_bounds(st, 0.0, 20.0)
    assert (lo, hi) == (0, 90)
    for k in range(3):
        spectrum, prob = fe.step(st)
        assert spectrum.shape == (2, 2, 80) and prob.shape == (2, 2) and spectrum.dtype == dtype
        assert st.output_sequence.shape == (2, 2 + k, 80)
    out = fe.postnet(torch.cat([spectrum, spectrum], dim=1))
    assert out.shape == (2, 4, 80)
