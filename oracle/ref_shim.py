"""ORACLE (test infrastructure): import recipe for the REAL reference modules.

Only usable where /root/reference exists (the build container).  Used by oracle/make_golden.py
to freeze fixtures under tests/golden/, and by tests marked `needs_reference` (skipped elsewhere).
Recipe from SURVEY.md App. B: five shims around un-installed third-party modules; the reference
sources are imported where they lie and are never copied.
"""
from __future__ import annotations

import os
import sys
import types

REF = os.environ.get("INFERNOS_REFERENCE", "/root/reference")
_loaded = None


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "HelloSippyTTSRT"))


def load():
    """Returns (RT module, P module, G711Codec class)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not present at {REF}")
    if REF not in sys.path:
        sys.path.insert(0, REF)

    def stub(n, **a):
        m = types.ModuleType(n)
        m.__dict__.update(a)
        sys.modules[n] = m
        return m

    stub("methodtools", lru_cache=lambda *a, **k: (lambda f: f))
    ap = stub("argostranslate")
    ap.package = stub("argostranslate.package")
    stub("argostranslate.translate", get_installed_languages=lambda: [])
    import transformers  # noqa: F401
    from transformers import SpeechT5ForTextToSpeech, SpeechT5HifiGan  # noqa: F401  (force lazy imports first)
    if "soundfile" not in sys.modules:
        stub("soundfile")
    import transformers.configuration_utils as cu
    orig = cu.PreTrainedConfig.__dict__["__init_subclass__"]
    cu.PreTrainedConfig.__init_subclass__ = classmethod(lambda cls, *a, **k: None)
    try:
        import HelloSippyTTSRT.HelloSippyRT as RT
    finally:
        cu.PreTrainedConfig.__init_subclass__ = orig
    import HelloSippyTTSRT.HelloSippyRTPipe as P
    from Core.Codecs.G711 import G711Codec
    _loaded = (RT, P, G711Codec)
    return _loaded


def real_hifigan(sd):
    from transformers import SpeechT5HifiGan, SpeechT5HifiGanConfig
    m = SpeechT5HifiGan(SpeechT5HifiGanConfig())
    missing, unexpected = m.load_state_dict(sd, strict=True)
    m.eval()
    return m


def real_chunker(sd):
    RT, _, _ = load()
    m = RT.AmendmentNetwork1(RT.AmendmentNetwork1Config())
    m.load_state_dict(sd, strict=True)
    m.eval()
    return m


def real_postnet(sd):
    """The real transformers SpeechT5SpeechDecoderPostnet (default SpeechT5Config) carrying the given `layers.*` tensors;
    feat_out / prob_out keep their default init (they are not on the post-net path)."""
    from transformers import SpeechT5Config
    from transformers.models.speecht5.modeling_speecht5 import SpeechT5SpeechDecoderPostnet
    m = SpeechT5SpeechDecoderPostnet(SpeechT5Config())
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith(("feat_out", "prob_out")) or k.endswith("num_batches_tracked") for k in missing), missing
    m.eval()
    return m
