"""ORACLE (test infrastructure): ctypes front-end of oracle/g711_oracle.c plus vectorised numpy
closed forms of the same G.711 arithmetic.  Follows
/root/reference/Core/Codecs/G711.py:7-47.  Never imported by the product package."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle_g711.so")
    src = os.path.join(_HERE, "g711_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle_g711.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        vp, sz, it = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        L.o_encode_pcm16.argtypes = [vp, sz, it, vp]
        L.o_encode_f32.argtypes = [vp, sz, it, vp]
        L.o_f32_to_pcm16.argtypes = [vp, sz, vp]
        L.o_decode_pcm16.argtypes = [vp, sz, it, vp]
        L.o_decode_f32.argtypes = [vp, sz, it, vp]
        L.o_resample_2to1.argtypes = [vp, sz, sz, vp, vp]
        L.o_resample_2to1_encode.argtypes = [vp, sz, sz, vp, it, vp]
        L.o_resample_1to2.argtypes = [vp, sz, sz, vp, vp]
        for f in ("o_encode_pcm16", "o_encode_f32", "o_f32_to_pcm16", "o_decode_pcm16", "o_decode_f32",
                  "o_resample_2to1", "o_resample_2to1_encode", "o_resample_1to2"):
            getattr(L, f).restype = None
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


LAW_ULAW, LAW_ALAW = 0, 1


def encode_pcm16(pcm: np.ndarray, law: int = 0) -> np.ndarray:
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    out = np.empty(pcm.shape, dtype=np.uint8)
    lib().o_encode_pcm16(_p(pcm), pcm.size, law, _p(out))
    return out


def encode_f32(x: np.ndarray, law: int = 0) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.shape, dtype=np.uint8)
    lib().o_encode_f32(_p(x), x.size, law, _p(out))
    return out


def f32_to_pcm16(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.shape, dtype=np.int16)
    lib().o_f32_to_pcm16(_p(x), x.size, _p(out))
    return out


def decode_pcm16(b: np.ndarray, law: int = 0) -> np.ndarray:
    b = np.ascontiguousarray(b, dtype=np.uint8)
    out = np.empty(b.shape, dtype=np.int16)
    lib().o_decode_pcm16(_p(b), b.size, law, _p(out))
    return out


def decode_f32(b: np.ndarray, law: int = 0) -> np.ndarray:
    b = np.ascontiguousarray(b, dtype=np.uint8)
    out = np.empty(b.shape, dtype=np.float32)
    lib().o_decode_f32(_p(b), b.size, law, _p(out))
    return out


def resample_2to1(x: np.ndarray, taps: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    assert x.ndim == 2
    taps = np.ascontiguousarray(taps, dtype=np.float32).reshape(28)
    rows, L = x.shape
    y = np.empty((rows, (L + 1) // 2), dtype=np.float32)
    lib().o_resample_2to1(_p(x), rows, L, _p(taps), _p(y))
    return y


def resample_2to1_encode(x: np.ndarray, taps: np.ndarray, law: int = 0) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    assert x.ndim == 2
    taps = np.ascontiguousarray(taps, dtype=np.float32).reshape(28)
    rows, L = x.shape
    y = np.empty((rows, (L + 1) // 2), dtype=np.uint8)
    lib().o_resample_2to1_encode(_p(x), rows, L, _p(taps), law, _p(y))
    return y


def resample_1to2(x: np.ndarray, taps: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    assert x.ndim == 2
    taps = np.ascontiguousarray(taps, dtype=np.float32).reshape(30)
    rows, L = x.shape
    y = np.empty((rows, 2 * L), dtype=np.float32)
    lib().o_resample_1to2(_p(x), rows, L, _p(taps), _p(y))
    return y


# ---- numpy closed forms (second, independent statement of the same tables) -----------------
def np_ulaw_enc(pcm: np.ndarray) -> np.ndarray:
    x = pcm.astype(np.int32) >> 2
    neg = x < 0
    x = np.where(neg, -x, x)
    x = np.minimum(x, 8159) + 0x21
    seg = np.floor(np.log2(x)).astype(np.int32) + 1 - 6
    code = np.where(seg >= 8, 0x7F, (seg << 4) | ((x >> (seg + 1)) & 0xF))
    return (code ^ np.where(neg, 0x7F, 0xFF)).astype(np.uint8)


def np_alaw_enc(pcm: np.ndarray) -> np.ndarray:
    x = pcm.astype(np.int32) >> 3
    neg = x < 0
    x = np.where(neg, -x - 1, x)
    bl = np.where(x > 0, np.floor(np.log2(np.maximum(x, 1))).astype(np.int32) + 1, 0)
    seg = np.maximum(0, bl - 5)
    code = np.where(seg >= 8, 0x7F, (seg << 4) | (np.where(seg < 2, x >> 1, x >> np.minimum(seg, 31)) & 0xF))
    return (code ^ np.where(neg, 0x55, 0xD5)).astype(np.uint8)
