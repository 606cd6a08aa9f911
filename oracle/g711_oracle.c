/*
 * ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * Plain-C restatement of the integer/byte half of the reference's TTS tail.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load the library built from this file.
 *
 * What it follows:
 *   - float -> int16 PCM        /root/reference/Core/Codecs/G711.py:27
 *                               clamp(x * 32767.0, -32768, 32767).to(int16)  (truncates toward zero)
 *   - int16 -> mu-law table     /root/reference/Core/Codecs/G711.py:7-12 builds it from CPython's
 *                               audioop.lin2ulaw (stdlib C, Sun g711.c lineage, 14-bit variant).  The
 *                               arithmetic is not under /root/reference; the published algorithm is
 *                               restated here and pinned exhaustively (all 65,536 / 256 inputs) against
 *                               audioop-generated tables whose sha256 are in tests/golden/g711_golden.json.
 *   - mu-law -> int16           /root/reference/Core/Codecs/G711.py:13-19 (audioop.ulaw2lin)
 *   - int16 -> float            /root/reference/Core/Codecs/G711.py:42   pcm.float() / 32767.0
 *   - A-law                     not in the reference (only PCMU/G722: /root/reference/SIP/InfernUAS.py:50);
 *                               north_star asks for it; oracle = audioop.lin2alaw / alaw2lin semantics.
 *   - 16k->8k resample          torchaudio/functional/functional.py:1416-1428 applied at
 *                               /root/reference/HelloSippyTTSRT/HelloSippyRTPipe.py:240:
 *                               y[j] = sum_i h[i] * xpad[2j+i], xpad = 13 zeros | x | 15 zeros, per call.
 *                               Here the sum is a defined pair of fmaf chains (o_down_dot below) so the CUDA
 *                               kernel can be bit-exact against it; torch's own summation order differs, and
 *                               tests bound that difference (|dPCM| <= 1).
 *   - 8k->16k resample          same module, orig 1 / new 2: 2 phases x 15 taps, pad (7, 8), stride 1,
 *                               used by /root/reference/Core/AudioChunk.py:19-24 after G711Codec.decode.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>

/* ---- float -> PCM16 (G711.py:27) ------------------------------------------------------- */
int16_t o_float_to_pcm16(float x)
{
    volatile float s = x * 32767.0f;           /* volatile: forbid contraction / excess precision */
    float c = s;
    if (c < -32768.0f) c = -32768.0f;
    if (c > 32767.0f) c = 32767.0f;
    if (c != c) return 0;                      /* NaN: torch's cast is undefined; pinned to 0 here */
    return (int16_t)c;                         /* C cast truncates toward zero, like torch .to(int16) */
}

/* ---- mu-law (14-bit Sun variant used by audioop) --------------------------------------- */
uint8_t o_ulaw_enc(int16_t s)
{
    int x = ((int)s) >> 2;                     /* arithmetic shift: 16-bit -> 14-bit */
    int mask = 0xFF;
    if (x < 0) { x = -x; mask = 0x7F; }
    if (x > 8159) x = 8159;                    /* clip */
    x += 0x21;                                 /* bias (33) */
    int seg = 0;                               /* seg = bit_length(x) - 6, x in [0x21, 0x2000] */
    for (int v = x >> 6; v; v >>= 1) seg++;
    int code = (seg >= 8) ? 0x7F : ((seg << 4) | ((x >> (seg + 1)) & 0xF));
    return (uint8_t)(code ^ mask);
}

int16_t o_ulaw_dec(uint8_t u)
{
    int v = (~u) & 0xFF;
    int t = (((v & 0x0F) << 3) + 0x84) << ((v & 0x70) >> 4);
    return (int16_t)((v & 0x80) ? (0x84 - t) : (t - 0x84));
}

/* ---- A-law (13-bit Sun variant used by audioop) ---------------------------------------- */
uint8_t o_alaw_enc(int16_t s)
{
    static const int seg_end[8] = {0x1F, 0x3F, 0x7F, 0xFF, 0x1FF, 0x3FF, 0x7FF, 0xFFF};
    int x = ((int)s) >> 3;
    int mask;
    if (x >= 0) mask = 0xD5; else { mask = 0x55; x = -x - 1; }
    int seg = 8;
    for (int i = 0; i < 8; i++) if (x <= seg_end[i]) { seg = i; break; }
    int code;
    if (seg >= 8) code = 0x7F;
    else code = (seg << 4) | (((seg < 2) ? (x >> 1) : (x >> seg)) & 0xF);
    return (uint8_t)(code ^ mask);
}

int16_t o_alaw_dec(uint8_t a)
{
    int v = a ^ 0x55;
    int t = (v & 0x0F) << 4;
    int seg = (v & 0x70) >> 4;
    if (seg == 0) t += 8;
    else if (seg == 1) t += 0x108;
    else t = (t + 0x108) << (seg - 1);
    return (int16_t)((v & 0x80) ? t : -t);
}

float o_pcm16_to_float(int16_t s) { return (float)s / 32767.0f; }   /* G711.py:42 */

/* ---- bulk helpers ------------------------------------------------------------------------ */
void o_encode_pcm16(const int16_t *pcm, size_t n, int law, uint8_t *out)
{
    for (size_t i = 0; i < n; i++) out[i] = law ? o_alaw_enc(pcm[i]) : o_ulaw_enc(pcm[i]);
}

void o_encode_f32(const float *x, size_t n, int law, uint8_t *out)
{
    for (size_t i = 0; i < n; i++) {
        int16_t s = o_float_to_pcm16(x[i]);
        out[i] = law ? o_alaw_enc(s) : o_ulaw_enc(s);
    }
}

void o_f32_to_pcm16(const float *x, size_t n, int16_t *out)
{
    for (size_t i = 0; i < n; i++) out[i] = o_float_to_pcm16(x[i]);
}

void o_decode_pcm16(const uint8_t *in, size_t n, int law, int16_t *out)
{
    for (size_t i = 0; i < n; i++) out[i] = law ? o_alaw_dec(in[i]) : o_ulaw_dec(in[i]);
}

void o_decode_f32(const uint8_t *in, size_t n, int law, float *out)
{
    for (size_t i = 0; i < n; i++) out[i] = o_pcm16_to_float(law ? o_alaw_dec(in[i]) : o_ulaw_dec(in[i]));
}

/* One output of the 16k->8k filter.  torchaudio's taps 0 and 27 are -0 by construction (the Hann window's zero) and are
 * skipped; the sum is DEFINED as two ascending fmaf chains from +0 -- e over the odd taps 1,3,..,25 and o over the even taps
 * 2,4,..,26 -- and one rounded add y = e + o.  (This is the order a packed two-lane fused multiply-add per tap pair produces;
 * the CUDA kernels are bit-exact against it.) */
static float o_down_dot(const float *xr, size_t L, size_t j, const float *h)
{
    float e = 0.0f, o = 0.0f;
    for (int i = 1; i < 27; i++) {
        long p = (long)(2 * j) + i - 13;
        float v = (p >= 0 && p < (long)L) ? xr[p] : 0.0f;
        if (i & 1) e = fmaf(h[i], v, e); else o = fmaf(h[i], v, o);
    }
    return e + o;
}

/* ---- 16k -> 8k: 28 taps, stride 2, zero pad (13, 15) per row ------------------------------
 * h: the 28 fp32 taps of torchaudio.transforms.Resample(16000, 8000).kernel (passed in; the tests
 * pass the values frozen in tests/golden/resample_taps.npz).  Rows are independent (per session
 * per call).  Output length per row = ceil(L/2). */
void o_resample_2to1(const float *x, size_t rows, size_t L, const float *h, float *y)
{
    size_t Lo = (L + 1) / 2;
    for (size_t r = 0; r < rows; r++) {
        const float *xr = x + r * L;
        float *yr = y + r * Lo;
        for (size_t j = 0; j < Lo; j++) {
            yr[j] = o_down_dot(xr, L, j, h);
        }
    }
}

/* fused reference: resample then encode (what the product's fused kernel must equal bit-for-bit) */
void o_resample_2to1_encode(const float *x, size_t rows, size_t L, const float *h, int law, uint8_t *out)
{
    size_t Lo = (L + 1) / 2;
    for (size_t r = 0; r < rows; r++) {
        const float *xr = x + r * L;
        for (size_t j = 0; j < Lo; j++) {
            int16_t s = o_float_to_pcm16(o_down_dot(xr, L, j, h));
            out[r * Lo + j] = law ? o_alaw_enc(s) : o_ulaw_enc(s);
        }
    }
}

/* ---- 8k -> 16k: kernel (2, 15), pad (7, 8), stride 1; out[2j+p] = sum_i h[p][i] * xpad[j+i] -- */
void o_resample_1to2(const float *x, size_t rows, size_t L, const float *h /* [2][15] */, float *y)
{
    for (size_t r = 0; r < rows; r++) {
        const float *xr = x + r * L;
        float *yr = y + r * 2 * L;
        for (size_t j = 0; j < L; j++) {
            for (int p = 0; p < 2; p++) {
                float acc = 0.0f;
                for (int i = 0; i < 15; i++) {
                    long q = (long)j + i - 7;
                    float v = (q >= 0 && q < (long)L) ? xr[q] : 0.0f;
                    acc = fmaf(h[p * 15 + i], v, acc);
                }
                yr[2 * j + p] = acc;
            }
        }
    }
}

/* decode + optional 8k->16k upsample (G711.py:34-47 + AudioChunk.py:19-24) */
void o_decode_upsample(const uint8_t *in, size_t rows, size_t L, int law, const float *h, float *tmp, float *y)
{
    o_decode_f32(in, rows * L, law, tmp);
    o_resample_1to2(tmp, rows, L, h, y);
}
