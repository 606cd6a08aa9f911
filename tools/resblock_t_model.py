#!/usr/bin/env python
"""Index-exact numpy model of the stacked-output ResBlock kernel (infernos_b200/csrc/conv_resblock_t.cu), float64, no operand rounding.

Every address the kernel computes is computed here the same way -- the slab plan (S, H, V, off, lim), the de-interleaved operand buffers
[chunk][block][row][8 ch] with their dilation-dependent row mapping, the extended-tap loop (block (j' - half) mod 4, row offset
((j' - half) div 4) * d, N sub-range rlo..rhi, reversed-tap slot window), the accumulator layouts of X and T1, the bias pre-load of T1 -- and
the result is compared with torch.conv1d.  CPU only: the check that the algebra is right before any GPU time is spent; tests/ runs it against
the library's own plan function.  Not part of the product."""
import numpy as np
import torch
import torch.nn.functional as F

KTG, KTBR = 10, 153


def plan(k, dil, T, post, align4=False):
    """mirror of resblock_t_plan()"""
    half = (k - 1) // 2
    cap = [4 * d * (128 // d) for d in dil]
    H = half * sum(d + 1 for d in dil) + (3 if post else 0)
    g, acc, S = [], 0, 512
    for i in range(3):
        g.append(acc)
        S = min(S, cap[i] + 2 * acc)
        acc += half * (dil[i] + 1)
    off = [max(0, (S - cap[i] + 1) // 2) for i in range(3)]
    lim = [min(cap[i], S - off[i]) for i in range(3)]
    for i in range(3):
        assert off[i] <= g[i] and S - off[i] - lim[i] <= g[i]
    if align4:
        H = (H + 3) & ~3
    vmax = ((S - 2 * H) & ~3) if align4 else S - 2 * H
    assert vmax >= 64
    tiles = -(-T // vmax)
    V = -(-T // tiles)
    if align4:
        V = (V + 3) & ~3
    return dict(S=S, H=H, V=V, tiles=tiles, off=off, lim=lim)


def rbt_map(r, d, off, lim):
    rr = r - off
    if rr < 0 or rr >= lim:
        return -1
    u, rho = divmod(rr, d)
    return (u & 3) * KTBR + KTG + (u >> 2) * d + rho


def lrelu(v, s):
    return np.maximum(v, s * v)


def write_rows(buf, row16, vals):
    """buf [4 chunks][4*KTBR rows][8]; vals (32,)"""
    if row16 >= 0:
        buf[:, row16, :] = vals.reshape(4, 8)


def mma_conv(buf, slots, acc, k, d):
    """the extended-tap loop of the MMA thread.  buf: operand buffer, slots [k][32 co][32 ci] (tap k-1-q in slot q), acc [128][128] += ..."""
    half = (k - 1) // 2
    have, freed = -1, 0
    for jp in range(k + 2, -1, -1):
        rlo, rhi = max(0, jp - (k - 1)), min(3, jp)
        qlo, qhi = k - 1 - jp + rlo, k - 1 - jp + rhi
        assert 0 <= qlo <= qhi < k and freed <= qlo
        have = max(have, qhi)
        e = jp - half
        a, bq = e >> 2, e & 3
        row0 = bq * KTBR + KTG + a * d
        assert 0 <= KTG + a * d and KTG + a * d + 128 <= KTBR
        A = buf[:, row0:row0 + 128, :].transpose(1, 0, 2).reshape(128, 32)              # [m][ci]
        B = slots[qlo:qhi + 1].reshape(-1, 32)                                            # [(r', co)][ci]
        acc[:, rlo * 32:(rhi + 1) * 32] += A @ B.T
        nqlo = (k - jp + max(0, jp - k)) if jp > 0 else k
        freed = max(freed, nqlo)
    assert have == k - 1 and freed == k


POISON = True          # the kernel does not clear its operand buffers: start them as NaN, valid outputs must not notice


def pack_convT(wT):
    """tail.cu:pack_convT -- torch ConvTranspose1d(k8, s4, p2) weight [Cin][Cout][8] -> [3 taps][4*Cout (phase, co)][Cin]"""
    Cin, Cout, _ = wT.shape
    q = np.zeros((3, 4 * Cout, Cin))
    for ph in range(4):
        for tap in range(3):
            kk = ph + 2 if tap == 1 else (ph + 6 if tap == 0 else ph - 2)
            if 0 <= kk <= 7:
                q[tap, ph * Cout:(ph + 1) * Cout, :] = wT[:, :, kk].T
    return q


def upsample_slab(u_in, wq, bT, t_base, upT):
    """the UP prologue of the kernel: operand rows [t_base/4 - 1, t_base/4 + 129) x 64 channels, X pre-loaded with the bias, six weight boxes
    (tap, 32-channel half): tap 0 feeds columns 0..63, tap 1 all 128, tap 2 columns 64..127.  -> X [128][128]"""
    assert t_base % 4 == 0
    t0 = t_base // 4
    U = np.zeros((130, 64))
    for row in range(130):
        tin = t0 - 1 + row
        if 0 <= tin < upT:
            U[row] = u_in[tin]
    X = np.tile(bT, 4)[None, :].repeat(128, axis=0).copy()
    for b in range(6):
        tap, h = b >> 1, b & 1
        col, nn = (64 if tap == 2 else 0), (128 if tap == 1 else 64)
        A = U[tap:tap + 128, 32 * h:32 * h + 32]                       # [m][ci]
        B = wq[tap, col:col + nn, 32 * h:32 * h + 32]                  # [(phase, co)][ci]
        X[:, col:col + nn] += A @ B.T
    return X


def run_slab(x, ws, bs, k, dil, slope, pl, t_base, T, up=None):
    """one CTA.  x (T, 32) float64 window; returns X rows [S][32] after the three pairs (bias included)."""
    S, off, lim = pl["S"], pl["off"], pl["lim"]
    A1 = np.full((4, 4 * KTBR, 8), np.nan if POISON else 0.0)
    A2 = np.full((4, 4 * KTBR, 8), np.nan if POISON else 0.0)
    X = np.zeros((128, 128))
    T1 = np.zeros((128, 128))
    inside = lambda r: r < S and 0 <= t_base + r < T
    slots = [np.stack([ws[c][:, :, k - 1 - q] for q in range(k)]) for c in range(6)]    # [q][co][ci]
    cbias = np.cumsum(np.stack([bs[1], bs[3], bs[5]]), axis=0)
    # load (or, UP: the stage's upsampler computed into X)
    if up is not None:
        X = upsample_slab(up[0], up[1], up[2], t_base, T // 4)
    for m in range(128):
        for q in range(4):
            r = 4 * m + q
            if up is None:
                v = x[t_base + r] if inside(r) else np.zeros(32)
                X[m, q * 32:(q + 1) * 32] = v
            else:
                v = X[m, q * 32:(q + 1) * 32]
            T1[m, q * 32:(q + 1) * 32] = bs[0]
            write_rows(A1, rbt_map(r, dil[0], off[0], lim[0]), lrelu(v, slope) if inside(r) else np.zeros(32))
    for i in range(3):
        d = dil[i]
        mma_conv(A1, slots[2 * i], T1, k, d)
        for m in range(128):
            v, rho = divmod(m, d)
            for q in range(4):
                acc = T1[m, q * 32:(q + 1) * 32].copy()
                if i < 2:
                    T1[m, q * 32:(q + 1) * 32] = bs[2 * (i + 1)]
                rr = d * (4 * v + q) + rho
                r = rr + off[i]
                valid = rr < lim[i]
                row16 = ((r & 3) * KTBR + KTG + (r >> 2)) if valid else -1
                keep = valid and 0 <= t_base + r < T
                write_rows(A2, row16, lrelu(acc, slope) if keep else np.zeros(32))
        mma_conv(A2, slots[2 * i + 1], X, k, 1)
        if i < 2:
            for m in range(128):
                for q in range(4):
                    r = 4 * m + q
                    acc = X[m, q * 32:(q + 1) * 32] + cbias[i]
                    write_rows(A1, rbt_map(r, dil[i + 1], off[i + 1], lim[i + 1]), lrelu(acc, slope) if inside(r) else np.zeros(32))
    out = np.zeros((512, 32))
    for m in range(128):
        for q in range(4):
            out[4 * m + q] = X[m, q * 32:(q + 1) * 32] + cbias[2]
    return out


def resblock_model(x, ws, bs, k, dil, slope, post=False, pl=None, up=None):
    """x (T, 32); ws [6] (32, 32, k); bs [6] (32,) -> (T, 32).  up = (u_in (T/4, 64), packed weights [3][128][64], bias (32,)): x is not read
    but computed from the upsampler's input, as the kernel's UP variant does."""
    T = x.shape[0]
    pl = pl or plan(k, dil, T, post, align4=up is not None)
    out = np.zeros((T, 32))
    for tile in range(pl["tiles"]):
        t_base = tile * pl["V"] - pl["H"]
        slab = run_slab(x, ws, bs, k, dil, slope, pl, t_base, T, up)
        for r in range(pl["H"], pl["H"] + pl["V"]):
            if t_base + r < T:
                out[t_base + r] = slab[r]
    return out


def resblock_ref(x, ws, bs, k, dil, slope):
    h = torch.from_numpy(x).T[None]
    for i, d in enumerate(dil):
        y = F.conv1d(F.leaky_relu(h, slope), torch.from_numpy(ws[2 * i]), torch.from_numpy(bs[2 * i]), dilation=d, padding=(k - 1) * d // 2)
        h = h + F.conv1d(F.leaky_relu(y, slope), torch.from_numpy(ws[2 * i + 1]), torch.from_numpy(bs[2 * i + 1]), padding=(k - 1) // 2)
    return h[0].T.numpy()


def check_up(k, dil, T, post=False, seed=0, pl=None):
    """the UP variant: x = ConvTranspose1d(k8, s4, p2)(u_in) computed inside the model, against torch's conv_transpose1d + ResBlock"""
    g = np.random.default_rng(seed)
    u_in = g.standard_normal((T // 4, 64))
    wT = g.standard_normal((64, 32, 8)) / 16.0
    bT = g.standard_normal(32) * 0.1
    ws = [g.standard_normal((32, 32, k)) / (32 * k) ** 0.5 for _ in range(6)]
    bs = [g.standard_normal(32) * 0.1 for _ in range(6)]
    x = F.conv_transpose1d(torch.from_numpy(u_in).T[None], torch.from_numpy(wT), torch.from_numpy(bT), stride=4, padding=2)[0].T.numpy()
    assert x.shape == (T, 32)
    got = resblock_model(np.full_like(x, np.nan), ws, bs, k, dil, 0.1, post, pl, up=(u_in, pack_convT(wT), bT))
    ref = resblock_ref(x, ws, bs, k, dil, 0.1)
    return float(np.abs(got - ref).max())


def check(k, dil, T, post=False, seed=0, pl=None):
    g = np.random.default_rng(seed)
    x = g.standard_normal((T, 32))
    ws = [g.standard_normal((32, 32, k)) / (32 * k) ** 0.5 for _ in range(6)]
    bs = [g.standard_normal(32) * 0.1 for _ in range(6)]
    got = resblock_model(x, ws, bs, k, dil, 0.1, post, pl)
    ref = resblock_ref(x, ws, bs, k, dil, 0.1)
    return float(np.abs(got - ref).max())


if __name__ == "__main__":
    for k, dil, T, post in ((3, (1, 3, 5), 700, False), (7, (1, 3, 5), 1100, False), (11, (1, 3, 5), 1300, True), (11, (1, 3, 5), 393, False),
                            (3, (1, 3, 5), 1, False), (7, (1, 3, 5), 100, False), (5, (2, 1, 4), 900, False), (11, (1, 3, 5), 500, False),
                            (11, (1, 3, 5), 501, False), (3, (1, 3, 5), 3072, False)):
        print(f"k={k} dil={dil} T={T} post={post}: plan {plan(k, dil, T, post)}  max |diff| = {check(k, dil, T, post):.2e}")
    for k, dil, T, post in ((3, (1, 3, 5), 3072, False), (7, (1, 3, 5), 1100, False), (11, (1, 3, 5), 1300, True), (11, (1, 3, 5), 392, True), (7, (1, 3, 5), 4, False)):
        print(f"UP k={k} dil={dil} T={T} post={post}: plan {plan(k, dil, T, post, True)}  max |diff| = {check_up(k, dil, T, post):.2e}")
