#!/usr/bin/env python
"""Round-2 groundwork: the thin HiFiGAN convolutions with the operand roles swapped.

Today a C = 32 conv is M = 128 time steps x N = 32 output channels per tcgen05.mma: the tensor core spends 32 cycles fetching the 4 KB
activation slice for 8 cycles of math (tools/mma_rate.cu).  Here the WEIGHTS become the M = 128 operand -- R = 4 time-shifted copies of the
32 output channels stacked into a block-Toeplitz matrix -- and the activations the N = 256 operand, so every MMA runs at full rate; the price
is K = (k + R - 1) * C instead of k * C.  A dilated conv is R-stacked inside each residue class of time modulo the dilation.

    out[co][t] = sum_j sum_ci W[co][ci][j] * x[ci][t + j*d - pad],   pad = d*(k-1)/2
    t = d*u + rho, u = R*v + r:
    D[(r, co)][(rho, v)] = sum_{j', ci} A[(r, co)][(j', ci)] * B[(j', ci)][(rho, v)]
    A[(r, co)][(j', ci)] = W[co][ci][j' - r] if 0 <= j' - r < k else 0                      (128 x (k + R - 1)*C, built once per layer)
    B[(j', ci)][(rho, v)] = x[ci][d*(R*v + j' - (k-1)/2) + rho]                             (zero outside the window)

This script checks the index algebra against torch.conv1d in float64 and prints the tensor-pipe cycle counts of both formulations
(cycles per tcgen05.mma M = 128, K = 16: 32 + N/4 for N <= 128, N/2 above; measured, profiles/r1b_mma_rate_microbench.txt).
CPU only; not part of the product."""
import torch
import torch.nn.functional as F

R = 4


def toeplitz_conv(x, W, d):
    C_out, C_in, k = W.shape
    T = x.shape[1]
    half = (k - 1) // 2
    A = torch.zeros(R * C_out, (k + R - 1) * C_in, dtype=x.dtype)
    for r in range(R):
        for j in range(k):
            A[r * C_out:(r + 1) * C_out, (r + j) * C_in:(r + j + 1) * C_in] = W[:, :, j]
    out = torch.zeros(C_out, T, dtype=x.dtype)
    for rho in range(d):
        nu = (T - rho + d - 1) // d                      # samples of this residue class
        nv = (nu + R - 1) // R
        B = torch.zeros((k + R - 1) * C_in, nv, dtype=x.dtype)
        for jp in range(k + R - 1):
            for v in range(nv):
                t = d * (R * v + jp - half) + rho
                if 0 <= t < T:
                    B[jp * C_in:(jp + 1) * C_in, v] = x[:, t]
        D = A @ B                                        # the MMA: M = R*C_out, N = nv, K = (k+R-1)*C_in
        for r in range(R):
            for v in range(nv):
                t = d * (R * v + r) + rho
                if t < T:
                    out[:, t] = D[r * C_out:(r + 1) * C_out, v]
    return out


def main():
    torch.manual_seed(0)
    for C, k, d, T in ((32, 3, 1, 200), (32, 7, 3, 333), (32, 11, 5, 512), (64, 11, 1, 100)):
        x = torch.randn(C, T, dtype=torch.float64)
        W = torch.randn(C, C, k, dtype=torch.float64)
        ref = F.conv1d(x[None], W, dilation=d, padding=d * (k - 1) // 2)[0]
        got = toeplitz_conv(x, W, d)
        print(f"C={C} k={k} d={d} T={T}: max |diff| = {float((ref - got).abs().max()):.2e}")
        assert torch.allclose(ref, got, atol=1e-10)
    print("\ntensor-pipe cycles per 1,024 output time steps (all C output channels):")
    for C in (32, 64):
        per_mma_now = 32 + C / 4                                         # M = 128 rows, N = C
        for k in (3, 7, 11):
            now = (1024 / 128) * k * (C / 16) * per_mma_now
            Rr = 128 // C
            new = (1024 / (Rr * 256)) * ((k + Rr - 1) * C / 16) * 128    # M = 128 = Rr*C, N = 256
            print(f"  C={C} k={k}: time-as-M {now:7.0f}   weights-as-M (R={Rr}) {new:7.0f}   x{now / new:.2f}")


if __name__ == "__main__":
    main()
