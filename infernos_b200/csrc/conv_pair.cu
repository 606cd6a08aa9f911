// Fused HiFiGAN ResBlock pair on tcgen05:   x' = x + conv2( lrelu( conv1( lrelu(x) ) ) )      (modeling_speecht5.py:2954-2962)
//
// One CTA computes conv1 for M1 = 128*mt rows (one window, consecutive time steps) into TMEM, turns the accumulator into
// the bf16 operand of conv2 *in shared memory* (bias, leaky-ReLU, zero outside the window = conv2's own "same" padding),
// runs conv2 from there into a second TMEM accumulator and finishes with the residual epilogue.  The intermediate never
// touches HBM: per element the pair moves 2 B (operand in) + 4 B (residual in) + 4 B + 2 B (out) instead of 16 B, and
// the load -> MMA -> epilogue latency chain of a CTA is paid once for two convolutions.
//
//   y row r   (0 <= r < M1)           <-> time  tile*Mout - pad2 + r          Mout = M1 - 2*pad2 output rows per tile
//   A1 row q  (operand of conv1)       <-> time  tile*Mout - pad2 - pad1 + q   tap j of conv1 reads rows r + j*dil1
//   A2 row q  (operand of conv2)       =   y row q                             tap j of conv2 reads rows m + j
//   out row m (0 <= m < Mout)          <-> time  tile*Mout + m
// Both operand buffers use the un-swizzled K-major interleaved UMMA layout, so every tap is a row-offset descriptor
// (see conv_umma.cu).  Warp roles as in k_conv_umma: 0-3 A producers then both epilogues, 4 TMA weight ring, 5 MMA issuer.
#include "conv_umma.cuh"
#include "umma_ptx.cuh"

#include <cuda.h>
#include <stdlib.h>
#include <algorithm>

namespace b2 {

static constexpr int kPairThreads = 192;

struct PairParams {
    const __nv_bfloat16 *in;
    const float *residual, *bias1, *bias2, *acc_src;
    float *out32;
    __nv_bfloat16 *outb;
    float outb_slope, mid_slope, div;
    int W, T, C, taps, dil1, pad1, pad2;
    int mt, M1, Mout, R1, R2, KB, nkb, stages, tiles_per_win;
    unsigned long long m_tpw;
};

template <int N>
__global__ void __launch_bounds__(kPairThreads) k_resblock_pair(const __grid_constant__ CUtensorMap tmap_w1, const __grid_constant__ CUtensorMap tmap_w2,
                                                                const PairParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int KB = p.KB;
    const uint32_t b_stage_bytes = (uint32_t)N * KB * 2;
    const uint32_t a1_bytes = ((uint32_t)p.R1 * p.C * 2 + 15) & ~15u;
    const uint32_t a2_bytes = ((uint32_t)p.R2 * p.C * 2 + 15) & ~15u;
    uint8_t *sB = smem;
    uint8_t *sA1 = smem + (size_t)p.stages * b_stage_bytes;
    uint8_t *sA2 = sA1 + a1_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sA2 + a2_bytes);
    // bars: [0,4) b_full  [4,8) b_empty  [8,10) a1_full[kb]  10 acc1_full  11 a2_full  12 acc2_full
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t bar_b_full = bar0, bar_b_empty = bar0 + 32, bar_a1 = bar0 + 64, bar_acc1 = bar0 + 80, bar_a2 = bar0 + 88, bar_acc2 = bar0 + 96;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 13);

    const int w = fdiv(blockIdx.x, p.m_tpw);
    const int tile = blockIdx.x - w * p.tiles_per_win;
    const int tout0 = tile * p.Mout;                 // time of out row 0
    const int ty0 = tout0 - p.pad2;                  // time of y row 0
    const int ta0 = ty0 - p.pad1;                    // time of A1 row 0

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) __trap();
        for (int s = 0; s < 4; s++) { mbar_init(bar_b_full + 8 * s, 1); mbar_init(bar_b_empty + 8 * s, 1); }
        mbar_init(bar_a1, 128); mbar_init(bar_a1 + 8, 128);
        mbar_init(bar_acc1, 1); mbar_init(bar_a2, 128); mbar_init(bar_acc2, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const uint32_t tmem_cols = (uint32_t)(2 * p.mt * N);
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == 4 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w2) : "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_acc2 = tmem_base + (uint32_t)(p.mt * N);

    if (warp < 4) {
        const int tid = threadIdx.x;
        // =========================== A1 producer ===========================
        {
            const uint32_t sA_u32 = smem_u32(sA1);
            const int cshift = (KB == 64) ? 3 : 2;
            const int cpr = 1 << cshift;
            const int pieces = p.R1 * cpr;
            const __nv_bfloat16 *in_w = p.in + (size_t)w * p.T * p.C;
            for (int kb = 0; kb < p.nkb; kb++) {
                for (int q = tid; q < pieces; q += 128) {
                    const int r = q >> cshift, c = q & (cpr - 1);
                    const int t = ta0 + r;
                    const bool ok = (t >= 0) && (t < p.T);
                    const __nv_bfloat16 *src = ok ? in_w + (size_t)t * p.C + kb * KB + c * 8 : p.in;
                    cp_async16(sA_u32 + (uint32_t)(((kb * cpr + c) * p.R1 + r) * 16), src, ok ? 16u : 0u);
                }
                cp_async_arrive_noinc(bar_a1 + 8 * kb);
            }
        }
        // =========================== epilogue 1: acc1 -> bf16 operand of conv2 in shared memory ===========================
        mbar_wait(bar_acc1, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
            const uint32_t sA2_u32 = smem_u32(sA2);
            for (int sub = 0; sub < p.mt; sub++) {
                const int r = sub * 128 + warp * 32 + lane;
                const int t = ty0 + r;
                const bool inside = (t >= 0) && (t < p.T);
#pragma unroll 1
                for (int c0 = 0; c0 < N; c0 += 32) {
                    uint32_t acc[32];
                    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(sub * N + c0), acc);
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        uint32_t pk[4];
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            const int c = 8 * i + 2 * e;
                            const float2 b = __ldg(reinterpret_cast<const float2 *>(p.bias1 + c0 + c));
                            float v0 = lrelu_f(__uint_as_float(acc[c]) + b.x, p.mid_slope);
                            float v1 = lrelu_f(__uint_as_float(acc[c + 1]) + b.y, p.mid_slope);
                            if (!inside) { v0 = 0.0f; v1 = 0.0f; }
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
                            pk[e] = *reinterpret_cast<uint32_t *>(&h2);
                        }
                        const uint32_t dst = sA2_u32 + (uint32_t)((((c0 >> 3) + i) * p.R2 + r) * 16);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                    }
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> UMMA (async proxy) reads
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(bar_a2);
        // =========================== epilogue 2: residual, MRF sum, dual write ===========================
        mbar_wait(bar_acc2, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int sub = 0; sub < p.mt; sub++) {
            const int m = sub * 128 + warp * 32 + lane;
            const int t = tout0 + m;
            const bool ok = (m < p.Mout) && (t < p.T);
            const size_t row_off = ((size_t)w * p.T + t) * N;
#pragma unroll 1
            for (int c0 = 0; c0 < N; c0 += 32) {
                uint32_t acc[32];
                tmem_ld32(tmem_acc2 + ((uint32_t)(warp * 32) << 16) + (uint32_t)(sub * N + c0), acc);
                if (!ok) continue;
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 b = __ldg(reinterpret_cast<const float4 *>(p.bias2 + c0 + i));
                    v[i] = __uint_as_float(acc[i]) + b.x; v[i + 1] = __uint_as_float(acc[i + 1]) + b.y;
                    v[i + 2] = __uint_as_float(acc[i + 2]) + b.z; v[i + 3] = __uint_as_float(acc[i + 3]) + b.w;
                }
                {
                    const float4 *rp = reinterpret_cast<const float4 *>(p.residual + row_off + c0);
#pragma unroll
                    for (int i = 0; i < 8; i++) { const float4 r4 = rp[i]; v[4 * i] += r4.x; v[4 * i + 1] += r4.y; v[4 * i + 2] += r4.z; v[4 * i + 3] += r4.w; }
                }
                if (p.acc_src) {
                    const float4 *ap = reinterpret_cast<const float4 *>(p.acc_src + row_off + c0);
#pragma unroll
                    for (int i = 0; i < 8; i++) { const float4 r4 = ap[i]; v[4 * i] = r4.x + v[4 * i]; v[4 * i + 1] = r4.y + v[4 * i + 1]; v[4 * i + 2] = r4.z + v[4 * i + 2]; v[4 * i + 3] = r4.w + v[4 * i + 3]; }
                }
                if (p.div != 1.0f) {
#pragma unroll
                    for (int i = 0; i < 32; i++) v[i] = __fdiv_rn(v[i], p.div);
                }
                if (p.out32) {
                    float4 *op = reinterpret_cast<float4 *>(p.out32 + row_off + c0);
#pragma unroll
                    for (int i = 0; i < 8; i++) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
                if (p.outb) {
                    uint4 *op = reinterpret_cast<uint4 *>(p.outb + row_off + c0);
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        uint32_t pk[4];
#pragma unroll
                        for (int e = 0; e < 4; e++) {
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(lrelu_f(v[8 * i + 2 * e], p.outb_slope), lrelu_f(v[8 * i + 2 * e + 1], p.outb_slope));
                            pk[e] = *reinterpret_cast<uint32_t *>(&h2);
                        }
                        op[i] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
            }
        }
    } else if (warp == 4) {
        // =========================== weight ring: conv1's taps, then conv2's ===========================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int conv = 0; conv < 2; conv++) {
                const CUtensorMap *tm = conv ? &tmap_w2 : &tmap_w1;
                for (int kb = 0; kb < p.nkb; kb++)
                    for (int j = 0; j < p.taps; j++) {
                        mbar_wait(bar_b_empty + 8 * stage, phase ^ 1);
                        mbar_expect_tx(bar_b_full + 8 * stage, b_stage_bytes);
                        tma_load_3d(smem_u32(sB + (size_t)stage * b_stage_bytes), tm, bar_b_full + 8 * stage, kb * KB, 0, j);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
            }
        }
    } else {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
            const uint32_t b_layout = (KB == 64) ? 2u : 4u;
            const uint32_t b_sbo = 8u * (uint32_t)KB * 2;
            const int ksteps = KB / 16;
            int stage = 0; uint32_t phase = 0;
            for (int conv = 0; conv < 2; conv++) {
                const uint32_t sA_u32 = smem_u32(conv ? sA2 : sA1);
                const int R = conv ? p.R2 : p.R1;
                const int dil = conv ? 1 : p.dil1;
                const uint32_t a_lbo = (uint32_t)R * 16;
                const uint32_t tacc = conv ? tmem_acc2 : tmem_base;
                uint32_t accum = 0;
                if (conv) {
                    mbar_wait(bar_a2, 0);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                for (int kb = 0; kb < p.nkb; kb++) {
                    if (!conv) {
                        mbar_wait(bar_a1 + 8 * kb, 0);
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    }
                    for (int j = 0; j < p.taps; j++) {
                        mbar_wait(bar_b_full + 8 * stage, phase);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t b_base = smem_u32(sB + (size_t)stage * b_stage_bytes);
                        for (int sub = 0; sub < p.mt; sub++)
                            for (int ks = 0; ks < ksteps; ks++) {
                                const uint32_t a_addr = sA_u32 + (uint32_t)((((kb * (KB / 8) + ks * 2) * R) + sub * 128 + j * dil) * 16);
                                umma_f16(tacc + (uint32_t)(sub * N), smem_desc(a_addr, a_lbo, 128u, 0u), smem_desc(b_base + ks * 32, 0u, b_sbo, b_layout),
                                         idesc, (accum | (uint32_t)ks) ? 1u : 0u);
                            }
                        accum = 1;
                        umma_commit(bar_b_empty + 8 * stage);
                        if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    }
                }
                umma_commit(conv ? bar_acc2 : bar_acc1);
            }
        }
        __syncwarp();
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

static bool g_pair_attr[64][3] = {};

template <int N>
static int launch_pair_n(const CUtensorMap &t1, const CUtensorMap &t2, const PairParams &p, unsigned grid, size_t smem, cudaStream_t st, int slot) {
    int dev = 0;
    B2_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 64 && !g_pair_attr[dev][slot]) {
        B2_CUDA_OK(cudaFuncSetAttribute(k_resblock_pair<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        g_pair_attr[dev][slot] = true;
    }
    k_resblock_pair<N><<<grid, kPairThreads, smem, st>>>(t1, t2, p);
    B2_LAUNCH_OK("k_resblock_pair");
    return 0;
}

bool pair_supported(const Layer &l1, const Layer &l2, int T) {
    // Measured on B200 (profiles/README.md): bit-for-bit the same numerics as the two-launch path and 25 % less HBM traffic,
    // but not faster -- the thin layers are not HBM-bound and the fused CTA's larger footprint lowers occupancy.  Opt-in.
    static const bool on = getenv("B2_PAIR_FUSION") != nullptr;
    if (!on) return false;
    return l1.tmap && l2.tmap && l1.Cin == l1.Cout && l2.Cin == l2.Cout && l1.Cin == l2.Cin && l1.taps == l2.taps && l2.dil == 1 &&
           (l1.Cin == 32 || l1.Cin == 64 || l1.Cin == 128) && T >= 128;
}

int launch_resblock_pair(const PairArgs &a, cudaStream_t st) {
    const Layer &l1 = *a.conv1, &l2 = *a.conv2;
    if (!pair_supported(l1, l2, a.T)) return set_error("resblock_pair: unsupported layer pair");
    if (!a.in || !a.residual) return set_error("resblock_pair: null input");
    if (a.W <= 0) return 0;
    PairParams p;
    p.in = a.in; p.residual = a.residual; p.bias1 = l1.bias; p.bias2 = l2.bias; p.acc_src = a.acc_src; p.out32 = a.out32; p.outb = a.outb;
    p.outb_slope = a.outb_slope; p.mid_slope = a.mid_slope; p.div = a.div;
    p.W = a.W; p.T = a.T; p.C = l1.Cin; p.taps = l1.taps; p.dil1 = l1.dil; p.pad1 = l1.pad; p.pad2 = l2.pad;
    static const int mt_env = getenv("B2_PAIR_MT") ? atoi(getenv("B2_PAIR_MT")) : 0;
    int mt = (p.C == 32) ? 2 : 1;
    if (mt_env > 0 && p.C == 32) mt = mt_env;
    while (mt > 1 && 128 * mt > ((a.T + 127) / 128) * 128) mt >>= 1;
    p.mt = mt; p.M1 = 128 * mt; p.Mout = p.M1 - 2 * p.pad2;
    p.R1 = (p.M1 + 2 * p.pad1) | 1;
    p.R2 = (p.M1 + 2 * p.pad2) | 1;
    p.KB = p.C >= 64 ? 64 : 32;
    p.nkb = p.C / p.KB;
    p.tiles_per_win = cdiv(a.T, p.Mout);
    p.m_tpw = ((1ull << 40) + (unsigned long long)p.tiles_per_win - 1) / (unsigned long long)p.tiles_per_win;
    const size_t a1 = ((size_t)p.R1 * p.C * 2 + 15) & ~(size_t)15, a2 = ((size_t)p.R2 * p.C * 2 + 15) & ~(size_t)15;
    const size_t b_stage = (size_t)p.C * p.KB * 2;
    p.stages = 4;
    const size_t smem = p.stages * b_stage + a1 + a2 + 16 * 8;
    if (smem > 227 * 1024) return set_error("resblock_pair: tile needs %zu bytes of shared memory", smem);
    const unsigned grid = (unsigned)((long long)a.W * p.tiles_per_win);
    const CUtensorMap &t1 = *reinterpret_cast<const CUtensorMap *>(l1.tmap), &t2 = *reinterpret_cast<const CUtensorMap *>(l2.tmap);
    switch (p.C) {
        case 32: return launch_pair_n<32>(t1, t2, p, grid, smem, st, 0);
        case 64: return launch_pair_n<64>(t1, t2, p, grid, smem, st, 1);
        default: return launch_pair_n<128>(t1, t2, p, grid, smem, st, 2);
    }
}

}  // namespace b2
