// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N and of the shared-memory layout of A.
// Answers one question for the conv kernels: is the tap-shift-friendly un-swizzled K-major layout of the activations
// (rows 16 bytes apart) slower to feed to the tensor core than the canonical 128-byte-swizzled one?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/mma_rate tools/mma_rate.cu && gpurun_out/mma_rate
#include "../infernos_b200/csrc/umma_ptx.cuh"
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

using namespace b2;

struct Cfg {
    int N;          // MMA N
    int a_mode;     // 0: un-swizzled interleaved [chunk][row][8]  (LBO = R*16, SBO = 128)
                    // 1: 128B swizzle K-major (rows of 64 bf16 = 128 B, SBO = 1024)
                    // 2: un-swizzled, canonical core-matrix order (LBO = 128, SBO = 256): the two K chunks of a row group adjacent
    int b_mode;     // 2: 128B swizzle, 4: 64B swizzle, 0: un-swizzled interleaved (like A mode 0)
    int iters;      // MMAs per timed run
    int distinct;   // how many different A start rows are cycled through (1 = same operand every time)
    int a_step;     // rows between consecutive A starts (tap shift)
    int commit_every; // 0: one commit at the end; n: a tcgen05.commit (to a second barrier) after every n MMAs
    int smem_kb;      // dynamic shared memory of the launch (100: two CTAs share an SM)
    int alt_k;        // 1: alternate between the two K16 halves of a 32-channel operand (A + 2 chunks, B + 32 bytes), as the C = 32 kernels do
};

__global__ void __launch_bounds__(128) k_rate(Cfg c, unsigned long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, bar2;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    // zero operands (timing does not depend on the values)
    for (int i = threadIdx.x; i < c.smem_kb * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (warp == 1) {
        const bool leader = elect_one();
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t sA = smem_u32(smem), sB = smem_u32(smem) + (uint32_t)(c.smem_kb - 32) * 1024;
        const int R = 701;
        uint64_t adesc0;
        if (c.a_mode == 0) adesc0 = smem_desc(sA, R * 16, 128u, 0u);
        else if (c.a_mode == 1) adesc0 = smem_desc(sA, 0u, 1024u, 2u);
        else adesc0 = smem_desc(sA, 128u, 256u, 0u);
        uint64_t bdesc0;
        if (c.b_mode == 2) bdesc0 = smem_desc(sB, 0u, 1024u, 2u);
        else if (c.b_mode == 4) bdesc0 = smem_desc(sB, 0u, 512u, 4u);
        else bdesc0 = smem_desc(sB, 300 * 16, 128u, 0u);
        const uint32_t row_units = (c.a_mode == 1) ? 8u : 1u;      // 16-byte units per row
        for (int rep = 0; rep < 3; rep++) {
            const unsigned long long t0 = clock64();
            int d = 0;
            for (int i = 0; i < c.iters; i++) {
                const uint64_t adesc = adesc0 + (uint64_t)((uint32_t)(d * c.a_step) * row_units);
                const bool hi = c.alt_k && (i & 1);
                if (leader) umma_f16(tmem, adesc + (hi ? (uint64_t)(2 * R) : 0ull), bdesc0 + (hi ? 2ull : 0ull), idesc, 1u);
                if (++d == c.distinct) d = 0;
                if (c.commit_every && ((i + 1) % c.commit_every) == 0 && leader) umma_commit(smem_u32(&bar2));
            }
            if (leader) umma_commit(smem_u32(&bar));
            __syncwarp();
            mbar_wait(smem_u32(&bar), (uint32_t)(rep & 1));
            const unsigned long long t1 = clock64();
            if (leader && rep == 2) out[blockIdx.x] = t1 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

int main() {
    unsigned long long *d_out;
    cudaMalloc(&d_out, 1024 * 8);
    cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int Ns[4] = {32, 64, 128, 256};
    printf("# cycles per tcgen05.mma M=128 K=16 bf16 (one CTA per SM, 148 CTAs; min / median over CTAs)\n");
    printf("# N a_mode b_mode distinct a_step : min med   [tensor floor = N/2]\n");
    for (int a_mode = 0; a_mode < 3; a_mode++)
        for (int ni = 0; ni < 4; ni++)
            for (int variant = 0; variant < 3; variant++) {
                Cfg c;
                c.N = Ns[ni]; c.a_mode = a_mode; c.iters = 2048;
                c.b_mode = (variant == 2) ? 0 : ((c.N == 32) ? 4 : 2);
                c.commit_every = 0; c.smem_kb = 160; c.alt_k = 0;
                c.distinct = (variant == 0) ? 1 : 11; c.a_step = (variant == 0) ? 0 : ((a_mode == 1) ? 8 : 5);
                if (a_mode == 1 && variant == 2) continue;
                for (int grid : {1, 148}) {
                    k_rate<<<grid, 128, 160 * 1024>>>(c, d_out);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                    std::vector<unsigned long long> h(grid);
                    cudaMemcpy(h.data(), d_out, grid * 8, cudaMemcpyDeviceToHost);
                    std::sort(h.begin(), h.end());
                    printf("N=%3d a_mode=%d b_mode=%d distinct=%2d a_step=%d grid=%3d : %6.1f %6.1f\n", c.N, c.a_mode, c.b_mode, c.distinct, c.a_step, grid,
                           (double)h[0] / c.iters, (double)h[grid / 2] / c.iters);
                }
            }
    printf("# effect of tcgen05.commit frequency (a_mode 0, 11 distinct A starts)\n");
    for (int ni = 0; ni < 4; ni++)
        for (int ce : {0, 32, 8, 4, 2, 1}) {
            Cfg c;
            c.N = Ns[ni]; c.a_mode = 0; c.iters = 2048; c.b_mode = (c.N == 32) ? 4 : 2; c.distinct = 11; c.a_step = 5; c.commit_every = ce; c.smem_kb = 160; c.alt_k = 0;
            k_rate<<<148, 128, 160 * 1024>>>(c, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
            std::vector<unsigned long long> h(148);
            cudaMemcpy(h.data(), d_out, 148 * 8, cudaMemcpyDeviceToHost);
            std::sort(h.begin(), h.end());
            printf("N=%3d commit_every=%2d : %6.1f cycles/MMA\n", c.N, ce, (double)h[74] / c.iters);
        }
    printf("# 64B-swizzled B (rows of 32 bf16) at every N, alternating K16 halves; one and two CTAs per SM (round 2, conv_resblock_t.cu's operands)\n");
    for (int N : {32, 64, 96, 128})
        for (int bm : {4, 2})
            for (int two : {0, 1}) {
                Cfg c;
                c.N = N; c.a_mode = 0; c.iters = 2048; c.b_mode = bm; c.distinct = 11; c.a_step = 5; c.commit_every = 0; c.smem_kb = two ? 100 : 160; c.alt_k = 1;
                const int grid = two ? 296 : 148;
                k_rate<<<grid, 128, c.smem_kb * 1024>>>(c, d_out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
                std::vector<unsigned long long> h(grid);
                cudaMemcpy(h.data(), d_out, grid * 8, cudaMemcpyDeviceToHost);
                std::sort(h.begin(), h.end());
                printf("N=%3d b_mode=%d ctas_per_sm=%d : %6.1f cycles/MMA per CTA\n", N, bm, two + 1, (double)h[grid / 2] / c.iters);
            }
    return 0;
}
