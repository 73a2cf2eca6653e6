// Developer tool: the int8 tensor-pipe ceiling of one B200 for the tile shapes the batched-query kernel can use.
// Operands sit in shared memory (filled once, 128-byte swizzle layout, contents irrelevant), one thread per CTA
// issues tcgen05.mma.kind::i8 back to back into a ring of TMEM accumulators and nothing reads them: what is left is
// the tensor pipe + its shared-memory operand fetch.  bench.py reads the number this prints (profiles/) as the
// measured INT8 peak of the batched path's roofline.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/umma_i8_peak tools/umma_i8_peak.cu
//   tools/bin/umma_i8_peak [iters]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc_sw128(const void* smem_ptr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem_ptr) & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
template <int CG>
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (CG == 1) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
    }
}
template <int CG>
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    if constexpr (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// CG = cta_group (1 or 2), NT = UMMA N (total), KB = K bytes per tile (multiple of 128), SWAP = A is s8 and B u8.
// cta_group::2: UMMA M = 256 (128 rows of A per CTA), each CTA holds NT / 2 rows of B.
template <int CG, int NT, int KB, bool SWAP>
__global__ void __launch_bounds__(640) peak_kernel(int iters, unsigned long long* cycles) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int KC = KB / 128;
    constexpr int NB = NT / CG;                     // rows of B in this CTA's shared memory
    constexpr int STAGES = 512 / NT;         // 192 -> 2 stages
    uint8_t* sa = smem;                             // [KC][128][128]
    uint8_t* sb = smem + KC * 128 * 128;            // [KC][NB][128]
    __shared__ __align__(8) uint64_t bars[4];
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < KC * (128 + NB) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01020304u * (uint32_t)(i | 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        if constexpr (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        }
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    const bool leader = CG == 1 || cluster_ctarank() == 0;
    __shared__ volatile int stop_flag;
    if (threadIdx.x == 0) stop_flag = 0;
    __syncthreads();
    if (warp >= 4) {
        // contention warps: read the accumulators with tcgen05.ld.32x32b.x32 as fast as they can until the MMA loop ends
        unsigned acc = 0;
        const uint32_t lane_base = ((uint32_t)(warp & 3) * 32u) << 16;
        uint32_t c = (uint32_t)(warp >> 2) * 64u;
        while (!stop_flag) {
            uint32_t r[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                  "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                  "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                  "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(tmem + lane_base + (c & 511u)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc ^= r[0] ^ r[31];
            c += 32;
        }
        if (acc == 0x12345u) cycles[0] = acc;
    }
    if (threadIdx.x == 0 && leader) {
        const uint32_t idesc = (2u << 4) | (SWAP ? (1u << 7) : (1u << 10)) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const int s = it % STAGES;
            if (it >= STAGES) mbar_wait(&bars[s], ((it / STAGES) - 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int kc = 0; kc < KC; ++kc) {
                const uint64_t da = make_desc_sw128(sa + kc * 128 * 128), db = make_desc_sw128(sb + kc * NB * 128);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_i8<CG>(tmem + s * NT, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), idesc, (kc | ks) ? 1u : 0u);
            }
            umma_commit<CG>(&bars[s]);
        }
        for (int it = iters; it < iters + STAGES; ++it) {
            const int s = it % STAGES;
            if (it >= STAGES) mbar_wait(&bars[s], ((it / STAGES) - 1) & 1);
        }
        cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
        stop_flag = 1;
    }
    if constexpr (CG == 2) { if (!leader && threadIdx.x == 0) { /* the peer's readers stop when the pair leaves the loop */ } }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();
    if (warp == 0) {
        if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
}

template <int CG, int NT, int KB, bool SWAP>
static double run(const char* name, int iters, int sms, int extra_warps = 0) {
    constexpr int KC = KB / 128, NB = NT / CG;
    const size_t smem = (size_t)KC * (128 + NB) * 128 + 1024;
    auto kern = peak_kernel<CG, NT, KB, SWAP>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned long long* dcyc;
    CK(cudaMalloc(&dcyc, sizeof(unsigned long long) * sms));
    CK(cudaMemset(dcyc, 0, sizeof(unsigned long long) * sms));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)sms);
    cfg.blockDim = dim3(128 + 32 * extra_warps);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(e0));
        CK(cudaLaunchKernelEx(&cfg, kern, iters, dcyc));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep && ms < best) best = ms;
    }
    unsigned long long cyc = 0;
    CK(cudaMemcpy(&cyc, dcyc, sizeof(cyc), cudaMemcpyDeviceToHost));
    // every accumulator tile: (128 * CG) x NT x KB MACs per pair of CG CTAs, i.e. 128 x NT x KB per CTA
    const double ops = 2.0 * 128.0 * NT * KB * (double)iters * sms;
    const double tops = ops / (best * 1e-3) / 1e12;
    printf("{\"ldtm_warps\": %d, \"shape\": \"%s\", \"cta_group\": %d, \"umma_m\": %d, \"umma_n\": %d, \"k_bytes\": %d, \"a_s8_b_u8\": %s, \"ms\": %.4f, "
           "\"int8_tops\": %.1f, \"cycles_per_tile\": %.1f, \"macs_per_clk_per_sm\": %.0f}\n",
           extra_warps, name, CG, 128 * CG, NT, KB, SWAP ? "true" : "false", best, tops, (double)cyc / iters, 128.0 * NT * KB / ((double)cyc / iters));
    fflush(stdout);
    cudaFree(dcyc);
    return tops;
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 20000;
    CK(cudaSetDevice(0));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    sms &= ~1;
    run<1, 128, 256, false>("cg1 128x128 u8*s8", iters, sms);
    run<1, 128, 256, true>("cg1 128x128 s8*u8", iters, sms);
    run<1, 256, 256, true>("cg1 128x256 s8*u8", iters / 2, sms);
    run<1, 256, 128, true>("cg1 128x256 s8*u8 k128", iters, sms);
    run<1, 128, 512, true>("cg1 128x128 s8*u8 k512", iters / 2, sms);
    run<1, 192, 256, true>("cg1 128x192 s8*u8", iters / 2, sms);
    if (argc > 2) {
        run<1, 256, 256, true>("cg1 128x256 s8*u8 + ldtm", iters / 2, sms, 16);
        run<2, 256, 256, true>("cg2 256x256 s8*u8", iters / 2, sms);
        run<2, 128, 256, true>("cg2 256x128 s8*u8", iters, sms);
    }
    return 0;
}
