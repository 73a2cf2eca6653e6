// Developer tool: the MMA <-> epilogue hand-off of the batched kernel in isolation (no TMA, operands resident):
// one thread issues 8 MMAs (K = 256) per 128 x 256 accumulator into a ring of two, commits; EPI warps wait, read the
// accumulator with tcgen05.ld (x64 twice per warp), arrive; the MMA thread waits for the arrivals before reusing it.
// Prints cycles per accumulator stage (1024 = tensor-pipe bound).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/umma_pipe_test tools/umma_pipe_test.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWL:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DN;\n\tbra WL;\n\tDN:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r; }
__device__ __forceinline__ void mbar_arrive_at(uint32_t a) { asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t make_desc_sw128(const void* smem_ptr) {
    uint64_t d = 0; d |= (uint64_t)((smem_u32(smem_ptr) & 0x3FFFF) >> 4); d |= (uint64_t)1 << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d;
}
template <int CG> __device__ __forceinline__ void umma_i8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if constexpr (CG == 1) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
    else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
template <int CG> __device__ __forceinline__ void umma_commit(uint64_t* bar) {
    if constexpr (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// EPI epilogue warps (multiple of 4), then one MMA warp.  NT = 256, K = 256 bytes.
template <int CG, int EPI, int KS>      // KS = K / 32 MMAs per stage: 8 = K 256, 2 = K 64
__global__ void __launch_bounds__(32 * (EPI + 1)) pipe_kernel(int iters, unsigned long long* cycles, unsigned* sink) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int NT = 256, NB = NT / CG;
    uint8_t* sa = smem; uint8_t* sb = smem + 2 * 128 * 128;
    __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 2 * (128 + NB) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01020304u * (uint32_t)(i | 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == EPI) {
        if constexpr (CG == 1) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
        else { asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;"); }
    }
    if (threadIdx.x == 0) { for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], CG * EPI); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    const uint32_t rank = CG == 1 ? 0u : cluster_ctarank();
    if (warp == EPI) {
        if (lane == 0 && rank == 0) {
            const uint32_t idesc = (2u << 4) | (1u << 7) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
            const long long t0 = clock64();
            for (int it = 0; it < iters; ++it) {
                const int s = it & 1;
                mbar_wait(&acc_empty[s], ((it >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kc = 0; kc < (KS + 3) / 4; ++kc) {
                    const uint64_t da = make_desc_sw128(sa + kc * 128 * 128), db = make_desc_sw128(sb + kc * NB * 128);
                    for (int ks = 0; ks < (KS < 4 ? KS : 4); ++ks) umma_i8<CG>(tmem + s * NT, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), idesc, (kc | ks) ? 1u : 0u);
                }
                umma_commit<CG>(&acc_full[s]);
            }
            cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
        }
    } else {
        const uint32_t quarter = warp & 3, slice = warp >> 2, cols = NT / (EPI / 4);
        const uint32_t acc_empty0 = CG == 1 ? smem_u32(&acc_empty[0]) : mapa_u32(smem_u32(&acc_empty[0]), 0);
        unsigned acc = 0;
        for (int it = 0; it < iters; ++it) {
            const int s = it & 1;
            mbar_wait(&acc_full[s], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (uint32_t c = 0; c < cols; c += 32) {
                uint32_t r[32];
                ld32(tmem + ((quarter * 32u) << 16) + s * NT + slice * cols + c, r);
                int m = (int)r[0];
#pragma unroll
                for (int i = 1; i < 32; ++i) m = max(m, (int)r[i]);
                acc ^= (unsigned)m;
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_at(acc_empty0 + s * 8u);
        }
        if (acc == 0x12345u) sink[0] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();
    if (warp == EPI) {
        if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
}
template <int CG, int EPI, int KS = 8> static void run(int iters) {
    const size_t smem = (size_t)2 * (128 + 256 / CG) * 128 + 1024;
    auto kern = pipe_kernel<CG, EPI, KS>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned long long* dcyc; unsigned* dsink;
    CK(cudaMalloc(&dcyc, 8 * 148)); CK(cudaMalloc(&dsink, 4)); CK(cudaMemset(dcyc, 0, 8 * 148));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(32 * (EPI + 1)); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) { CK(cudaLaunchKernelEx(&cfg, kern, iters, dcyc, dsink)); CK(cudaDeviceSynchronize()); }
    unsigned long long cyc = 0; CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
    printf("{\"cta_group\": %d, \"epilogue_warps\": %d, \"K\": %d, \"cycles_per_stage\": %.1f}\n", CG, EPI, KS * 32, (double)cyc / iters); fflush(stdout);
    cudaFree(dcyc); cudaFree(dsink);
}
int main() {
    CK(cudaSetDevice(0));
    run<1, 4>(20000); run<1, 8>(20000); run<1, 16>(20000);
    run<2, 4>(20000); run<2, 8>(20000); run<2, 16>(20000);
    run<1, 16, 2>(20000); run<2, 4, 2>(20000); run<2, 8, 2>(20000); run<2, 16, 2>(20000);     // K = 64: the hand-off chain itself
    run<2, 16, 1>(20000); run<2, 16, 4>(20000);
    return 0;
}
