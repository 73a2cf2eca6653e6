// Developer tool: TMEM -> register bandwidth of tcgen05.ld on one B200 (what bounds the batched kernel's epilogue:
// every score has to leave TMEM through this path).  No MMA: each warp loads its 32-lane quarter over and over.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/tmem_ld_bw tools/tmem_ld_bw.cu && tools/bin/tmem_ld_bw
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X> struct Ld;
template <> struct Ld<32> {
    static __device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&r)[32]) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
    }
};
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// MODE 0: x32, wait after every load.  MODE 1: x32, two loads in flight.  MODE 2: x16, wait after every load.
// MODE 3: x32 + the 16-instruction max tree and a compare (the batched epilogue's fast path).
template <int MODE>
__global__ void __launch_bounds__(512) ld_kernel(int iters, unsigned long long* cycles, unsigned* sink) {
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    const uint32_t lane_base = ((uint32_t)(warp & 3) * 32u) << 16;
    const uint32_t nslice = blockDim.x / 128;                 // warps per lane quarter
    const uint32_t slice = (uint32_t)warp >> 2;
    const uint32_t cols = 512u / nslice;                      // columns of this warp
    unsigned acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if constexpr (MODE == 0 || MODE == 3) {
            for (uint32_t c = 0; c < cols; c += 32) {
                uint32_t r[32];
                Ld<32>::ld(tmem + lane_base + slice * cols + c, r);
                ld_wait();
                if constexpr (MODE == 3) {
                    int t[11];
#pragma unroll
                    for (int i = 0; i < 10; ++i) t[i] = max(max((int)r[3 * i], (int)r[3 * i + 1]), (int)r[3 * i + 2]);
                    t[10] = max((int)r[30], (int)r[31]);
                    int u0 = max(max(t[0], t[1]), t[2]), u1 = max(max(t[3], t[4]), t[5]), u2 = max(max(t[6], t[7]), t[8]), u3 = max(t[9], t[10]);
                    int mx = max(max(max(u0, u1), u2), u3);
                    if (__any_sync(0xFFFFFFFFu, mx >= 0x7FFFFFF0)) acc ^= (unsigned)mx;
                } else {
                    acc ^= r[0] ^ r[31];
                }
            }
        } else if constexpr (MODE == 1) {
            for (uint32_t c = 0; c < cols; c += 64) {
                uint32_t r0[32], r1[32];
                Ld<32>::ld(tmem + lane_base + slice * cols + c, r0);
                Ld<32>::ld(tmem + lane_base + slice * cols + c + 32, r1);
                ld_wait();
                acc ^= r0[0] ^ r0[31] ^ r1[0] ^ r1[31];
            }
        } else {
            for (uint32_t c = 0; c < cols; c += 16) {
                uint32_t r[16];
                ld16(tmem + lane_base + slice * cols + c, r);
                ld_wait();
                acc ^= r[0] ^ r[15];
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345678u) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

template <int MODE>
static void run(const char* name, int warps, int iters) {
    unsigned long long* dcyc; unsigned* dsink;
    CK(cudaMalloc(&dcyc, 8 * 148)); CK(cudaMalloc(&dsink, 4));
    ld_kernel<MODE><<<148, warps * 32>>>(iters, dcyc, dsink);
    CK(cudaDeviceSynchronize());
    ld_kernel<MODE><<<148, warps * 32>>>(iters, dcyc, dsink);
    CK(cudaDeviceSynchronize());
    unsigned long long cyc = 0;
    CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
    const double bytes = 128.0 * 512 * 4 * iters;             // the whole TMEM once per iteration
    printf("{\"mode\": \"%s\", \"warps\": %d, \"bytes_per_clk_per_sm\": %.1f, \"cycles_per_128x256_accumulator\": %.0f}\n", name, warps,
           bytes / (double)cyc, 131072.0 / (bytes / (double)cyc));
    fflush(stdout);
    cudaFree(dcyc); cudaFree(dsink);
}

int main() {
    CK(cudaSetDevice(0));
    for (int warps : {4, 8, 16}) {
        run<0>("x32 wait each", warps, 2000);
        run<1>("x32 two in flight", warps, 2000);
        run<2>("x16 wait each", warps, 2000);
        run<3>("x32 + max tree", warps, 2000);
    }
    return 0;
}
