// Developer test: one tcgen05.mma.kind::i8 tile (M=128, N=256, K=256, u8 x u8 -> s32) fed by TMA with
// 128-byte swizzle, checked against the CPU.  Validates the tensor-map / shared-memory descriptor /
// instruction-descriptor encodings used by the batched-query kernel before they are buried in a pipeline.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/umma_i8_test tools/umma_i8_test.cu && /tmp/umma_i8_test
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint64_t make_desc_sw128(const void* smem_ptr) {
    // K-major, 128-byte swizzle: start address >> 4, LBO (unused) = 1, SBO = 1024 B (8 rows x 128 B), version 1, layout 2
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem_ptr) & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}

constexpr int M = 128, N = 256, K = 256;

__global__ void __launch_bounds__(128) umma_test(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int* out) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzled tiles need 1024-byte alignment
    uint8_t* sa = smem;                         // 2 x [128][128]
    uint8_t* sb = smem + 2 * M * 128;           // 2 x [256][128]
    __shared__ uint64_t bar_full, bar_mma;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        mbar_init(&bar_full, 1);
        mbar_init(&bar_mma, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar_full, 2 * M * 128 + 2 * N * 128);
        tma_load_2d(sa, &map_a, &bar_full, 0, 0);
        tma_load_2d(sa + M * 128, &map_a, &bar_full, 128, 0);
        tma_load_2d(sb, &map_b, &bar_full, 0, 0);
        tma_load_2d(sb + N * 128, &map_b, &bar_full, 128, 0);
        mbar_wait(&bar_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // instruction descriptor: c=S32 (2<<4), a,b = u8 (0), K-major both, N>>3 at bit 17, M>>4 at bit 24
#ifdef SWAP_SIGN      // A = s8 (1 at bit 7), B = u8: the operand roles of the transposed batched kernel (queries are A)
        const uint32_t idesc = (2u << 4) | (1u << 7) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
#else
        const uint32_t idesc = (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
#endif
        for (int ks = 0; ks < K / 32; ++ks) {
            const int sub = ks / 4, koff = (ks % 4) * 32;
            const uint64_t da = make_desc_sw128(sa + sub * M * 128) + (uint64_t)(koff >> 4);
            const uint64_t db = make_desc_sw128(sb + sub * N * 128) + (uint64_t)(koff >> 4);
            umma_i8(tmem, da, db, idesc, ks > 0 ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
    }
    __syncwarp();
    mbar_wait(&bar_mma, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * N + c0 + i] = (int)r[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeTiled enc, void* base, uint64_t rows, uint32_t box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)K, rows};
    cuuint64_t strides[1] = {(cuuint64_t)K};
    cuuint32_t box[2] = {128, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    return m;
}

int main() {
    CK(cudaSetDevice(0));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    EncodeTiled enc = (EncodeTiled)fn;
    std::vector<uint8_t> ha(M * K), hb(N * K);
    srand(1);
    for (auto& v : ha) v = rand() & 255;
    for (auto& v : hb) v = rand() & 255;
    uint8_t *da, *db;
    int* dout;
    CK(cudaMalloc(&da, ha.size())); CK(cudaMalloc(&db, hb.size())); CK(cudaMalloc(&dout, M * N * sizeof(int)));
    CK(cudaMemcpy(da, ha.data(), ha.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), hb.size(), cudaMemcpyHostToDevice));
    CUtensorMap ma = make_map(enc, da, M, 128), mb = make_map(enc, db, N, 256);
    const size_t smem = 2 * M * 128 + 2 * N * 128 + 1024;
    CK(cudaFuncSetAttribute(umma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_test<<<1, 128, smem>>>(ma, mb, dout);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<int> ho(M * N);
    CK(cudaMemcpy(ho.data(), dout, ho.size() * sizeof(int), cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            int s = 0;
#ifdef SWAP_SIGN
            for (int k = 0; k < K; ++k) s += (int)(int8_t)ha[m * K + k] * (int)hb[n * K + k];
#else
            for (int k = 0; k < K; ++k) s += (int)ha[m * K + k] * (int)hb[n * K + k];
#endif
            if (s != ho[m * N + n]) { if (bad < 5) printf("mismatch m=%d n=%d want %d got %d\n", m, n, s, ho[m * N + n]); ++bad; }
        }
    printf("umma_i8_test: %s (%ld mismatches of %d)\n", bad ? "FAILED" : "OK", bad, M * N);
    return bad ? 1 : 0;
}
