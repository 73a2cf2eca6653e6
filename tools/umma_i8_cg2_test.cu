// Developer test: one tcgen05.mma.cta_group::2.kind::i8 tile (M = 256, N = 256, K = 256, s8 x u8 -> s32) on a CTA pair,
// with exactly the protocol pieces the batched-query kernel uses in its cta_group::2 form:
//   * every CTA loads ITS half of A (128 rows) and ITS half of B (128 rows) with TMA (.cta_group::2) into its own shared
//     memory, completing the transaction count on the LEADER's mbarrier (address from mapa),
//   * the non-leader arms that barrier remotely (mbarrier.arrive.expect_tx.shared::cluster),
//   * the leader issues the MMAs and commits with multicast to both CTAs' barriers,
//   * each CTA reads its 128 TMEM lanes x 256 columns and arrives (remotely for the non-leader) on a leader barrier.
// Checked against the CPU.  Also prints which half of D / B each CTA holds, i.e. validates the operand split.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/umma_i8_cg2_test tools/umma_i8_cg2_test.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// arrive + expect_tx on a barrier given by its shared::cluster address (own or peer CTA)
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t bar_cluster_addr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_cluster_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
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
__device__ __forceinline__ void tma_load_2d_cg2(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(x), "r"(y) : "memory");
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
__device__ __forceinline__ void umma_i8_cg2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

constexpr int M = 256, N = 256, K = 256;       // per CTA: 128 rows of A, 128 rows of B

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
umma_cg2_test(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int* out, int* done_flag) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sa = smem;                         // 2 x [128][128]
    uint8_t* sb = smem + 2 * 128 * 128;         // 2 x [128][128]
    __shared__ __align__(8) uint64_t bar_full, bar_mma, bar_drained;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        mbar_init(&bar_full, 2);                // one arrive (with its expect_tx) per CTA of the pair
        mbar_init(&bar_mma, 1);
        mbar_init(&bar_drained, 2 * 4);         // four warps per CTA
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    const uint32_t full_leader = mapa_u32(smem_u32(&bar_full), 0);
    const uint32_t drained_leader = mapa_u32(smem_u32(&bar_drained), 0);
    if (threadIdx.x == 0) {
        // every CTA: its halves of A and B, completion on the leader's barrier
        mbar_expect_tx_cluster(full_leader, 4 * 128 * 128);
        tma_load_2d_cg2(sa, &map_a, full_leader, 0, (int)rank * 128);
        tma_load_2d_cg2(sa + 128 * 128, &map_a, full_leader, 128, (int)rank * 128);
        tma_load_2d_cg2(sb, &map_b, full_leader, 0, (int)rank * 128);
        tma_load_2d_cg2(sb + 128 * 128, &map_b, full_leader, 128, (int)rank * 128);
        if (rank == 0) {
            mbar_wait(&bar_full, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // c = s32, A = s8 (bit 7), B = u8, K-major both, N >> 3 at bit 17, M >> 4 at bit 24
            const uint32_t idesc = (2u << 4) | (1u << 7) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
            for (int ks = 0; ks < K / 32; ++ks) {
                const int sub = ks / 4, koff = (ks % 4) * 32;
                const uint64_t da = make_desc_sw128(sa + sub * 128 * 128) + (uint64_t)(koff >> 4);
                const uint64_t db = make_desc_sw128(sb + sub * 128 * 128) + (uint64_t)(koff >> 4);
                umma_i8_cg2(tmem, da, db, idesc, ks > 0 ? 1u : 0u);
            }
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(smem_u32(&bar_mma)), "h"((uint16_t)3) : "memory");
        }
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
        for (int i = 0; i < 32; ++i) out[((int)rank * 128 + warp * 32 + lane) * N + c0 + i] = (int)r[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(drained_leader);                 // remote for the non-leader
    if (rank == 0 && threadIdx.x == 0) { mbar_wait(&bar_drained, 0); *done_flag = 1; }
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u));
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
    srand(2);
    for (auto& v : ha) v = rand() & 255;
    for (auto& v : hb) v = rand() & 255;
    uint8_t *da, *db;
    int *dout, *dflag;
    CK(cudaMalloc(&da, ha.size())); CK(cudaMalloc(&db, hb.size())); CK(cudaMalloc(&dout, M * N * sizeof(int))); CK(cudaMalloc(&dflag, 4));
    CK(cudaMemset(dout, 0xFF, M * N * sizeof(int))); CK(cudaMemset(dflag, 0, 4));
    CK(cudaMemcpy(da, ha.data(), ha.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), hb.size(), cudaMemcpyHostToDevice));
    CUtensorMap ma = make_map(enc, da, M, 128), mb = make_map(enc, db, N, 128);
    const size_t smem = 4 * 128 * 128 + 1024;
    CK(cudaFuncSetAttribute(umma_cg2_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_cg2_test<<<2, 128, smem>>>(ma, mb, dout, dflag);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<int> ho(M * N);
    int flag = 0;
    CK(cudaMemcpy(ho.data(), dout, ho.size() * sizeof(int), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&flag, dflag, 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            int s = 0;
            for (int k = 0; k < K; ++k) s += (int)(int8_t)ha[m * K + k] * (int)hb[n * K + k];
            if (s != ho[m * N + n]) { if (bad < 8) printf("mismatch m=%d n=%d want %d got %d\n", m, n, s, ho[m * N + n]); ++bad; }
        }
    printf("umma_i8_cg2_test: %s (%ld mismatches of %d), drained barrier %s\n", bad ? "FAILED" : "OK", bad, M * N, flag ? "ok" : "NOT REACHED");
    return bad || !flag ? 1 : 0;
}
