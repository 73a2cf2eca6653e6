// batch.cuh -- kernel B: the batched-query path.  A batch of queries against the corpus is a dense
// u8 x s8 -> s32 contraction  S[r, q] = sum_i r_i (q_i - 128),  run on the 5th-generation tensor cores
// (tcgen05.mma.kind::i8, accumulators in TMEM), with the top-k selection fused into the epilogue so the
// N x Q score matrix (10^10 entries for 10M x 1024) never exists in memory.
//
//   operands       :  A = corpus bytes r (u8), B = query bytes centred on 128, q' = q - 128 = q ^ 0x80 (s8);  S = sum r q'
//   exact integers:  dot_i = sum c(q)c(r) = 4 S + (2 sum r - 255 d) + (-510 sum q')   (c(v) = 2v - 255)
//                    the per-row term varies little across rows (it is 2 sum r, not 510 sum r), so the epilogue
//                    can pre-test the RAW accumulator against a per-column integer bound: one compare per score
//   ranking key   :  kappa' = fl(fl(dot_i) * inv_norm_r)      (the per-query factor 1/|c(q)| is applied later)
//
// One CTA = one query group (QG <= 512 queries resident in shared memory, loaded once by TMA) x a strided set of
// 128-row corpus tiles.  Warp 0 streams corpus tiles with TMA (128-byte swizzle, 16 KB K-chunks, 4-stage
// mbarrier ring); one thread of warp 1 issues the MMAs (M = 128 rows, N = 128 queries, K = 32 bytes per
// instruction) into a ring of four TMEM accumulators; warps 2..17 are the epilogue: each owns 32 TMEM lanes (rows)
// and 32 columns of every accumulator, reads them with one tcgen05.ld.32x32b.x32, hands the accumulator straight
// back to the MMA warp, and tests every score with ONE integer compare against a per-(warp, column) bound that is
// rebuilt per tile from the query's current threshold and the norm / row-term range of the warp's 32 rows.  Only
// columns in which some lane passes take the exact test (int -> float, one multiply, one compare); the few that
// beat the threshold are staged in shared memory and pushed 32 at a time into the query's candidate buffer.
// Round 0 (2048 rows) has no threshold yet: a flood variant writes every score to slot = row, no atomics.  Later
// rounds cover geometrically growing row ranges; between rounds batch_tighten_kernel cuts the buffers back to
// `keep`, inside a round the thresholds tighten from per-query histograms of the accepted keys.  The final
// candidates go through the same bit-exact re-rank and certificate as the single-query path (finalize_kernel<true>).
#pragma once
#include <cuda.h>
#include <cstdio>
#include "rerank.cuh"

namespace pbx {

constexpr int kBatchEpiWarps = 16;           // four per TMEM lane quarter, each takes a quarter of an accumulator's columns
constexpr int kBatchThreads = 64 + 32 * kBatchEpiWarps;   // warp 0 TMA, warp 1 MMA, warps 2.. epilogue
constexpr int kBatchMaxStages = 8;           // corpus K-chunk ring: up to 8 x 16 KB (the host sizes it to the shared memory left)
constexpr int kBatchTileRows = 128;          // UMMA M
constexpr int kBatchAccStages = 4;           // TMEM: 4 accumulators of 128 rows x 128 queries (512 columns), a ring between MMA and epilogue
constexpr uint32_t kBatchAccCols = 128;      // UMMA N
constexpr uint32_t kBatchCap = 4096;         // candidate buffer entries per query (small k)
constexpr uint32_t kBatchCapLarge = 16384;   // ... for keep > 512 (k = 1000 class)
constexpr uint32_t kBatchQueryBytes = 128 * 1024;   // resident queries per CTA
constexpr uint32_t kBatchHistBins = 256;            // per-query histogram of accepted keys over kappa in [-1, 1]

// ---- PTX helpers -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the waiting thread sleeps in hardware between polls instead of spinning on the
// issue port.  (Measured against a tight spin and against try_wait + nanosleep(32): within 2.5 % of each other.)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "PBX_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra PBX_DONE;\n\t"
        "bra PBX_WAIT;\n\t"
        "PBX_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
// K-major operand, 128-byte swizzle: start address >> 4, LBO unused (1), SBO = 1024 B (8 rows x 128 B), version 1
__device__ __forceinline__ uint64_t umma_desc_sw128(const void* smem_ptr) {
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
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
}

// ---- per-batch query preparation: one CTA per (padded) query ---------------------------------------------
struct BatchPrepParams {
    const uint8_t* queries;     // [nq][dim]
    uint32_t nq, dim, pitch;
    uint8_t* qpad;              // [nq_pad][pitch] raw bytes, zero padded (rows beyond nq are zero): the MMA's B operand
    int16_t* q16;               // [nq][pitch] centred (re-rank)
    uint8_t* qbytes;            // [nq][pitch]
    QueryHeader* qh;            // [nq]
    int* colterm;               // [nq_pad]  -510 * sum (q_i - 128)
    float* thr;                 // [nq_pad]  -inf for real queries, +inf for padding
    uint32_t* cand_cnt;         // [nq_pad]
    uint32_t* overflow;         // [nq_pad]
    uint32_t* bhist;            // [nq_pad][kBatchHistBins], zeroed here
    float* inv_q;               // [nq_pad]
    uint32_t flood_rows;        // rows of round 0: every real query starts with exactly these candidates (slot = row)
};

__global__ void batch_prep_kernel(const BatchPrepParams p) {
    const uint32_t q = blockIdx.x;
    const bool real = q < p.nq;
    int s = 0, n2 = 0, raw = 0;
    for (uint32_t i = threadIdx.x; i < p.pitch; i += blockDim.x) {
        uint32_t v = 0;
        int c = 0;
        if (real && i < p.dim) { v = p.queries[(size_t)q * p.dim + i]; c = centre(v); }
        p.qpad[(size_t)q * p.pitch + i] = (real && i < p.dim) ? (uint8_t)(v ^ 0x80u) : (uint8_t)0;   // q - 128 as s8
        if (real) { p.q16[(size_t)q * p.pitch + i] = (int16_t)c; p.qbytes[(size_t)q * p.pitch + i] = (uint8_t)v; }
        s += c; n2 += c * c; raw += (int)v;
    }
    __shared__ int ss[32], sn[32], sr[32];
    for (int off = 16; off; off >>= 1) {
        s += __shfl_xor_sync(~0u, s, off); n2 += __shfl_xor_sync(~0u, n2, off); raw += __shfl_xor_sync(~0u, raw, off);
    }
    if ((threadIdx.x & 31) == 0) { ss[threadIdx.x >> 5] = s; sn[threadIdx.x >> 5] = n2; sr[threadIdx.x >> 5] = raw; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int S = 0, N = 0, R = 0;
        for (uint32_t w = 0; w < (blockDim.x + 31) / 32; ++w) { S += ss[w]; N += sn[w]; R += sr[w]; }
        if (real) {
            QueryHeader h;
            h.sum_cq = S; h.norm2_q = N; h.inv_q = (float)(1.0 / sqrt((double)N)); h.sa = 0.0f;
            p.qh[q] = h;
        }
        p.colterm[q] = -510 * (R - 128 * (int)p.dim);          // -510 * sum (q_i - 128)   (R = 0 for padding queries)
        p.thr[q] = real ? -__int_as_float(0x7f800000) : __int_as_float(0x7f800000);
        p.cand_cnt[q] = real ? p.flood_rows : 0u;
        p.overflow[q] = 0;
        p.inv_q[q] = real ? (float)(1.0 / sqrt((double)N)) : 0.0f;
    }
    for (uint32_t i = threadIdx.x; i < kBatchHistBins; i += blockDim.x) p.bhist[(size_t)q * kBatchHistBins + i] = 0;
}

// ---- the contraction + selection kernel ---------------------------------------------------------------------
struct BatchMmaParams {
    CUtensorMap map_rows;       // corpus [capacity][pitch] u8, box {128 B, 128 rows}, 128-byte swizzle
    CUtensorMap map_q;          // padded queries [nq_pad][pitch] u8, box {128 B, min(256, QG) rows}
    const float* inv_norm;
    const int* row_sum;
    const int* colterm;         // [nq_pad]
    const float* thr;           // [nq_pad] thresholds on kappa' for this round
    u64* cand;                  // [nq_pad][cap]
    uint32_t* cand_cnt;         // [nq_pad]
    uint32_t* overflow;         // [nq_pad]
    uint32_t* bhist;            // [nq_pad][kBatchHistBins] accepted keys per query and kappa bin (in-round tightening)
    const float* inv_q;         // [nq_pad] 1 / |c(q)|  (0 for padding queries)
    float* thr_live;            // == thr, written: thresholds tightened while the round runs
    uint32_t keep;
    uint32_t cap;               // candidate buffer entries per query
    uint32_t n;                 // rows visible to this search
    uint32_t dim;
    uint32_t kc;                // K-chunks of 128 bytes per row (pitch / 128)
    uint32_t qg;                // queries per group (resident per CTA): 128, 256 or 512
    uint32_t groups;            // query groups; gridDim.x is a multiple of it
    uint32_t stages;            // corpus K-chunk ring depth (16 KB each)
    uint32_t tile_begin, tile_end;   // 128-row tiles of this round
};

template <bool FLOOD>
__global__ void __launch_bounds__(kBatchThreads, 1)
batch_mma_kernel(const __grid_constant__ BatchMmaParams p) {
    extern __shared__ __align__(16) uint8_t bsm_raw[];     // NOT declared 1024-aligned: the compiler would fold the fix-up below away
    // the 128-byte swizzle of TMA and UMMA is a function of the shared-memory address: tiles must sit on 1024-byte
    // boundaries, and dynamic shared memory only starts after the static variables (the host adds 1 KB of slack)
    uint8_t* bsm = bsm_raw + ((1024u - (smem_u32(bsm_raw) & 1023u)) & 1023u);
    const uint32_t QG = p.qg, KC = p.kc;
    const uint32_t NMMA = QG < 256 ? QG : 256;          // rows of one TMA box of the query operand
    const uint32_t NS = QG / kBatchAccCols;             // accumulator stages per tile (4, 2 or 1): 128 queries each
    uint8_t* sQ = bsm;                                  // [KC][QG][128]
    uint8_t* sA = bsm + (size_t)QG * KC * 128;          // [stages][128][128]
    const uint32_t STAGES = p.stages;
    int* s_colterm = reinterpret_cast<int*>(sA + (size_t)STAGES * kBatchTileRows * 128);
    float* s_thr = reinterpret_cast<float*>(s_colterm + QG);
    float* s_invq = s_thr + QG;
    float2* s_pre = reinterpret_cast<float2*>(s_invq + QG);   // per column {threshold clamped to +-1e30, (float)colterm}: inputs of the pre-test bound
    int* s_uw = reinterpret_cast<int*>(s_pre + QG);     // [epilogue warp][2 * 64] integer pre-test bounds, rebuilt per tile
    // per-epilogue-warp staging of accepted candidates: pushes to global memory go out 32 at a time, so the
    // ~1 us round trip of the slot atomic is paid once per 32 candidates instead of once per candidate
    __shared__ u64 st_key[kBatchEpiWarps][64];
    __shared__ uint32_t st_q[kBatchEpiWarps][64];
    __shared__ __align__(8) uint64_t q_full, a_full[kBatchMaxStages], a_empty[kBatchMaxStages], acc_full[kBatchAccStages], acc_empty[kBatchAccStages];
    __shared__ uint32_t tmem_base;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t g = blockIdx.x % p.groups;
    const uint32_t ci = blockIdx.x / p.groups, cstride = gridDim.x / p.groups;

    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        mbar_init(&q_full, 1);
        for (int i = 0; i < kBatchMaxStages; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < kBatchAccStages; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kBatchEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (uint32_t i = threadIdx.x; i < QG; i += blockDim.x) {
        s_colterm[i] = p.colterm[g * QG + i];
        s_thr[i] = p.thr[g * QG + i];
        s_invq[i] = p.inv_q[g * QG + i];
        s_pre[i] = make_float2(fminf(fmaxf(s_thr[i], -1.0e30f), 1.0e30f), (float)s_colterm[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
#ifdef PBX_DEBUG_BATCH
            if (blockIdx.x == 0) printf("bsm_raw %x bsm %x sQ %x sA %x q_full %x map_q %p map_rows %p QG %u KC %u NMMA %u\n", smem_u32(bsm_raw), smem_u32(bsm),
                                        smem_u32(sQ), smem_u32(sA), smem_u32(&q_full), (const void*)&p.map_q, (const void*)&p.map_rows, QG, KC, NMMA);
#endif
            mbar_expect_tx(&q_full, QG * KC * 128);
            for (uint32_t kc = 0; kc < KC; ++kc)
                for (uint32_t h = 0; h < QG; h += NMMA)
                    tma_load_2d(sQ + ((size_t)kc * QG + h) * 128, &p.map_q, &q_full, (int)(kc * 128), (int)(g * QG + h));
            uint32_t it = 0;
            for (uint32_t t = p.tile_begin + ci; t < p.tile_end; t += cstride) {
                for (uint32_t kc = 0; kc < KC; ++kc, ++it) {
                    const uint32_t st = it % STAGES, ph = (it / STAGES) & 1u;
                    mbar_wait(&a_empty[st], ph ^ 1u);
                    mbar_expect_tx(&a_full[st], kBatchTileRows * 128);
                    tma_load_2d(sA + (size_t)st * kBatchTileRows * 128, &p.map_rows, &a_full[st], (int)(kc * 128), (int)(t * kBatchTileRows));
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // instruction descriptor: D = s32 (2 << 4), A = u8 (0 at bit 7), B = s8 (1 at bit 10), both K-major,
            // N >> 3 at bit 17, M >> 4 at bit 24
            const uint32_t idesc = (2u << 4) | (1u << 10) | ((kBatchAccCols >> 3) << 17) | ((uint32_t)(kBatchTileRows >> 4) << 24);
            mbar_wait(&q_full, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t it = 0, acc_it = 0;
            for (uint32_t t = p.tile_begin + ci; t < p.tile_end; t += cstride) {
                const uint32_t it0 = it;
                for (uint32_t sg = 0; sg < NS; ++sg, ++acc_it) {
                    const uint32_t ab = acc_it % kBatchAccStages, par = (acc_it / kBatchAccStages) & 1u;
                    mbar_wait(&acc_empty[ab], par ^ 1u);                          // the epilogue has drained its previous use
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t d_tmem = tmem + ab * kBatchAccCols;
                    for (uint32_t kc = 0; kc < KC; ++kc) {
                        const uint32_t itk = it0 + kc;
                        const uint32_t st = itk % STAGES, ph = (itk / STAGES) & 1u;
                        if (sg == 0) {
                            mbar_wait(&a_full[st], ph);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        }
                        const uint64_t da = umma_desc_sw128(sA + (size_t)st * kBatchTileRows * 128);
                        const uint64_t db = umma_desc_sw128(sQ + ((size_t)kc * QG + (size_t)sg * kBatchAccCols) * 128);
#pragma unroll
                        for (uint32_t ks = 0; ks < 4; ++ks)
                            umma_i8(d_tmem, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), idesc, (kc | ks) ? 1u : 0u);
                        if (sg == NS - 1) umma_commit(&a_empty[st]);              // the stage is free once these MMAs retire
                    }
                    umma_commit(&acc_full[ab]);
                }
                it = it0 + KC;
            }
        }
    } else {
        // ===== epilogue: warp w owns TMEM lanes 32*(w%4) .. +31 (a hardware rule) and 32 of the 128 columns of every
        // accumulator stage (slice).  The MMA warp may run up to four stages ahead of the slowest epilogue warp. =====
        const uint32_t quarter = (uint32_t)warp & 3u;
        const uint32_t slice = (uint32_t)(warp - 2) >> 2;            // 0..3
        uint32_t tile_iter = 0, acc_it = 0;
        const int dterm = -255 * (int)p.dim;
        u64* my_key = st_key[warp - 2];
        uint32_t* my_q = st_q[warp - 2];
        int* my_u = s_uw + (size_t)(warp - 2) * 128;
        uint32_t staged = 0;                                         // warp-uniform
        auto flush = [&]() {
            __syncwarp();
            for (uint32_t base = 0; base < staged; base += 32) {
                const uint32_t e = base + (uint32_t)lane;
                if (e < staged) {
                    const uint32_t qi = my_q[e];
                    const u64 key = my_key[e];
                    const uint32_t slot = atomicAdd(p.cand_cnt + qi, 1u);
                    if (slot < p.cap) p.cand[(size_t)qi * p.cap + slot] = key;
                    else p.overflow[qi] = 1u;
                    // every accepted key is counted once in its query's kappa histogram
                    const float kap = __fmul_rn(key64_kappa(key), p.inv_q[qi]);
                    const int bin = min(max(__float2int_rd(__fmul_rn(__fadd_rn(kap, 1.0f), (float)(kBatchHistBins / 2))), 0), (int)kBatchHistBins - 1);
                    atomicAdd(p.bhist + (size_t)qi * kBatchHistBins + bin, 1u);
                }
            }
            staged = 0;
            __syncwarp();
        };
        // In-round tightening.  At refresh points this CTA turns the histograms of the queries it is responsible for
        // (column j with j % CTAs-in-group == its index) into thresholds -- the lower edge of the highest bin with at
        // least `keep` accepted keys at or above it, a valid bound because those keys are real rows -- publishes them
        // with an atomic max, and every epilogue thread re-reads the published threshold of one column.
        const uint32_t ethread = (uint32_t)threadIdx.x - 64u;          // 0 .. 32 * kBatchEpiWarps - 1
        auto refresh = [&]() {
            flush();
            for (uint32_t j = ethread; j < QG; j += 32u * kBatchEpiWarps) {
                const uint32_t qi = g * QG + j;
                if (j % cstride == ci % cstride && s_invq[j] > 0.0f) {
                    const uint4* hp = reinterpret_cast<const uint4*>(p.bhist + (size_t)qi * kBatchHistBins);
                    uint32_t run = 0;
                    int bstar = -1;
                    for (int c4 = (int)kBatchHistBins / 4 - 1; c4 >= 0 && bstar < 0; --c4) {
                        const uint4 v = __ldcg(hp + c4);
                        const uint32_t h[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int i = 3; i >= 0; --i) {
                            run += h[i];
                            if (bstar < 0 && run >= p.keep) bstar = 4 * c4 + i;
                        }
                    }
                    if (bstar > 0) {
                        // kappa-bin edge back to kappa' units, nudged down so that no key of bin >= b* falls below it
                        const float edge = (float)bstar * (2.0f / (float)kBatchHistBins) - 1.0f;
                        float t = edge / s_invq[j];
                        t = t - fabsf(t) * 4.0e-6f - 1.0e-3f;
                        // float atomic max (thresholds may be negative): compare-and-swap on the bit pattern
                        float* addr = p.thr_live + qi;
                        float old = *reinterpret_cast<volatile float*>(addr);
                        while (t > old) {
                            const uint32_t prev = atomicCAS(reinterpret_cast<uint32_t*>(addr), __float_as_uint(old), __float_as_uint(t));
                            if (prev == __float_as_uint(old)) break;
                            old = __uint_as_float(prev);
                        }
                    }
                }
                const float live = *reinterpret_cast<volatile float*>(p.thr_live + qi);
                if (live > s_thr[j]) { s_thr[j] = live; s_pre[j].x = fminf(fmaxf(live, -1.0e30f), 1.0e30f); }
            }
            __syncwarp();
        };
        // per-row metadata of a tile: fetched one tile ahead, so its L2 round trip hides behind the current tile
        auto fetch_meta = [&](uint32_t tile, float& inv, int& rs) {
            const uint32_t rw = tile * kBatchTileRows + quarter * 32u + (uint32_t)lane;
            const bool ok = tile < p.tile_end && rw < p.n;
            inv = ok ? __ldg(p.inv_norm + rw) : 0.0f;
            rs = ok ? __ldg(p.row_sum + rw) : 0;
        };
        float inv_next;
        int rs_next;
        fetch_meta(p.tile_begin + ci, inv_next, rs_next);
        for (uint32_t t = p.tile_begin + ci; t < p.tile_end; t += cstride, ++tile_iter) {
            if (tile_iter >= 4 && ((tile_iter & (tile_iter - 1)) == 0 || (tile_iter & 63u) == 0)) refresh();
            const uint32_t row = t * kBatchTileRows + quarter * 32u + (uint32_t)lane;
            const bool row_ok = row < p.n;
            const float inv_r = inv_next;
            const int rowterm = dterm + 2 * rs_next;
            fetch_meta(t + cstride, inv_next, rs_next);
            if constexpr (!FLOOD) {
                // Integer pre-test on the raw accumulator.  A score passes when fl(fl(dot_i) * inv_r) >= thr with
                // dot_i = 4 S + rowterm + colterm.  With norm = 1 / inv_r inside [norm_lo, norm_hi] and rowterm <= rt_max
                // over the 32 rows of this warp, passing implies
                //     S >= (thr * (thr >= 0 ? norm_lo : norm_hi) - colterm - rt_max - slack) / 4 =: v[col]
                // (slack covers every rounding), so ONE integer compare per score never rejects a passing one; the
                // survivors take the exact test.  rowterm = 2 sum r - 255 d moves little from row to row, so the bound
                // loses almost nothing to rt_max; what it loses to the norm spread is the price of a per-warp bound.
                float inv_lo = row_ok ? inv_r : __int_as_float(0x7f800000), inv_hi = row_ok ? inv_r : 0.0f;
                int rt_max = row_ok ? rowterm : INT_MIN;
#pragma unroll
                for (int off = 16; off; off >>= 1) {
                    inv_lo = fminf(inv_lo, __shfl_xor_sync(0xFFFFFFFFu, inv_lo, off));
                    inv_hi = fmaxf(inv_hi, __shfl_xor_sync(0xFFFFFFFFu, inv_hi, off));
                    rt_max = max(rt_max, __shfl_xor_sync(0xFFFFFFFFu, rt_max, off));
                }
                const float norm_lo = inv_hi > 0.0f ? 1.0f / inv_hi : 0.0f;
                const float norm_hi = inv_hi > 0.0f ? 1.0f / inv_lo : 0.0f;
                const float rt_f = (float)rt_max;                       // |rowterm| < 2^24: exact
                __syncwarp();
                // a handful of float ops per column; the float -> int conversion saturates, so the clamped +-1e30
                // thresholds (-inf: nothing seen yet, +inf: padding column) become INT_MIN / INT_MAX
#pragma unroll
                for (uint32_t sg = 0; sg < (uint32_t)kBatchAccStages; ++sg) {
                    if (sg < NS) {
                        const uint32_t col = sg * kBatchAccCols + slice * 32u + (uint32_t)lane;
                        const float2 pc = s_pre[col];
                        const float tt = pc.x * (pc.x >= 0.0f ? norm_lo : norm_hi);
                        const float y = (tt - pc.y) - rt_f;
                        const float m = fabsf(tt) + fabsf(pc.y) + fabsf(rt_f);   // every rounding above is relative to one of these
                        my_u[sg * 32u + (uint32_t)lane] = __float2int_rd(0.25f * (fmaf(m, -4.0e-6f, y) - 8.0f));
                    }
                }
                __syncwarp();
            }
            for (uint32_t sg = 0; sg < NS; ++sg, ++acc_it) {
                const uint32_t ab = acc_it % kBatchAccStages;
                mbar_wait(&acc_full[ab], (acc_it / kBatchAccStages) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                do {                                                  // one 32-column block per warp and stage
                    uint32_t r[32];
                    tmem_ld32(tmem + ((quarter * 32u) << 16) + ab * kBatchAccCols + slice * 32u, r);
                    // the scores are in registers: hand the accumulator back to the MMA warp before looking at them
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[ab]);
                    const uint32_t colbase = sg * kBatchAccCols + slice * 32u;
                    const int* my_us = my_u + sg * 32u;
                    if constexpr (FLOOD) {
                        // round 0: every score of a real query is a candidate; its slot in the query's buffer is its row
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const uint32_t col = colbase + (uint32_t)i;
                            const int dot_i = 4 * (int)r[i] + rowterm + s_colterm[col];
                            const float kf = __fmul_rn((float)dot_i, inv_r);
                            if (row_ok && s_invq[col] > 0.0f) p.cand[(size_t)(g * QG + col) * p.cap + row] = make_key64(kf, row);
                        }
                        continue;
                    }
                    // one compare per score against the raw accumulator, OR-ed into two predicates (two dependency chains)
                    bool some = false, some2 = false;
#pragma unroll
                    for (int i4 = 0; i4 < 8; ++i4) {
                        const int4 u4 = *reinterpret_cast<const int4*>(my_us + 4 * i4);
                        some |= ((int)r[4 * i4 + 0] >= u4.x);
                        some2 |= ((int)r[4 * i4 + 1] >= u4.y);
                        some |= ((int)r[4 * i4 + 2] >= u4.z);
                        some2 |= ((int)r[4 * i4 + 3] >= u4.w);
                    }
                    if (!__any_sync(0xFFFFFFFFu, (some || some2) && row_ok)) continue;
                    // some lane passed the pre-test in some column: find which (bounds re-read through an opaque pointer,
                    // so that the compiler does not keep the 32 bounds of the fast pass alive across the branch)
                    uint32_t mask = 0;
                    const int* u_s = my_us;
                    asm volatile("" : "+l"(u_s));
#pragma unroll
                    for (int i4 = 0; i4 < 8; ++i4) {
                        const int4 u4 = *reinterpret_cast<const int4*>(u_s + 4 * i4);
                        const int us[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if ((int)r[4 * i4 + j] >= us[j]) mask |= 1u << (4 * i4 + j);
                    }
                    if (!row_ok) mask = 0;
                    // rare: for every column in which some lane survived the pre-test, those lanes take the exact test
                    uint32_t any = __reduce_or_sync(0xFFFFFFFFu, mask);
                    while (any) {
                        const uint32_t i = (uint32_t)__ffs(any) - 1u;       // warp-uniform
                        any &= any - 1u;
                        uint32_t sv = 0;
                        switch (i) {
#define PBX_SEL(n) case n: sv = r[n]; break;
                            PBX_SEL(0) PBX_SEL(1) PBX_SEL(2) PBX_SEL(3) PBX_SEL(4) PBX_SEL(5) PBX_SEL(6) PBX_SEL(7)
                            PBX_SEL(8) PBX_SEL(9) PBX_SEL(10) PBX_SEL(11) PBX_SEL(12) PBX_SEL(13) PBX_SEL(14) PBX_SEL(15)
                            PBX_SEL(16) PBX_SEL(17) PBX_SEL(18) PBX_SEL(19) PBX_SEL(20) PBX_SEL(21) PBX_SEL(22) PBX_SEL(23)
                            PBX_SEL(24) PBX_SEL(25) PBX_SEL(26) PBX_SEL(27) PBX_SEL(28) PBX_SEL(29) PBX_SEL(30) PBX_SEL(31)
#undef PBX_SEL
                        }
                        const uint32_t col = colbase + i;
                        bool hit = false;
                        float kf = 0.0f;
                        if (mask & (1u << i)) {
                            const int dot_i = 4 * (int)sv + rowterm + s_colterm[col];
                            kf = __fmul_rn((float)dot_i, inv_r);
                            hit = kf >= s_thr[col];
                        }
                        const uint32_t hb = __ballot_sync(0xFFFFFFFFu, hit);
                        if (hit) {
                            const uint32_t pos = staged + (uint32_t)__popc(hb & ((1u << lane) - 1u));
                            my_key[pos] = make_key64(kf, row);
                            my_q[pos] = g * QG + col;
                        }
                        staged += (uint32_t)__popc(hb);
                        if (staged >= 32) flush();
                    }
                } while (false);
            }
        }
        flush();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

// ---- between rounds: cut every query's candidate buffer back to `keep` and tighten its threshold ----------
struct BatchTightenParams {
    u64* cand;
    uint32_t* cand_cnt;
    float* thr;
    uint32_t keep;
    uint32_t nq;
    uint32_t cap;               // candidate buffer entries per query
};

__global__ void __launch_bounds__(256)
batch_tighten_kernel(const BatchTightenParams p) {
    extern __shared__ __align__(16) unsigned char tsm[];
    u64* buf = reinterpret_cast<u64*>(tsm);              // [cap]
    __shared__ uint32_t s_cnt;
    __shared__ u64 s_tau;
    __shared__ SelectScratch sel;
    const uint32_t q = blockIdx.x;
    const uint32_t c = min(p.cand_cnt[q], p.cap);
    if (c <= p.keep) return;                             // nothing to cut: the threshold stays where it is
    u64* src = p.cand + (size_t)q * p.cap;
    for (uint32_t i = threadIdx.x; i < c; i += blockDim.x) buf[i] = src[i];
    if (threadIdx.x == 0) { s_cnt = c; s_tau = 0; }
    __syncthreads();
    TopBuf<u64> tb{buf, &s_cnt, &s_tau, p.cap, p.keep};
    block_select_top(tb, &sel);
    const uint32_t kept = s_cnt;
    for (uint32_t i = threadIdx.x; i < kept; i += blockDim.x) src[i] = buf[i];
    if (threadIdx.x == 0) {
        p.cand_cnt[q] = kept;
        p.thr[q] = key64_kappa(s_tau);                   // every kept key has kappa' >= this
    }
}

}  // namespace pbx
