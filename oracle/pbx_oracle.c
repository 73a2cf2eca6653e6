/*
 * pbx_oracle.c -- CPU restatement of PixelBox's similarity-search hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product library
 * (pixelbox_b200/csrc) never links or calls anything in this directory.
 *
 * What it restates (all citations relative to the reference checkout):
 *   - cosine_distance               src/engine.rs:572-588   (f32, strict left-to-right folds)
 *   - the SQLite UDF wrapper        src/engine.rs:608-622   (blob copies, f32 -> f64 widening)
 *   - the similarity query          src/engine.rs:375-383   (WHERE dist < ? ORDER BY dist ASC LIMIT k)
 *   - the quantizer                 src/image_hashes/efficientnet.rs:39 (f32 -> u8, centre 128)
 *
 * PARITY PINNING STATUS: the reference cannot be compiled here (no rustc/cargo in the image),
 * and its own tests pin this path only through three inequality asserts
 * (src/engine.rs:705-707) plus the README encoding example (README.md:54).  Those are all
 * checked in tests/test_oracle.py.  Beyond them the arithmetic is pinned by an independent
 * numpy-float32 restatement (tests/np_restatement.py) and the selection semantics
 * (filter / order / limit / tie order) by running the verbatim SQL of src/engine.rs:375-382
 * inside Python's SQLite with this file's UDF (tests/sqlite_oracle.py).  So: arithmetic
 * "pinned by KATs + independent restatement", top-k order "parity unpinned by upstream tests,
 * pinned here against SQLite 3.45.1".
 *
 * Build: see oracle/Makefile.  -O2 -ffp-contract=off, never -ffast-math: Rust never contracts
 * a*b+c into an FMA and never reassociates float sums, so neither may we.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

#define PBX_ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * R1: cosine_distance, literal.  src/engine.rs:572-588.
 *
 *   let u8_to_float = |u8s| u8s.iter().map(|v| ((*v as f32 / 255.0) * 2.0) - 1.0).collect::<Vec<f32>>();   :575-577
 *   let magnitude = a.fold(0f32, |i,x| i + x*x).sqrt() * b.fold(0f32, |i,x| i + x*x).sqrt();               :580-581
 *   if magnitude < 1e-6 { return 0.0; }                                                                     :582-584
 *   let dot = a.zip(b).fold(0f32, |i,(a,b)| i + a*b);                                                      :585
 *   (1.0 / (dot / magnitude).max(1e-6)) - 1.0                                                              :586-587
 *
 * The two Vec<f32> heap allocations are kept on purpose: this function is also the timed CPU
 * baseline and the reference pays for them on every row.
 * ------------------------------------------------------------------------------------------ */
static inline float decode_u8(uint8_t v) {
    float x = (float)v;      /* `*v as f32` : exact */
    x = x / 255.0f;          /* f32 division, round-to-nearest */
    x = x * 2.0f;            /* exact */
    x = x - 1.0f;            /* rounded */
    return x;
}

static inline float rust_f32_max(float a, float b) {
    /* f32::max: if one argument is NaN the other is returned. */
    if (a != a) return b;
    if (b != b) return a;
    return a > b ? a : b;
}

PBX_ORACLE_API
float pbx_oracle_cosine_distance(const uint8_t* hash_a, size_t len_a, const uint8_t* hash_b, size_t len_b) {
    float* a = (float*)malloc((len_a ? len_a : 1) * sizeof(float));
    float* b = (float*)malloc((len_b ? len_b : 1) * sizeof(float));
    for (size_t i = 0; i < len_a; ++i) a[i] = decode_u8(hash_a[i]);
    for (size_t i = 0; i < len_b; ++i) b[i] = decode_u8(hash_b[i]);

    float sa = 0.0f;
    for (size_t i = 0; i < len_a; ++i) { float sq = a[i] * a[i]; sa = sa + sq; }
    float sb = 0.0f;
    for (size_t i = 0; i < len_b; ++i) { float sq = b[i] * b[i]; sb = sb + sq; }
    float magnitude = sqrtf(sa) * sqrtf(sb);
    float result;
    if (magnitude < 1e-6f) {
        result = 0.0f;
    } else {
        size_t n = len_a < len_b ? len_a : len_b; /* zip stops at the shorter */
        float dot = 0.0f;
        for (size_t i = 0; i < n; ++i) { float p = a[i] * b[i]; dot = dot + p; }
        float cosine_similarity = dot / magnitude;
        result = (1.0f / rust_f32_max(cosine_similarity, 1e-6f)) - 1.0f;
    }
    free(a);
    free(b);
    return result;
}

/* The f32 cosine the reference forms on the way (before 1/max(.)-1); exposed so tests can
 * measure |cos_f32 - cos_exact| against the error bound the GPU certificate relies on. */
PBX_ORACLE_API
float pbx_oracle_cosine_similarity_f32(const uint8_t* hash_a, const uint8_t* hash_b, size_t d) {
    float sa = 0.0f, sb = 0.0f, dot = 0.0f;
    for (size_t i = 0; i < d; ++i) { float x = decode_u8(hash_a[i]); float sq = x * x; sa = sa + sq; }
    for (size_t i = 0; i < d; ++i) { float x = decode_u8(hash_b[i]); float sq = x * x; sb = sb + sq; }
    float magnitude = sqrtf(sa) * sqrtf(sb);
    if (magnitude < 1e-6f) return 0.0f;
    for (size_t i = 0; i < d; ++i) { float p = decode_u8(hash_a[i]) * decode_u8(hash_b[i]); dot = dot + p; }
    return dot / magnitude;
}

/* R2: the SQLite scalar UDF body, src/engine.rs:613-620: copies both blobs (`to_vec`),
 * calls cosine_distance, widens f32 -> f64. */
PBX_ORACLE_API
double pbx_oracle_udf_cosine_distance(const uint8_t* lhs, size_t len_l, const uint8_t* rhs, size_t len_r) {
    uint8_t* l = (uint8_t*)malloc(len_l ? len_l : 1);
    uint8_t* r = (uint8_t*)malloc(len_r ? len_r : 1);
    memcpy(l, lhs, len_l);
    memcpy(r, rhs, len_r);
    float dist = pbx_oracle_cosine_distance(l, len_l, r, len_r);
    free(l);
    free(r);
    return (double)dist;
}

/* ------------------------------------------------------------------------------------------
 * Exact-integer restatement: c(v) = 2v - 255, the 1/255 scales cancel.
 *   dot_i = sum c(a)c(b),  n(a) = sum c(a)^2,  cos = dot_i / sqrt(n(a) n(b)).
 * ------------------------------------------------------------------------------------------ */
PBX_ORACLE_API
void pbx_oracle_int_terms(const uint8_t* q, const uint8_t* r, size_t d, int32_t* dot, int32_t* norm2_q, int32_t* norm2_r) {
    int64_t s = 0, nq = 0, nr = 0;
    for (size_t i = 0; i < d; ++i) {
        int64_t cq = 2 * (int64_t)q[i] - 255;
        int64_t cr = 2 * (int64_t)r[i] - 255;
        s += cq * cr;
        nq += cq * cq;
        nr += cr * cr;
    }
    *dot = (int32_t)s;
    *norm2_q = (int32_t)nq;
    *norm2_r = (int32_t)nr;
}

PBX_ORACLE_API
double pbx_oracle_cosine_exact(const uint8_t* q, const uint8_t* r, size_t d) {
    int32_t dot, nq, nr;
    pbx_oracle_int_terms(q, r, d, &dot, &nq, &nr);
    if (nq == 0 || nr == 0) return 0.0;
    long double m = sqrtl((long double)nq) * sqrtl((long double)nr);
    return (double)((long double)dot / m);
}

/* ------------------------------------------------------------------------------------------
 * R3: the similarity query, src/engine.rs:375-383, as a bare loop:
 *   scan rows, dist = cosine_distance(query, hash) widened to f64, keep dist < max_dist,
 *   ORDER BY dist ASC (ties: ascending image_id -- the order SQLite 3.45.1 produces for this
 *   SQL, checked in tests/test_sqlite_oracle.py, and the rule BASELINE.json's north_star states),
 *   LIMIT k.
 * Outputs: ids, f32 distance, exact int dot and row norm^2 of every returned row.
 * ------------------------------------------------------------------------------------------ */
typedef struct { float dist; int64_t id; uint64_t row; } hit_t;

static inline int hit_less(const hit_t* a, const hit_t* b) {
    if ((double)a->dist < (double)b->dist) return 1;
    if ((double)a->dist > (double)b->dist) return 0;
    return a->id < b->id;
}

/* bounded insertion list: keeps the k smallest under hit_less */
static void topk_insert(hit_t* heap, uint32_t* count, uint32_t k, hit_t h) {
    if (k == 0) return;
    if (*count == k && !hit_less(&h, &heap[k - 1])) return;
    uint32_t pos = (*count < k) ? (*count)++ : k - 1;
    while (pos > 0 && hit_less(&h, &heap[pos - 1])) { heap[pos] = heap[pos - 1]; --pos; }
    heap[pos] = h;
}

static void scan_range(const uint8_t* corpus, const int64_t* ids, uint64_t r0, uint64_t r1, uint32_t d,
                       const uint8_t* query, uint32_t k, double max_dist, hit_t* heap, uint32_t* count) {
    for (uint64_t r = r0; r < r1; ++r) {
        float dist = pbx_oracle_cosine_distance(query, d, corpus + r * (uint64_t)d, d);
        if (!((double)dist < max_dist)) continue;                       /* WHERE dist < ?  :379 */
        hit_t h = { dist, ids ? ids[r] : (int64_t)(r + 1), r };
        topk_insert(heap, count, k, h);
    }
}

typedef struct {
    const uint8_t* corpus; const int64_t* ids; uint64_t r0, r1; uint32_t d; const uint8_t* query;
    uint32_t k; double max_dist; hit_t* heap; uint32_t count;
} scan_job_t;

static void* scan_job_main(void* arg) {
    scan_job_t* j = (scan_job_t*)arg;
    j->count = 0;
    scan_range(j->corpus, j->ids, j->r0, j->r1, j->d, j->query, j->k, j->max_dist, j->heap, &j->count);
    return NULL;
}

/* threads == 1 is the faithful baseline: upstream runs the whole scan inside one SQLite statement
 * on the UI thread (rayon is commented out, Cargo.toml:20).  threads > 1 splits the rows over
 * pthreads and merges the per-thread lists under the same order -- a courtesy baseline. */
PBX_ORACLE_API
int pbx_oracle_topk(const uint8_t* corpus, const int64_t* ids, uint64_t n, uint32_t d, const uint8_t* query,
                    uint32_t k, double max_dist, int threads,
                    int64_t* out_ids, float* out_dist, int32_t* out_dot, int32_t* out_norm2, uint32_t* out_count) {
    if (k == 0) { *out_count = 0; return 0; }
    int nt = threads > 0 ? threads : 1;
    if (nt > 1024) nt = 1024;
    hit_t* heaps = (hit_t*)malloc((size_t)nt * k * sizeof(hit_t));
    scan_job_t* jobs = (scan_job_t*)calloc((size_t)nt, sizeof(scan_job_t));
    pthread_t* tids = (pthread_t*)calloc((size_t)nt, sizeof(pthread_t));
    if (!heaps || !jobs || !tids) { free(heaps); free(jobs); free(tids); return -1; }
    for (int t = 0; t < nt; ++t) {
        scan_job_t j = { corpus, ids, n * (uint64_t)t / (uint64_t)nt, n * (uint64_t)(t + 1) / (uint64_t)nt,
                         d, query, k, max_dist, heaps + (size_t)t * k, 0 };
        jobs[t] = j;
    }
    if (nt == 1) {
        scan_job_main(&jobs[0]);
    } else {
        for (int t = 0; t < nt; ++t) pthread_create(&tids[t], NULL, scan_job_main, &jobs[t]);
        for (int t = 0; t < nt; ++t) pthread_join(tids[t], NULL);
    }
    /* merge per-thread lists under the same (dist, id) order */
    hit_t* fin = heaps;
    uint32_t cnt = jobs[0].count;
    for (int t = 1; t < nt; ++t)
        for (uint32_t i = 0; i < jobs[t].count; ++i) topk_insert(fin, &cnt, k, heaps[(size_t)t * k + i]);
    for (uint32_t i = 0; i < cnt; ++i) {
        out_ids[i] = fin[i].id;
        out_dist[i] = fin[i].dist;
        int32_t dot, nq, nr;
        pbx_oracle_int_terms(query, corpus + fin[i].row * (uint64_t)d, d, &dot, &nq, &nr);
        if (out_dot) out_dot[i] = dot;
        if (out_norm2) out_norm2[i] = nr;
    }
    *out_count = cnt;
    free(heaps);
    free(jobs);
    free(tids);
    return 0;
}

/* distances of every row (for small-corpus tests / full-order checks) */
PBX_ORACLE_API
void pbx_oracle_all_distances(const uint8_t* corpus, uint64_t n, uint32_t d, const uint8_t* query, float* out) {
    for (uint64_t r = 0; r < n; ++r) out[r] = pbx_oracle_cosine_distance(query, d, corpus + r * (uint64_t)d, d);
}

/* ------------------------------------------------------------------------------------------
 * The other two registered distances.
 * byte_distance, src/engine.rs:590-592: f32 fold of |a - b| (every partial sum is an integer below 2^24,
 *   so the fold is exact), divided by (255f32 * len as f32).
 * hamming_distance, src/engine.rs:594-604: per byte the number of differing bits as a u8, `.sum::<u8>()`,
 *   divided by (8f32 * len as f32).  The u8 sum overflows past 255 differing bits: a debug build panics,
 *   a release build wraps modulo 256; this restatement wraps (the behaviour of the shipped binary) and
 *   reports the true bit count separately.  Upstream KATs: test_hamming_distance, src/engine.rs:693-701.
 * ------------------------------------------------------------------------------------------ */
PBX_ORACLE_API
float pbx_oracle_byte_distance(const uint8_t* a, const uint8_t* b, size_t len) {
    float acc = 0.0f;
    for (size_t i = 0; i < len; ++i) {
        float d = (float)a[i] - (float)b[i];
        acc = acc + (d < 0.0f ? -d : d);
    }
    return acc / (255.0f * (float)len);
}

PBX_ORACLE_API
float pbx_oracle_hamming_distance(const uint8_t* a, const uint8_t* b, size_t len, uint32_t* true_bits) {
    uint8_t sum = 0;
    uint32_t all = 0;
    for (size_t i = 0; i < len; ++i) {
        uint8_t diff = (uint8_t)(a[i] ^ b[i]);
        uint8_t bits_set = 0;
        while (diff != 0) { bits_set = (uint8_t)(bits_set + (diff & 1)); diff >>= 1; }
        sum = (uint8_t)(sum + bits_set);          /* wrapping, as in a release build */
        all += bits_set;
    }
    if (true_bits) *true_bits = all;
    return (float)sum / (8.0f * (float)len);
}

/* ------------------------------------------------------------------------------------------
 * Quantizer, src/image_hashes/efficientnet.rs:39:
 *   128u8.saturating_add_signed((f*128.0f32).max(-128.0).min(128.0) as i8)
 * `as i8` from f32 truncates toward zero and saturates (so +128.0 -> 127); NaN -> 0.
 * ------------------------------------------------------------------------------------------ */
PBX_ORACLE_API
uint8_t pbx_oracle_quantize(float f) {
    float x = f * 128.0f;
    x = rust_f32_max(x, -128.0f);        /* NaN.max(-128) == -128, so x is a number from here on */
    x = x < 128.0f ? x : 128.0f;         /* f32::min */
    int i;
    if (x >= 127.0f) i = 127;            /* saturating `as i8` */
    else if (x <= -128.0f) i = -128;
    else i = (int)x; /* trunc toward zero */
    int v = 128 + i;
    if (v < 0) v = 0;
    if (v > 255) v = 255;
    return (uint8_t)v;
}

/* ------------------------------------------------------------------------------------------
 * Synthetic corpus definition (bench/test data, not reference behaviour): counter-based,
 * 8 bytes per splitmix64 output, keyed by (seed, global row, 8-byte chunk index).
 * The CUDA generator in pixelbox_b200/csrc and the numpy one in pixelbox_b200/synth.py
 * implement the same function; tests check all three agree.
 * ------------------------------------------------------------------------------------------ */
static inline uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

PBX_ORACLE_API
void pbx_oracle_synth_rows(uint64_t seed, uint64_t first_row, uint64_t nrows, uint32_t d, uint8_t* out) {
    uint32_t chunks = (d + 7) / 8;
    for (uint64_t r = 0; r < nrows; ++r) {
        uint64_t row = first_row + r;
        uint64_t rk = splitmix64(seed ^ (row * 0xD1342543DE82EF95ull));
        for (uint32_t c = 0; c < chunks; ++c) {
            uint64_t x = splitmix64(rk + (uint64_t)c);
            for (uint32_t b = 0; b < 8 && c * 8 + b < d; ++b) out[r * (uint64_t)d + c * 8 + b] = (uint8_t)(x >> (8 * b));
        }
    }
}

PBX_ORACLE_API
int pbx_oracle_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
