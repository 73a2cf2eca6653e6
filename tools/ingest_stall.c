/* Ingest under load (src/engine.rs:186-203 inserts one image at a time while the UI thread searches):
 * one thread appends N single rows through pbx_corpus_append, another searches in a loop.  Reports the append rate, the
 * search latency distribution, and the slowest searches / appends with their start times so that a stall can be attributed
 * (growth step, coalesced-block upload, first-use allocation).
 *   gcc -O2 -Iinclude tools/ingest_stall.c -o tools/bin/ingest_stall -Lpixelbox_b200/lib -l:libpixelbox_b200.so -lpthread -lm
 *   LD_LIBRARY_PATH=pixelbox_b200/lib tools/bin/ingest_stall [rows] [dim] [preload_rows] */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "pixelbox_b200.h"

static pbx_corpus* corpus;
static uint8_t* rows;
static int64_t* ids;
static uint64_t N = 1000000, PRE = 0;
static uint32_t DIM = 256;
static volatile int writer_done = 0;
static double t_origin;

static double now(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

typedef struct { double t, dt; } ev;
static ev* s_ev; static size_t s_n = 0, s_cap = 0;
static ev slow_app[16]; static int n_slow = 0;
static double app_secs = 0;

static void* writer(void* arg) {
    (void)arg;
    const double t0 = now();
    for (uint64_t i = PRE; i < N; ++i) {
        const double a = now();
        if (pbx_corpus_append(corpus, ids + i, rows + (i % 65536) * DIM, 1) != PBX_OK) { fprintf(stderr, "append: %s\n", pbx_last_error()); exit(2); }
        const double d = now() - a;
        if (d > 2e-4) {                       /* keep the 16 slowest */
            int at = n_slow < 16 ? n_slow++ : -1;
            if (at < 0) { int m = 0; for (int j = 1; j < 16; ++j) if (slow_app[j].dt < slow_app[m].dt) m = j; if (slow_app[m].dt < d) at = m; }
            if (at >= 0) { slow_app[at].t = a - t_origin; slow_app[at].dt = d; }
        }
    }
    pbx_corpus_flush(corpus);
    app_secs = now() - t0;
    __atomic_store_n(&writer_done, 1, __ATOMIC_SEQ_CST);
    return NULL;
}

static void* reader(void* arg) {
    (void)arg;
    int64_t out_ids[10]; float out_dist[10]; uint32_t cnt;
    while (!__atomic_load_n(&writer_done, __ATOMIC_SEQ_CST)) {
        const double a = now();
        if (pbx_search(corpus, rows + 5 * DIM, 1, 10, 1e3, out_ids, out_dist, NULL, NULL, &cnt) != PBX_OK) { fprintf(stderr, "search: %s\n", pbx_last_error()); exit(2); }
        const double d = now() - a;
        if (s_n == s_cap) { s_cap = s_cap ? 2 * s_cap : 65536; s_ev = realloc(s_ev, s_cap * sizeof(ev)); }
        s_ev[s_n].t = a - t_origin; s_ev[s_n].dt = d; ++s_n;
    }
    return NULL;
}

static int by_dt(const void* a, const void* b) { double x = ((const ev*)a)->dt, y = ((const ev*)b)->dt; return x < y ? -1 : x > y; }

int main(int argc, char** argv) {
    if (argc > 1) N = strtoull(argv[1], 0, 10);
    if (argc > 2) DIM = (uint32_t)atoi(argv[2]);
    if (argc > 3) PRE = strtoull(argv[3], 0, 10);
    N += PRE;                                   /* argv[1] = rows appended one by one, after PRE preloaded rows */
    rows = malloc((size_t)65536 * DIM);
    ids = malloc(N * sizeof(int64_t));
    uint64_t x = 88172645463325252ull;
    for (size_t i = 0; i < (size_t)65536 * DIM; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; rows[i] = (uint8_t)(x >> 24); }
    for (uint64_t i = 0; i < N; ++i) ids[i] = (int64_t)i + 1;
    if (pbx_corpus_create(DIM, 1024, 0, &corpus) != PBX_OK) { fprintf(stderr, "create: %s\n", pbx_last_error()); return 2; }
    for (uint64_t i = 0; i < PRE; i += 65536) {
        const uint64_t m = PRE - i < 65536 ? PRE - i : 65536;
        if (pbx_corpus_append(corpus, ids + i, rows, m) != PBX_OK) { fprintf(stderr, "preload: %s\n", pbx_last_error()); return 2; }
    }
    {   /* warm both paths: first-use allocations are not what is measured */
        int64_t oi[10]; float od[10]; uint32_t c;
        pbx_corpus_append(corpus, ids, rows, 0);
        for (int i = 0; i < 3; ++i) pbx_search(corpus, rows, 1, 10, 1e3, oi, od, NULL, NULL, &c);
    }
    t_origin = now();
    pthread_t w, r;
    pthread_create(&r, NULL, reader, NULL);
    pthread_create(&w, NULL, writer, NULL);
    pthread_join(w, NULL);
    pthread_join(r, NULL);
    uint64_t n = 0;
    pbx_corpus_size(corpus, &n);
    qsort(s_ev, s_n, sizeof(ev), by_dt);
    printf("{\"rows_appended\": %llu, \"dim\": %u, \"preloaded\": %llu, \"append_rows_per_s\": %.0f, \"searches\": %zu, "
           "\"search_ms_median\": %.4f, \"search_ms_p99\": %.4f, \"search_ms_p999\": %.4f, \"search_ms_max\": %.4f, \"searches_over_1ms\": %zu}\n",
           (unsigned long long)(N - PRE), DIM, (unsigned long long)PRE, (N - PRE) / app_secs, s_n,
           s_n ? 1e3 * s_ev[s_n / 2].dt : 0.0, s_n ? 1e3 * s_ev[(size_t)(s_n * 0.99)].dt : 0.0, s_n ? 1e3 * s_ev[(size_t)(s_n * 0.999)].dt : 0.0,
           s_n ? 1e3 * s_ev[s_n - 1].dt : 0.0, ({ size_t k = 0; for (size_t i = 0; i < s_n; ++i) k += s_ev[i].dt > 1e-3; k; }));
    for (size_t i = s_n > 6 ? s_n - 6 : 0; i < s_n; ++i) printf("  slow search: t=%.4f s  %.3f ms\n", s_ev[i].t, 1e3 * s_ev[i].dt);
    for (int i = 0; i < n_slow; ++i) printf("  slow append: t=%.4f s  %.3f ms\n", slow_app[i].t, 1e3 * slow_app[i].dt);
    pbx_corpus_destroy(corpus);
    return n == N ? 0 : 1;
}
