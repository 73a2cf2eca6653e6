/* TSAN host test ("race detection"): the reference's concurrency on this path is a search on the
 * UI thread while the indexer's writer thread inserts (src/ui/search.rs:22, src/engine.rs:186-203).  Two threads drive
 * the C ABI the same way: one appends blocks of rows, one searches; both the host code of this driver and of the
 * library (built with -fsanitize=thread, see tools/tsan.sh) are instrumented.
 *   gcc -O1 -g -fsanitize=thread -Iinclude tools/tsan_host.c -o tools/bin/tsan_host -Lpixelbox_b200/lib/exp -l:lib_tsan.so -lpthread */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "pixelbox_b200.h"

#define DIM 256
#define N 40000
#define BLOCK 500
static pbx_corpus* corpus;
static uint8_t* rows;
static int64_t* ids;
static volatile int writer_done = 0;
static int failures = 0;

static void* writer(void* arg) {
    (void)arg;
    for (uint64_t at = 0; at < N; at += BLOCK)
        if (pbx_corpus_append(corpus, ids + at, rows + at * DIM, BLOCK) != PBX_OK) { fprintf(stderr, "append: %s\n", pbx_last_error()); __atomic_add_fetch(&failures, 1, __ATOMIC_SEQ_CST); }
    __atomic_store_n(&writer_done, 1, __ATOMIC_SEQ_CST);
    return NULL;
}

static void* reader(void* arg) {
    (void)arg;
    int64_t out_ids[2 * 50];
    float out_dist[2 * 50];
    uint32_t cnt[2];
    int searches = 0;
    while (!__atomic_load_n(&writer_done, __ATOMIC_SEQ_CST) || searches < 20) {
        uint64_t before = 0, after = 0;
        pbx_corpus_size(corpus, &before);
        if (pbx_search(corpus, rows + 123 * DIM, 2, 50, 1e3, out_ids, out_dist, NULL, NULL, cnt) != PBX_OK) { fprintf(stderr, "search: %s\n", pbx_last_error()); __atomic_add_fetch(&failures, 1, __ATOMIC_SEQ_CST); }
        pbx_corpus_size(corpus, &after);
        /* a search sees a committed prefix: every id it returns had been appended when it finished */
        for (uint32_t i = 0; i < cnt[0]; ++i)
            if (out_ids[i] < 1 || (uint64_t)out_ids[i] > after) { fprintf(stderr, "id %lld beyond the committed prefix %llu\n", (long long)out_ids[i], (unsigned long long)after); __atomic_add_fetch(&failures, 1, __ATOMIC_SEQ_CST); }
        ++searches;
    }
    printf("reader: %d searches concurrent with the writer\n", searches);
    return NULL;
}

int main(void) {
    rows = malloc((size_t)N * DIM);
    ids = malloc((size_t)N * sizeof(int64_t));
    uint64_t x = 88172645463325252ull;
    for (size_t i = 0; i < (size_t)N * DIM; ++i) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; rows[i] = (uint8_t)(x >> 24); }
    for (int i = 0; i < N; ++i) ids[i] = i + 1;
    if (pbx_corpus_create(DIM, 1000, 0, &corpus) != PBX_OK) { fprintf(stderr, "create: %s\n", pbx_last_error()); return 2; }
    pthread_t w, r;
    pthread_create(&w, NULL, writer, NULL);
    pthread_create(&r, NULL, reader, NULL);
    pthread_join(w, NULL);
    pthread_join(r, NULL);
    uint64_t n = 0;
    pbx_corpus_size(corpus, &n);
    pbx_corpus_destroy(corpus);
    printf("tsan_host: %llu rows, %d failures\n", (unsigned long long)n, failures);
    return failures ? 1 : 0;
}
