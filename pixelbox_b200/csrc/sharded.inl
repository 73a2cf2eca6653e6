// sharded.inl -- pbx_sharded_*: the row-sharded corpus of ONE process over several GPUs of one box (SURVEY.md 8b / 8e:
// the reference's host is a single process that calls Engine on the UI thread, src/ui/search.rs:20-31).  Included at the
// end of api.cu (one translation unit).
//
// One pbx_corpus shard per listed device, one persistent worker thread per shard.  A search copies the queries into
// pinned (portable) host memory once; every worker enqueues, on its shard's stream, H2D of the queries -> the shard's
// complete local search (fast pass + certificate + exact pass on demand: a locally exact top-k) -> a peer copy of its
// k records per query into the root device's gather buffer; the calling thread then makes the root stream wait for all
// shards' events, merges the lists under (dist, image_id) with merge_hits_kernel and copies the result to the host.
// No shard waits for another inside a kernel, so the path has no dependence on CTA dispatch order or co-residency (the
// fused peer-memory kernel of the multi-process driver, pbx_exchange_*, trades that for one launch less).  Devices may
// repeat in the list (several shards on one GPU: the 1-GPU test box).

#include <condition_variable>
#include <memory>
#include <thread>

struct pbx_sharded {
    struct Shard {
        pbx_corpus* c = nullptr;
        int device = 0;
        uint8_t* d_q = nullptr;            // [max_nq][dim]
        pbx_hit* d_hits = nullptr;         // [max_nq][max_k] local result
        uint32_t* d_cnt = nullptr;         // [max_nq]
        uint32_t* h_cnt = nullptr;         // pinned copy of the local counts (error markers)
        cudaEvent_t ev_done = nullptr;     // local search + peer copy enqueued on c->stream
        std::thread th;
        std::mutex mu;
        std::condition_variable cv;
        std::atomic<uint64_t> req{0}, done{0};
        std::atomic<bool> quit{false};
        int rc = PBX_OK;
        char err[256] = "";
    };
    uint32_t dim = 0;
    std::vector<std::unique_ptr<Shard>> shards;
    // the call in flight (written by the caller before req is bumped)
    uint32_t nq = 0, k = 0;
    double max_dist = 0.0;
    // buffers
    uint32_t cap_nq = 0, cap_k = 0;
    uint8_t* h_q = nullptr;                // pinned, portable: [cap_nq][dim]
    pbx_hit* d_gathered = nullptr;         // on the root device: [n_shards][cap_nq * cap_k] (laid out per call as [n_shards][nq][k])
    pbx_hit* d_out = nullptr;              // root: [cap_nq][cap_k] merged hits, then [cap_nq] counts
    pbx_hit* h_out = nullptr;              // pinned
    std::mutex mu;                         // one search / structural change at a time
};

static void sharded_worker(pbx_sharded* s, pbx_sharded::Shard* sh, uint32_t index) {
    uint64_t seen = 0;
    for (;;) {
        // wait for the next request: spin briefly (interactive queries arrive back to back), then sleep
        uint64_t want = sh->req.load(std::memory_order_acquire);
        for (int spin = 0; want == seen && spin < 20000 && !sh->quit.load(std::memory_order_relaxed); ++spin)
            want = sh->req.load(std::memory_order_acquire);
        if (want == seen) {
            std::unique_lock<std::mutex> lk(sh->mu);
            sh->cv.wait(lk, [&] { return sh->req.load(std::memory_order_acquire) != seen || sh->quit.load(); });
            want = sh->req.load(std::memory_order_acquire);
        }
        if (sh->quit.load() && want == seen) return;
        seen = want;
        // ---- the shard's half of a search ----
        int rc = PBX_OK;
        const uint32_t nq = s->nq, k = s->k;
        pbx_sharded::Shard* root = s->shards[0].get();
        cudaError_t e = cudaSetDevice(sh->device);
        cudaStream_t st = sh->c->stream;
        if (e == cudaSuccess) e = cudaMemcpyAsync(sh->d_q, s->h_q, (size_t)nq * s->dim, cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) {
            rc = pbx_search_device(sh->c, sh->d_q, nq, k, s->max_dist, sh->d_hits, sh->d_cnt, nullptr);
            if (rc != PBX_OK) snprintf(sh->err, sizeof(sh->err), "%s", pbx_last_error());
        }
        if (e == cudaSuccess && rc == PBX_OK) {
            pbx_hit* dst = s->d_gathered + (size_t)index * nq * k;
            e = cudaMemcpyPeerAsync(dst, root->device, sh->d_hits, sh->device, (size_t)nq * k * sizeof(pbx_hit), st);
        }
        if (e == cudaSuccess && rc == PBX_OK) e = cudaMemcpyAsync(sh->h_cnt, sh->d_cnt, (size_t)nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess && rc == PBX_OK) e = cudaEventRecord(sh->ev_done, st);
        if (e != cudaSuccess) {
            rc = PBX_E_CUDA;
            snprintf(sh->err, sizeof(sh->err), "shard %u on device %d: %s", index, sh->device, cudaGetErrorString(e));
            cudaGetLastError();
        }
        sh->rc = rc;
        sh->done.store(seen, std::memory_order_release);
    }
}

static int sharded_ensure_buffers(pbx_sharded* s, uint32_t nq, uint32_t k) {
    if (nq <= s->cap_nq && k <= s->cap_k) return PBX_OK;
    const uint32_t cnq = std::max<uint32_t>(std::max(nq, s->cap_nq), 64u), ck = std::max<uint32_t>(std::max(k, s->cap_k), 100u);
    const size_t n_sh = s->shards.size();
    for (auto& sh : s->shards) {
        CU_TRY(cudaSetDevice(sh->device));
        CU_TRY(cudaDeviceSynchronize());
        cudaFree(sh->d_q); cudaFree(sh->d_hits); cudaFree(sh->d_cnt); cudaFreeHost(sh->h_cnt);
        sh->d_q = nullptr; sh->d_hits = nullptr; sh->d_cnt = nullptr; sh->h_cnt = nullptr;
        CU_TRY(cudaMalloc(&sh->d_q, (size_t)cnq * s->dim));
        CU_TRY(cudaMalloc(&sh->d_hits, (size_t)cnq * ck * sizeof(pbx_hit)));
        CU_TRY(cudaMalloc(&sh->d_cnt, (size_t)cnq * sizeof(uint32_t)));
        CU_TRY(cudaHostAlloc(&sh->h_cnt, (size_t)cnq * sizeof(uint32_t), cudaHostAllocPortable));
    }
    CU_TRY(cudaSetDevice(s->shards[0]->device));
    cudaFree(s->d_gathered); cudaFree(s->d_out); cudaFreeHost(s->h_q); cudaFreeHost(s->h_out);
    s->d_gathered = nullptr; s->d_out = nullptr; s->h_q = nullptr; s->h_out = nullptr; s->cap_nq = 0; s->cap_k = 0;
    const size_t out_records = (size_t)cnq * ck + ((size_t)cnq * sizeof(uint32_t) + sizeof(pbx_hit) - 1) / sizeof(pbx_hit) + 1;
    CU_TRY(cudaMalloc(&s->d_gathered, n_sh * cnq * ck * sizeof(pbx_hit)));
    CU_TRY(cudaMalloc(&s->d_out, out_records * sizeof(pbx_hit)));
    CU_TRY(cudaHostAlloc(&s->h_q, (size_t)cnq * s->dim, cudaHostAllocPortable));
    CU_TRY(cudaHostAlloc(&s->h_out, out_records * sizeof(pbx_hit), cudaHostAllocPortable));
    s->cap_nq = cnq; s->cap_k = ck;
    return PBX_OK;
}

extern "C" int pbx_sharded_create(uint32_t dim, uint64_t capacity_hint, const int* devices, int n_devices, pbx_sharded** out) {
    if (!out) return fail(PBX_E_INVALID, "out is NULL");
    *out = nullptr;
    if (!devices || n_devices <= 0 || n_devices > (int)PBX_MAX_SHARDS) return fail(PBX_E_INVALID, "1 <= n_devices <= %u required", PBX_MAX_SHARDS);
    pbx_sharded* s = new (std::nothrow) pbx_sharded();
    if (!s) return fail(PBX_E_OOM, "host allocation failed");
    s->dim = dim;
    const uint64_t per = (capacity_hint + (uint64_t)n_devices - 1) / (uint64_t)n_devices;
    for (int i = 0; i < n_devices; ++i) {
        std::unique_ptr<pbx_sharded::Shard> sh(new (std::nothrow) pbx_sharded::Shard());
        if (!sh) { pbx_sharded_destroy(s); return fail(PBX_E_OOM, "host allocation failed"); }
        sh->device = devices[i];
        int rc = pbx_corpus_create(dim, per, devices[i], &sh->c);
        if (rc == PBX_OK && cudaEventCreateWithFlags(&sh->ev_done, cudaEventDisableTiming) != cudaSuccess) rc = fail(PBX_E_CUDA, "event creation failed");
        s->shards.push_back(std::move(sh));
        if (rc != PBX_OK) { pbx_sharded_destroy(s); return rc; }
    }
    // peer access between distinct devices makes the record copies direct NVLink writes (without it the copy is staged
    // by the driver: slower, same result)
    for (auto& a : s->shards)
        for (auto& b : s->shards) {
            if (a->device == b->device) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, a->device, b->device) == cudaSuccess && can) {
                cudaSetDevice(a->device);
                cudaError_t e = cudaDeviceEnablePeerAccess(b->device, 0);
                if (e != cudaSuccess) cudaGetLastError();            // already enabled (or refused): not an error
            }
        }
    for (uint32_t i = 0; i < s->shards.size(); ++i) s->shards[i]->th = std::thread(sharded_worker, s, s->shards[i].get(), i);
    *out = s;
    return PBX_OK;
}

extern "C" void pbx_sharded_destroy(pbx_sharded* s) {
    if (!s) return;
    for (auto& sh : s->shards) {
        if (sh->th.joinable()) {
            { std::lock_guard<std::mutex> lk(sh->mu); sh->quit.store(true); }
            sh->cv.notify_all();
            sh->th.join();
        }
    }
    for (auto& sh : s->shards) {
        cudaSetDevice(sh->device);
        cudaDeviceSynchronize();
        cudaFree(sh->d_q); cudaFree(sh->d_hits); cudaFree(sh->d_cnt); cudaFreeHost(sh->h_cnt);
        if (sh->ev_done) cudaEventDestroy(sh->ev_done);
        if (sh->c) pbx_corpus_destroy(sh->c);
    }
    if (!s->shards.empty()) cudaSetDevice(s->shards[0]->device);
    cudaFree(s->d_gathered); cudaFree(s->d_out); cudaFreeHost(s->h_q); cudaFreeHost(s->h_out);
    cudaGetLastError();
    delete s;
}

extern "C" int pbx_sharded_shards(const pbx_sharded* s, uint32_t* n_shards) {
    if (!s || !n_shards) return fail(PBX_E_INVALID, "NULL argument");
    *n_shards = (uint32_t)s->shards.size();
    return PBX_OK;
}

extern "C" int pbx_sharded_shard(pbx_sharded* s, uint32_t index, pbx_corpus** out) {
    if (!s || !out || index >= s->shards.size()) return fail(PBX_E_INVALID, "bad shard index");
    *out = s->shards[index]->c;
    return PBX_OK;
}

extern "C" int pbx_sharded_size(const pbx_sharded* s, uint64_t* n_rows) {
    if (!s || !n_rows) return fail(PBX_E_INVALID, "NULL argument");
    uint64_t t = 0;
    for (auto& sh : s->shards) t += sh->c->n.load() + sh->c->pending.load();
    *n_rows = t;
    return PBX_OK;
}

// Contiguous blocks of the id-ordered table, one per shard (the partition of SURVEY.md 8e; ids stay global).
extern "C" int pbx_sharded_load(pbx_sharded* s, const int64_t* image_ids, const uint8_t* hashes, uint64_t n) {
    if (!s) return fail(PBX_E_INVALID, "corpus is NULL");
    if (n && (!image_ids || !hashes)) return fail(PBX_E_INVALID, "NULL ids or hashes with n > 0");
    std::lock_guard<std::mutex> lk(s->mu);
    const uint64_t w = s->shards.size(), base = n / w, rem = n % w;
    for (uint64_t i = 0; i < w; ++i) {
        const uint64_t first = i * base + std::min<uint64_t>(i, rem), cnt = base + (i < rem ? 1 : 0);
        int rc = pbx_corpus_load(s->shards[i]->c, image_ids + first, hashes + first * s->dim, cnt);
        if (rc != PBX_OK) return rc;
    }
    return PBX_OK;
}

// Appended rows go to the shard that holds the fewest (shards stay balanced as the indexer runs, src/engine.rs:186-203).
extern "C" int pbx_sharded_append(pbx_sharded* s, const int64_t* image_ids, const uint8_t* hashes, uint64_t n) {
    if (!s) return fail(PBX_E_INVALID, "corpus is NULL");
    if (n == 0) return PBX_OK;
    if (!image_ids || !hashes) return fail(PBX_E_INVALID, "NULL ids or hashes with n > 0");
    pbx_sharded::Shard* best = s->shards[0].get();
    for (auto& sh : s->shards)
        if (sh->c->n.load() + sh->c->pending.load() < best->c->n.load() + best->c->pending.load()) best = sh.get();
    return pbx_corpus_append(best->c, image_ids, hashes, n);
}

extern "C" int pbx_sharded_fill_synthetic(pbx_sharded* s, uint64_t rows_per_shard, uint64_t seed) {
    if (!s) return fail(PBX_E_INVALID, "corpus is NULL");
    std::lock_guard<std::mutex> lk(s->mu);
    for (uint64_t i = 0; i < s->shards.size(); ++i) {
        int rc = pbx_corpus_fill_synthetic(s->shards[i]->c, rows_per_shard, seed, i * rows_per_shard);
        if (rc != PBX_OK) return rc;
    }
    return PBX_OK;
}

extern "C" int pbx_sharded_search_hits(pbx_sharded* s, const uint8_t* queries, uint32_t nq, uint32_t k, double max_dist, pbx_hit* out_hits,
                                       uint32_t* out_count) {
    if (!s) return fail(PBX_E_INVALID, "corpus is NULL");
    int rc = check_search_args(s->shards[0]->c, queries, nq, k);
    if (rc != PBX_OK) return rc;
    if (nq == 0) return PBX_OK;
    if (!out_hits || !out_count) return fail(PBX_E_INVALID, "NULL output");
    std::lock_guard<std::mutex> lk(s->mu);
    const uint32_t n_sh = (uint32_t)s->shards.size();
    pbx_sharded::Shard* root = s->shards[0].get();
    const uint32_t batch_max = 1024;
    for (uint32_t q0 = 0; q0 < nq; q0 += batch_max) {
        const uint32_t b = std::min<uint32_t>(batch_max, nq - q0);
        rc = sharded_ensure_buffers(s, b, k);
        if (rc != PBX_OK) return rc;
        memcpy(s->h_q, queries + (size_t)q0 * s->dim, (size_t)b * s->dim);
        s->nq = b; s->k = k; s->max_dist = max_dist;
        for (auto& sh : s->shards) {
            { std::lock_guard<std::mutex> wl(sh->mu); sh->req.fetch_add(1, std::memory_order_release); }
            sh->cv.notify_one();
        }
        // every shard has enqueued its half (events recorded) before the root stream is told to wait for them
        for (auto& sh : s->shards) {
            const uint64_t want = sh->req.load(std::memory_order_acquire);
            while (sh->done.load(std::memory_order_acquire) != want) std::this_thread::yield();
        }
        for (auto& sh : s->shards)
            if (sh->rc != PBX_OK) {
                for (auto& o : s->shards) { cudaSetDevice(o->device); cudaStreamSynchronize(o->c->stream); }
                return fail(sh->rc, "%s", sh->err);
            }
        CU_TRY(cudaSetDevice(root->device));
        cudaStream_t st = root->c->stream;
        for (auto& sh : s->shards) CU_TRY(cudaStreamWaitEvent(st, sh->ev_done, 0));
        uint32_t* d_cnt_out = reinterpret_cast<uint32_t*>(s->d_out + (size_t)b * k);
        rc = pbx_merge_hits_device(root->device, s->d_gathered, nullptr, n_sh, b, k, s->d_out, d_cnt_out, st);
        if (rc != PBX_OK) return rc;
        const size_t out_bytes = (size_t)b * k * sizeof(pbx_hit) + (size_t)b * sizeof(uint32_t);
        CU_TRY(cudaMemcpyAsync(s->h_out, s->d_out, out_bytes, cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        for (auto& sh : s->shards) {                     // the other shards' count copies (error markers) have landed too
            if (sh.get() != root) { CU_TRY(cudaSetDevice(sh->device)); CU_TRY(cudaStreamSynchronize(sh->c->stream)); }
            for (uint32_t q = 0; q < b; ++q)
                if (sh->h_cnt[q] > k) return fail(PBX_E_INTERNAL, "shard on device %d could not certify query %u (count marker 0x%x)", sh->device, q0 + q, sh->h_cnt[q]);
        }
        memcpy(out_hits + (size_t)q0 * k, s->h_out, (size_t)b * k * sizeof(pbx_hit));
        memcpy(out_count + q0, reinterpret_cast<const uint32_t*>(s->h_out + (size_t)b * k), (size_t)b * sizeof(uint32_t));
    }
    return PBX_OK;
}

extern "C" int pbx_sharded_search(pbx_sharded* s, const uint8_t* queries, uint32_t nq, uint32_t k, double max_dist, int64_t* out_ids,
                                  float* out_dist, int32_t* out_dot, int32_t* out_norm2, uint32_t* out_count) {
    if (!s) return fail(PBX_E_INVALID, "corpus is NULL");
    if (nq == 0) return PBX_OK;
    if (!out_ids || !out_count) return fail(PBX_E_INVALID, "NULL output");
    if (k == 0 || k > PBX_MAX_K) return fail(k == 0 ? PBX_E_INVALID : PBX_E_K, "k = %u outside [1, %u]", k, PBX_MAX_K);
    std::vector<pbx_hit> hits;
    try { hits.resize((size_t)nq * k); } catch (...) { return fail(PBX_E_OOM, "host allocation failed"); }
    int rc = pbx_sharded_search_hits(s, queries, nq, k, max_dist, hits.data(), out_count);
    if (rc != PBX_OK) return rc;
    for (size_t i = 0; i < hits.size(); ++i) {
        const bool valid = (i % k) < out_count[i / k];
        out_ids[i] = valid ? hits[i].image_id : 0;
        if (out_dist) out_dist[i] = valid ? hits[i].dist : 0.f;
        if (out_dot) out_dot[i] = valid ? hits[i].dot : 0;
        if (out_norm2) out_norm2[i] = valid ? hits[i].norm2 : 0;
    }
    return PBX_OK;
}
