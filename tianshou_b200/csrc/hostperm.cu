// HOST code (no device work): the reference's per-pass minibatch order for a whole update, pipelined.
//
// Batch.split draws one np.random.permutation(N) per pass from numpy's global legacy RandomState
// (batch.py:1209); nothing else consumes that stream inside Algorithm.update(), so the `repeat` draws of one
// update can be produced ahead of the passes that use them.  The draw itself is sequential in two ways: the
// MT19937 stream (data-dependent length because random_interval rejects) and the Fisher-Yates swaps.  A job
// splits them three ways: a GENERATOR thread runs MT19937 ahead (the word stream does not depend on how many words a
// permutation consumes: mt_gen + vectorised tempering into a ring of blocks), ONE walker thread does the data-dependent
// part only (mask, compare, advance: records the accepted index j for every i), and worker threads apply the
// swaps of different passes concurrently (~2.5 ms each) straight into the caller's (pinned) int32 rows.
// Bit-identical to np.random.permutation, including the final generator state (ts_host_perm_job_finish).
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {
constexpr int kMtN = 624, kMtM = 397;
inline void mt_gen(uint32_t* key) {
    constexpr uint32_t A = 0x9908b0dfu, UP = 0x80000000u, LO = 0x7fffffffu;
    int i = 0;
    uint32_t y;
    for (; i < kMtN - kMtM; ++i) { y = (key[i] & UP) | (key[i + 1] & LO); key[i] = key[i + kMtM] ^ (y >> 1) ^ ((0u - (y & 1u)) & A); }
    for (; i < kMtN - 1; ++i) { y = (key[i] & UP) | (key[i + 1] & LO); key[i] = key[i + (kMtM - kMtN)] ^ (y >> 1) ^ ((0u - (y & 1u)) & A); }
    y = (key[kMtN - 1] & UP) | (key[0] & LO);
    key[kMtN - 1] = key[kMtM - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
}

// One generator block: the MT19937 state after its mt_gen (numpy's `key`) and the tempered outputs of its 624 words.
struct MtBlock {
    uint32_t key[kMtN];
    uint32_t temp[kMtN];
};
__attribute__((target_clones("avx2", "default"))) void mt_temper(const uint32_t* key, uint32_t* out) {
    for (int d = 0; d < kMtN; ++d) {        // vectorisable: no loop-carried dependence
        uint32_t y = key[d];
        y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
        out[d] = y;
    }
}

struct PermJob {
    uint32_t key[kMtN];
    int pos = 0;
    int64_t n = 0;
    int repeat = 0;
    int32_t* out = nullptr;                      // [repeat][n]
    std::vector<std::unique_ptr<uint32_t[]>> js; // js[r][i] = accepted j for position i (i >= 1); uninitialised storage
    std::vector<int> state;                      // 0 = pending, 1 = walk started (j storage exists), 2 = permutation ready
    std::unique_ptr<std::atomic<int64_t>[]> progress;   // per pass: every position ABOVE this one has its final j (streamed to the applier)
    std::mutex mu;
    std::condition_variable cv;
    std::thread generator, producer;
    std::vector<std::thread> workers;
    std::atomic<int> next_apply{0};
    int applied = 0;                             // passes completely applied (guarded by mu)
    static constexpr int kAhead = 6;             // the walker stays at most this many passes ahead of the workers
    // generator -> walker ring.  The MT19937 word stream does not depend on how many words a permutation consumes, so a
    // generator thread runs ahead (mt_gen + tempering, vectorised) while the walker does only the data-dependent part
    // (mask, compare, advance): ~0.7 ms per 524 288-element permutation instead of ~1.7 ms in one thread.
    static constexpr int64_t kRing = 256;        // blocks (5 KB each): the generator may run ~160 k words ahead
    std::unique_ptr<MtBlock[]> ring;           // uninitialised storage (no 1.3 MB memset per job)
    std::atomic<int64_t> produced{0}, consumed{0};
    std::atomic<bool> stop{false};

    void generate() {
        uint32_t k[kMtN];
        std::memcpy(k, key, sizeof(k));
        // block 0 = the caller's state as it stands (its words [pos, 624) are unconsumed); block b > 0 = mt_gen of block b - 1
        for (int64_t b = 0;; ++b) {
            while (b - consumed.load(std::memory_order_acquire) >= kRing) {
                if (stop.load(std::memory_order_acquire)) return;
                std::this_thread::yield();
            }
            if (stop.load(std::memory_order_acquire)) return;
            if (b > 0) mt_gen(k);
            MtBlock& blk = ring[(size_t)(b % kRing)];
            std::memcpy(blk.key, k, sizeof(k));
            mt_temper(k, blk.temp);
            produced.store(b + 1, std::memory_order_release);
        }
    }
    const MtBlock& block(int64_t b) {
        while (produced.load(std::memory_order_acquire) <= b) std::this_thread::yield();
        return ring[(size_t)(b % kRing)];
    }

    void produce() {
        int64_t b = 0;                 // current block
        int p = pos;                   // next unconsumed word of it
        const MtBlock* blk = &block(0);
        for (int r = 0; r < repeat; ++r) {
            {   // back-pressure: bounded memory (4 n bytes per pass in flight) whatever `repeat` is
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return r - applied < kAhead; });
            }
            js[r].reset(new uint32_t[(size_t)(n > 0 ? n : 1)]);
            progress[r].store(n - 1, std::memory_order_relaxed);
            { std::lock_guard<std::mutex> lk(mu); state[r] = 1; }     // the applier may start: it follows `progress`
            cv.notify_all();
            // One iteration per DRAW (not per position): write the candidate, step to the next position only when it
            // is accepted (v <= i).  No data-dependent branch -- random_interval's rejection loop mispredicts ~30 % of
            // the time when written as do/while.  Four draws per bounds check: i drops by at most one per draw.
            uint32_t* jd = js[r].get();
            int64_t i = n - 1;
            while (i >= 1) {
                if (p == kMtN) {
                    consumed.store(b + 1, std::memory_order_release);
                    progress[r].store(i, std::memory_order_release);
                    ++b; p = 0;
                    blk = &block(b);
                }
                const uint32_t* tmp = blk->temp + p;
                const int avail = kMtN - p;
                int d = 0;
                while (d < avail && i >= 1) {
                    // positions i in (lower, mask] share one mask: the loop-carried chain is just compare + subtract
                    const uint32_t mask = 0xffffffffu >> __builtin_clz((uint32_t)i);
                    const int64_t lower = (int64_t)(mask >> 1);
                    for (; d + 4 <= avail && i - 4 > lower; d += 4) {
                        const uint32_t v0 = tmp[d] & mask;     jd[i] = v0; i -= (int64_t)(v0 <= (uint32_t)i);
                        const uint32_t v1 = tmp[d + 1] & mask; jd[i] = v1; i -= (int64_t)(v1 <= (uint32_t)i);
                        const uint32_t v2 = tmp[d + 2] & mask; jd[i] = v2; i -= (int64_t)(v2 <= (uint32_t)i);
                        const uint32_t v3 = tmp[d + 3] & mask; jd[i] = v3; i -= (int64_t)(v3 <= (uint32_t)i);
                    }
                    for (; d < avail && i > lower; ++d) {
                        const uint32_t v = tmp[d] & mask;
                        jd[i] = v;
                        i -= (int64_t)(v <= (uint32_t)i);
                    }
                }
                p += d;
            }
            progress[r].store(0, std::memory_order_release);
        }
        // final generator state = numpy's (key, pos) after these draws: the current block's key, next unconsumed word
        std::memcpy(key, blk->key, sizeof(key));
        pos = p;
        stop.store(true, std::memory_order_release);
    }
    void apply_loop() {
        for (;;) {
            const int r = next_apply.fetch_add(1);
            if (r >= repeat) return;
            { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return state[r] >= 1; }); }
            int32_t* o = out + (int64_t)r * n;
            for (int64_t i = 0; i < n; ++i) o[i] = (int32_t)i;
            const uint32_t* j = js[r].get();
            // The swaps follow the walker as it goes (positions above progress[r] are final): the first pass of an update is
            // ready ~max(walk, apply) after the start instead of walk + apply.  The swap partners are known in advance:
            // prefetch them (the 4 n-byte row does not fit a core's L1; a dependent miss per swap is what this loop would wait on).
            constexpr int64_t kPf = 24;
            int64_t safe = progress[r].load(std::memory_order_acquire);       // positions > safe may be applied
            for (int64_t i = n - 1; i >= 1; --i) {
                while (i <= safe && safe > 0) {
                    std::this_thread::yield();
                    safe = progress[r].load(std::memory_order_acquire);
                }
                if (i - kPf > safe) __builtin_prefetch(o + j[(size_t)(i - kPf)], 1, 1);
                const uint32_t v = j[(size_t)i]; const int32_t t = o[v]; o[v] = o[i]; o[i] = t;
            }
            js[r].reset();
            { std::lock_guard<std::mutex> lk(mu); state[r] = 2; ++applied; }
            cv.notify_all();
        }
    }
};
}  // namespace

extern "C" int ts_host_perm_job_start(const uint32_t* key, int32_t pos, int64_t n, int32_t repeat, int32_t* out,
                                      int32_t n_workers, void** job_out) {
    TS_REQUIRE(key && out && job_out && n >= 0 && n <= 0x7fffffffLL && repeat >= 1 && pos >= 0 && pos <= kMtN,
               "ts_host_perm_job_start: bad arguments");
    PermJob* job = nullptr;
    try {
        job = new PermJob();
        std::memcpy(job->key, key, sizeof(job->key));
        job->pos = pos; job->n = n; job->repeat = repeat; job->out = out;
        job->js.resize((size_t)repeat);
        job->state.assign((size_t)repeat, 0);
        job->progress.reset(new std::atomic<int64_t>[(size_t)repeat]);
        for (int r = 0; r < repeat; ++r) job->progress[r].store(n, std::memory_order_relaxed);
        job->ring.reset(new MtBlock[(size_t)PermJob::kRing]);
        const int nw = n_workers < 1 ? 1 : (n_workers > repeat ? repeat : n_workers);
        job->generator = std::thread([job] { job->generate(); });
        job->producer = std::thread([job] { job->produce(); });
        for (int w = 0; w < nw; ++w) {
            try {
                job->workers.emplace_back([job] { job->apply_loop(); });
            } catch (const std::exception&) {
                if (job->workers.empty()) throw;      // no worker at all: give up; otherwise run with fewer
                break;
            }
        }
    } catch (const std::exception& e) {     // out of memory / thread limit: the caller falls back to the serial draw
        if (job) {
            // unblock everything before joining: the walker may be waiting on back-pressure with no worker to relieve it
            { std::lock_guard<std::mutex> lk(job->mu); job->applied = 1 << 30; }
            job->cv.notify_all();
            if (job->producer.joinable()) job->producer.join();
            job->stop.store(true, std::memory_order_release);
            if (job->generator.joinable()) job->generator.join();
            for (auto& w : job->workers) w.join();
            delete job;
        }
        tsb::set_error("ts_host_perm_job_start: %s", e.what());
        return 1;
    }
    *job_out = job;
    return 0;
}

extern "C" int ts_host_perm_job_wait(void* handle, int32_t r) {
    PermJob* job = static_cast<PermJob*>(handle);
    TS_REQUIRE(job && r >= 0 && r < job->repeat, "ts_host_perm_job_wait: bad arguments");
    std::unique_lock<std::mutex> lk(job->mu);
    job->cv.wait(lk, [&] { return job->state[(size_t)r] == 2; });
    return 0;
}

extern "C" int ts_host_perm_job_finish(void* handle, uint32_t* key_out, int32_t* pos_out) {
    PermJob* job = static_cast<PermJob*>(handle);
    TS_REQUIRE(job && key_out && pos_out, "ts_host_perm_job_finish: bad arguments");
    job->producer.join();
    job->generator.join();
    for (auto& w : job->workers) w.join();
    std::memcpy(key_out, job->key, sizeof(job->key));
    *pos_out = job->pos;
    delete job;
    return 0;
}
