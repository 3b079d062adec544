// HOST code (no device work): the reference's per-pass minibatch order for a whole update, pipelined.
//
// Batch.split draws one np.random.permutation(N) per pass from numpy's global legacy RandomState
// (batch.py:1209); nothing else consumes that stream inside Algorithm.update(), so the `repeat` draws of one
// update can be produced ahead of the passes that use them.  The draw itself is sequential in two ways: the
// MT19937 stream (data-dependent length because random_interval rejects) and the Fisher-Yates swaps.  A job
// splits them: ONE producer thread walks the generator and records the accepted index j for every i
// (~1.5 ms per 524 288-element permutation, no memory traffic besides the record), and worker threads apply the
// swaps of different passes concurrently (~2.5 ms each) straight into the caller's (pinned) int32 rows.
// Pass r becomes available ~4 + 1.5 r ms after the start, which keeps pace with the GPU's ~1.5 ms per pass.
// Bit-identical to np.random.permutation, including the final generator state (ts_host_perm_job_finish).
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
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

struct PermJob {
    uint32_t key[kMtN];
    int pos = 0;
    int64_t n = 0;
    int repeat = 0;
    int32_t* out = nullptr;                      // [repeat][n]
    std::vector<std::vector<uint32_t>> js;       // js[r][i] = accepted j for position i (i >= 1)
    std::vector<int> state;                      // 0 = pending, 1 = j-sequence ready, 2 = permutation ready
    std::mutex mu;
    std::condition_variable cv;
    std::thread producer;
    std::vector<std::thread> workers;
    std::atomic<int> next_apply{0};
    int applied = 0;                             // passes completely applied (guarded by mu)
    static constexpr int kAhead = 6;             // the producer stays at most this many passes ahead of the workers

    void produce() {
        for (int r = 0; r < repeat; ++r) {
            {   // back-pressure: bounded memory (4 n bytes per pass in flight) whatever `repeat` is
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return r - applied < kAhead; });
            }
            std::vector<uint32_t>& j = js[r];
            j.resize((size_t)(n > 0 ? n : 1));
            int p = pos;
            // One iteration per DRAW (not per position): write the candidate, step to the next position only when it
            // is accepted (v <= i).  No data-dependent branch -- random_interval's rejection loop mispredicts ~30 % of
            // the time when written as do/while.  The generator block is tempered in a separate (vectorisable) loop.
            uint32_t* jd = j.data();
            uint32_t tmp[kMtN];
            int64_t i = n - 1;
            while (i >= 1) {
                if (p == kMtN) { mt_gen(key); p = 0; }
                const int avail = kMtN - p;
                for (int d = 0; d < avail; ++d) {
                    uint32_t y = key[p + d];
                    y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
                    tmp[d] = y;
                }
                int d = 0;
                while (d < avail && i >= 1) {
                    // positions i in (lower, mask] share one mask: the loop-carried chain is just compare + subtract
                    const uint32_t mask = 0xffffffffu >> __builtin_clz((uint32_t)i);
                    const int64_t lower = (int64_t)(mask >> 1);
                    for (; d < avail && i > lower; ++d) {
                        const uint32_t v = tmp[d] & mask;
                        jd[i] = v;
                        i -= (int64_t)(v <= (uint32_t)i);
                    }
                }
                p += d;
            }
            pos = p;
            { std::lock_guard<std::mutex> lk(mu); state[r] = 1; }
            cv.notify_all();
        }
    }
    void apply_loop() {
        for (;;) {
            const int r = next_apply.fetch_add(1);
            if (r >= repeat) return;
            { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return state[r] >= 1; }); }
            int32_t* o = out + (int64_t)r * n;
            for (int64_t i = 0; i < n; ++i) o[i] = (int32_t)i;
            const std::vector<uint32_t>& j = js[r];
            for (int64_t i = n - 1; i >= 1; --i) { const uint32_t v = j[(size_t)i]; const int32_t t = o[v]; o[v] = o[i]; o[i] = t; }
            std::vector<uint32_t>().swap(js[r]);
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
        const int nw = n_workers < 1 ? 1 : (n_workers > repeat ? repeat : n_workers);
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
            if (job->producer.joinable()) job->producer.join();
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
    for (auto& w : job->workers) w.join();
    std::memcpy(key_out, job->key, sizeof(job->key));
    *pos_out = job->pos;
    delete job;
    return 0;
}
