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
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include <pthread.h>
#include <sched.h>
#include <sys/resource.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "common.cuh"
#include "hostperm_simd.h"

namespace {
using tsb_hp::kMtN;

// One generator block: the MT19937 state after its transition (numpy's `key`) and the tempered outputs of its 624 words.
struct MtBlock {
    uint32_t key[kMtN];
    uint32_t temp[kMtN];
};

// Threads parked between jobs: creating the six threads of a job costs ~0.3 ms at the top of every update() (measured on
// the GPU box, profiles/r2g_default_order_hosttrace.txt), waking parked ones a few tens of microseconds.  A task gets a
// parked thread if one is free, else a new thread (tasks of one job depend on each other: none may queue behind another).
class Crew {
  public:
    static Crew& get() {
        static std::once_flag once;
        std::call_once(once, [] { pthread_atfork(nullptr, nullptr, [] { instance() = nullptr; }); });
        Crew*& c = instance();
        if (c == nullptr) c = new Crew();        // never destroyed: parked threads may outlive static destructors
        return *c;
    }
    void run(std::function<void()> f) {
        std::unique_lock<std::mutex> lk(mu_);
        q_.push_back(std::move(f));
        if ((size_t)idle_ >= q_.size()) { cv_.notify_one(); return; }
        try {
            std::thread([this] { loop(); }).detach();
        } catch (...) {
            q_.pop_back();
            throw;
        }
    }
  private:
    static Crew*& instance() { static Crew* c = nullptr; return c; }
    void loop() {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> lk(mu_);
                ++idle_;
                cv_.wait(lk, [&] { return !q_.empty(); });
                --idle_;
                f = std::move(q_.front());
                q_.pop_front();
            }
            f();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> q_;
    int idle_ = 0;
};

// NUMA placement: the rows live in pinned host memory allocated by the calling thread; a job whose appliers land on the other
// socket does its ~5 M dependent random swaps over the inter-socket link (the same job took 6.2 .. 10.8 ms on different
// visits of a 2-socket box).  The crew threads of a job are confined to the CPUs of the caller's NUMA node (intersected with
// the process affinity; no pinning if that leaves fewer than 8 CPUs, on a single-node machine, or with TS_B200_PERM_PIN=0).
struct NodeCpus {
    bool valid = false;
    cpu_set_t set;
};
inline bool parse_cpulist(const char* s, cpu_set_t* out) {
    CPU_ZERO(out);
    bool any = false;
    while (*s) {
        char* end = nullptr;
        const long a = std::strtol(s, &end, 10);
        if (end == s) break;
        long b = a;
        s = end;
        if (*s == '-') { b = std::strtol(s + 1, &end, 10); s = end; }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET((int)c, out); any = true; }
        if (*s == ',') ++s;
    }
    return any;
}
// The machine's NUMA nodes, read from sysfs ONCE per process (a job is started at the top of every update()).
inline const std::vector<cpu_set_t>& numa_nodes() {
    static std::vector<cpu_set_t> nodes;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* base = std::getenv("TS_B200_SYSFS_NODE_DIR");          // tests point this at a fake topology
        if (base == nullptr || std::strlen(base) > 160) base = "/sys/devices/system/node";
        for (int node = 0; node < 64; ++node) {
            char path[256];
            std::snprintf(path, sizeof(path), "%s/node%d/cpulist", base, node);
            FILE* f = std::fopen(path, "r");
            if (f == nullptr) continue;
            char buf[4096];
            cpu_set_t cpus;
            if (std::fgets(buf, sizeof(buf), f) != nullptr && parse_cpulist(buf, &cpus)) nodes.push_back(cpus);
            std::fclose(f);
        }
    });
    return nodes;
}
inline NodeCpus caller_node_cpus() {
    NodeCpus r;
    const char* env = std::getenv("TS_B200_PERM_PIN");
    if (env != nullptr && env[0] == '0') return r;
    const std::vector<cpu_set_t>& nodes = numa_nodes();
    if (nodes.size() < 2) return r;           // one node: nothing to choose
    const int cpu = sched_getcpu();
    cpu_set_t allowed;
    if (cpu < 0 || sched_getaffinity(0, sizeof(allowed), &allowed) != 0) return r;
    for (const cpu_set_t& node : nodes) {
        if (!CPU_ISSET(cpu, &node)) continue;
        cpu_set_t cpus;
        CPU_AND(&cpus, &node, &allowed);
        const char* min_env = std::getenv("TS_B200_PERM_PIN_MIN_CPUS");
        const int min_cpus = min_env != nullptr ? std::atoi(min_env) : 8;
        if (CPU_COUNT(&cpus) >= min_cpus) { r.valid = true; r.set = cpus; }
        break;
    }
    return r;
}

struct PermJob {
    uint32_t key[kMtN];
    int pos = 0;
    int isa = tsb_hp::detect_isa();
    int64_t n = 0;
    int repeat = 0;
    int32_t* out = nullptr;                      // [repeat][n]
    // partner lists in walk order (entry k of a pass = the accepted j of position n - 1 - k): kAhead slots of `stride` words, pass r
    // uses slot r % kAhead.  The storage (and the generator ring) comes from a process-wide cache: a fresh 2 MB-per-slot
    // allocation costs its page faults on the first pass of every update() otherwise.
    std::unique_ptr<uint32_t[]> js_store;
    size_t js_words = 0, stride = 0;
    uint32_t* js(int r) { return js_store.get() + (size_t)(r % kAhead) * stride; }
    std::vector<int> state;                      // 0 = pending, 1 = walk started (j storage exists), 2 = permutation ready
    std::unique_ptr<std::atomic<int64_t>[]> progress;   // per pass: every position ABOVE this one has its final j (streamed to the applier)
    std::mutex mu;
    std::condition_variable cv;
    NodeCpus node = caller_node_cpus();          // where the caller (and its pinned rows) live
    // TS_B200_PERM_NICE=<n>: run the crew at a lower priority than the thread that feeds the GPU (several ranks per box)
    // (default: 10 when torchrun started >= 4 local ranks -- measured at N = 4, profiles/r2_ab_n4.txt -- else 0)
    int nice_value = [] {
        if (const char* e = std::getenv("TS_B200_PERM_NICE")) return std::atoi(e);
        const char* w = std::getenv("LOCAL_WORLD_SIZE");
        return (w != nullptr && std::atoi(w) >= 4) ? 10 : 0;
    }();
    int live = 0;                                // tasks of this job still running on crew threads (guarded by mu)
    int n_appliers = 0;
    std::atomic<int> next_apply{0};
    int applied = 0;                             // passes completely applied (guarded by mu)
    bool abort_job = false;                      // error path of ts_host_perm_job_start (guarded by mu)
    static constexpr int kAhead = 6;             // the walker stays at most this many passes ahead of the workers
    // generator -> walker ring.  The MT19937 word stream does not depend on how many words a permutation consumes, so a
    // generator thread runs ahead (mt_gen + tempering, vectorised) while the walker does only the data-dependent part
    // (mask, compare, advance): ~0.7 ms per 524 288-element permutation instead of ~1.7 ms in one thread.
    static constexpr int64_t kRing = 256;        // blocks (5 KB each): the generator may run ~160 k words ahead
    std::unique_ptr<MtBlock[]> ring;           // uninitialised storage (no 1.3 MB memset per job)
    std::atomic<int64_t> produced{0}, consumed{0};
    std::atomic<bool> stop{false};
    // TS_B200_PERM_TRACE=1: per-pass stage times (ms since the job started), printed by ts_host_perm_job_finish
    bool trace = std::getenv("TS_B200_PERM_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    std::vector<double> t_walk0, t_walk1, t_app0, t_app1;
    double now_ms() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }

    // Run a member task on a crew thread; `live` counts the tasks still running (the last decrement happens under `mu`, so
    // the finisher, who waits under the same mutex, cannot free the job while a task still touches it).
    template <class F> void launch(F f) {
        { std::lock_guard<std::mutex> lk(mu); ++live; }
        try {
            Crew::get().run([this, f] {
                if (node.valid) sched_setaffinity(0, sizeof(node.set), &node.set);      // this thread only
                if (nice_value != 0) setpriority(PRIO_PROCESS, (id_t)syscall(SYS_gettid), nice_value);   // per-thread on Linux
                f();
                std::lock_guard<std::mutex> lk(mu);
                --live;
                cv.notify_all();
            });
        } catch (...) {
            std::lock_guard<std::mutex> lk(mu);
            --live;
            throw;
        }
    }
    void wait_all_tasks() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return live == 0; });
    }
    void generate() {
        // block 0 = the caller's state as it stands (its words [pos, 624) are unconsumed); block b > 0 = the transition of block b - 1
        for (int64_t b = 0;; ++b) {
            while (b - consumed.load(std::memory_order_acquire) >= kRing) {
                if (stop.load(std::memory_order_acquire)) return;
                std::this_thread::yield();
            }
            if (stop.load(std::memory_order_acquire)) return;
            MtBlock& blk = ring[(size_t)(b % kRing)];
            if (b == 0) std::memcpy(blk.key, key, sizeof(key));
            else tsb_hp::mt_next(ring[(size_t)((b - 1) % kRing)].key, blk.key);
            tsb_hp::mt_temper(blk.key, blk.temp);
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
            {   // back-pressure: bounded memory (4 n bytes per pass in flight) whatever `repeat` is -- the slot's previous pass is applied
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return r < kAhead || state[(size_t)(r - kAhead)] == 2 || abort_job; });
                if (abort_job) break;
            }
            progress[r].store(n - 1, std::memory_order_relaxed);
            { std::lock_guard<std::mutex> lk(mu); state[r] = 1; }     // the applier may start: it follows `progress`
            cv.notify_all();
            // The partner list is in walk order (entry k = position n - 1 - k); tsb_hp::walk decides whole groups of draws
            // with vector compares (hostperm_simd.cpp).
            if (trace) t_walk0[(size_t)r] = now_ms();
            tsb_hp::Walk w{n - 1, js(r)};
            while (w.i >= 1) {
                if (p == kMtN) {
                    consumed.store(b + 1, std::memory_order_release);
                    progress[r].store(w.i, std::memory_order_release);
                    ++b; p = 0;
                    blk = &block(b);
                }
                p += tsb_hp::walk(isa, blk->temp + p, kMtN - p, &w);
            }
            progress[r].store(0, std::memory_order_release);
            if (trace) t_walk1[(size_t)r] = now_ms();
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
            if (trace) t_app0[(size_t)r] = now_ms();
            for (int64_t i = 0; i < n; ++i) o[i] = (int32_t)i;
            const uint32_t* j = js(r);
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
                if (i - kPf > safe) __builtin_prefetch(o + j[(size_t)(n - 1 - (i - kPf))], 1, 1);
                const uint32_t v = j[(size_t)(n - 1 - i)]; const int32_t t = o[v]; o[v] = o[i]; o[i] = t;
            }
            if (trace) t_app1[(size_t)r] = now_ms();
            { std::lock_guard<std::mutex> lk(mu); state[r] = 2; ++applied; }
            cv.notify_all();
        }
    }
};
// Storage of the last finished job, kept for the next one (one update() at a time per process is the normal case).
struct Cache {
    std::mutex mu;
    std::unique_ptr<uint32_t[]> js_store;
    size_t js_words = 0;
    std::unique_ptr<MtBlock[]> ring;
};
Cache& cache() { static Cache c; return c; }
void take_cached(PermJob* job) {
    Cache& c = cache();
    const size_t need = job->stride * (size_t)PermJob::kAhead;
    {
        std::lock_guard<std::mutex> lk(c.mu);
        if (c.js_words >= need) { job->js_store = std::move(c.js_store); job->js_words = c.js_words; c.js_words = 0; }
        job->ring = std::move(c.ring);
    }
    if (!job->js_store) { job->js_store.reset(new uint32_t[need]); job->js_words = need; }
    if (!job->ring) job->ring.reset(new MtBlock[(size_t)PermJob::kRing]);
}
void give_back(PermJob* job) {
    Cache& c = cache();
    std::lock_guard<std::mutex> lk(c.mu);
    if (job->js_words > c.js_words) { c.js_store = std::move(job->js_store); c.js_words = job->js_words; }
    if (!c.ring) c.ring = std::move(job->ring);
}
}  // namespace

extern "C" int ts_host_perm_job_start(const uint32_t* key, int32_t pos, int64_t n, int32_t repeat, int32_t* out,
                                      int32_t n_workers, void** job_out) {
    TS_REQUIRE(key && out && job_out && n >= 0 && n <= 0x7fffffffLL && repeat >= 1 && pos >= 0 && pos <= kMtN,
               "ts_host_perm_job_start: bad arguments");
    PermJob* job = nullptr;
    bool walker_launched = false;
    try {
        job = new PermJob();
        std::memcpy(job->key, key, sizeof(job->key));
        job->pos = pos; job->n = n; job->repeat = repeat; job->out = out;
        job->stride = (size_t)(n > 0 ? n : 1) + tsb_hp::kWalkSlack;
        take_cached(job);
        job->state.assign((size_t)repeat, 0);
        for (auto* v : {&job->t_walk0, &job->t_walk1, &job->t_app0, &job->t_app1}) v->assign((size_t)repeat, 0.0);
        job->progress.reset(new std::atomic<int64_t>[(size_t)repeat]);
        for (int r = 0; r < repeat; ++r) job->progress[r].store(n, std::memory_order_relaxed);
        const int nw = n_workers < 1 ? 1 : (n_workers > repeat ? repeat : n_workers);
        job->launch([job] { job->generate(); });
        job->launch([job] { job->produce(); });
        walker_launched = true;
        for (int w = 0; w < nw; ++w) {
            try {
                job->launch([job] { job->apply_loop(); });
                ++job->n_appliers;
            } catch (const std::exception&) {
                if (job->n_appliers == 0) throw;      // no applier at all: give up; otherwise run with fewer
                break;
            }
        }
    } catch (const std::exception& e) {     // out of memory / thread limit: the caller falls back to the serial draw
        if (job) {
            // unblock everything before joining: the walker may be waiting on back-pressure with no worker to relieve it
            { std::lock_guard<std::mutex> lk(job->mu); job->abort_job = true; }
            job->cv.notify_all();
            // the walker leaves through abort_job and then stops the generator itself (it may still need words until then)
            if (!walker_launched) job->stop.store(true, std::memory_order_release);
            job->wait_all_tasks();
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

// ---- asynchronous feed of the job's rows to the device -------------------------------------------------------------
// The host never has to stand between the walker / appliers and the GPU: for every pass a host function on an internal copy
// stream blocks THAT STREAM until the row is complete, the row is copied to the device and an event is recorded; the update
// (ts_ppo_update with `row_feed`) waits for event r on its own stream before pass r.  All passes of an update are then
// enqueued by ONE asynchronous C call, exactly as with a device-generated order.
namespace {
struct RowFeed {
    struct Arg { PermJob* job; int r; };
    std::vector<Arg> args;
    std::vector<cudaEvent_t> ev;
    cudaStream_t copy_stream = nullptr;
};
void CUDART_CB feed_wait_row(void* p) {      // runs on a driver thread: no CUDA calls in here
    auto* a = static_cast<RowFeed::Arg*>(p);
    std::unique_lock<std::mutex> lk(a->job->mu);
    a->job->cv.wait(lk, [&] { return a->job->state[(size_t)a->r] == 2; });
}
cudaStream_t feed_stream() {
    static cudaStream_t streams[tsb::kMaxDevices] = {};
    static std::mutex mu;
    const int dev = tsb::device_ordinal();
    std::lock_guard<std::mutex> lk(mu);
    if (streams[dev] == nullptr && cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    return streams[dev];
}
}  // namespace

extern "C" int ts_host_perm_feed_start(void* handle, const int32_t* host_rows, int32_t* dev_rows, int64_t n, int32_t repeat,
                                       void** feed_out) {
    PermJob* job = static_cast<PermJob*>(handle);
    TS_REQUIRE(job && host_rows && dev_rows && feed_out && n == job->n && repeat >= 1 && repeat <= job->repeat,
               "ts_host_perm_feed_start: bad arguments");
    cudaStream_t cs = feed_stream();
    TS_REQUIRE(cs != nullptr, "ts_host_perm_feed_start: cannot create the copy stream");
    auto* feed = new RowFeed();
    feed->copy_stream = cs;
    feed->args.resize((size_t)repeat);
    feed->ev.assign((size_t)repeat, nullptr);
    auto fail = [&](const char* what, cudaError_t e) {
        tsb::set_error("ts_host_perm_feed_start: %s: %s", what, cudaGetErrorString(e));
        cudaStreamSynchronize(cs);          // host functions already enqueued reference feed->args
        for (cudaEvent_t ev : feed->ev) if (ev) cudaEventDestroy(ev);
        delete feed;
        return 1;
    };
    for (int r = 0; r < repeat; ++r) {
        feed->args[(size_t)r] = {job, r};
        cudaError_t e = cudaEventCreateWithFlags(&feed->ev[(size_t)r], cudaEventDisableTiming);
        if (e != cudaSuccess) return fail("cudaEventCreate", e);
        if ((e = cudaLaunchHostFunc(cs, feed_wait_row, &feed->args[(size_t)r])) != cudaSuccess) return fail("cudaLaunchHostFunc", e);
        if ((e = cudaMemcpyAsync(dev_rows + (int64_t)r * n, host_rows + (int64_t)r * n, (size_t)n * sizeof(int32_t),
                                 cudaMemcpyHostToDevice, cs)) != cudaSuccess) return fail("cudaMemcpyAsync", e);
        if ((e = cudaEventRecord(feed->ev[(size_t)r], cs)) != cudaSuccess) return fail("cudaEventRecord", e);
    }
    *feed_out = feed;
    return 0;
}

extern "C" int ts_host_perm_feed_wait_row(void* handle, int32_t r, ts_stream_t stream) {
    auto* feed = static_cast<RowFeed*>(handle);
    TS_REQUIRE(feed && r >= 0 && r < (int32_t)feed->ev.size(), "ts_host_perm_feed_wait_row: bad arguments");
    TS_CUDA(cudaStreamWaitEvent(tsb::as_stream(stream), feed->ev[(size_t)r], 0));
    return 0;
}

extern "C" int ts_host_perm_feed_finish(void* handle) {
    auto* feed = static_cast<RowFeed*>(handle);
    TS_REQUIRE(feed, "ts_host_perm_feed_finish: bad arguments");
    const cudaError_t e = cudaStreamSynchronize(feed->copy_stream);      // every host function has returned: `args` may go
    for (cudaEvent_t ev : feed->ev) if (ev) cudaEventDestroy(ev);
    delete feed;
    if (e != cudaSuccess) { tsb::set_error("ts_host_perm_feed_finish: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}

extern "C" int ts_host_perm_job_finish(void* handle, uint32_t* key_out, int32_t* pos_out) {
    PermJob* job = static_cast<PermJob*>(handle);
    TS_REQUIRE(job && key_out && pos_out, "ts_host_perm_job_finish: bad arguments");
    job->wait_all_tasks();
    std::memcpy(key_out, job->key, sizeof(job->key));
    *pos_out = job->pos;
    if (job->trace) {
        fprintf(stderr, "[ts_host_perm] n=%lld repeat=%d isa=%d workers=%zu numa_pinned=%d (%d cpus)  (ms since start: walk begin-end | apply begin-end)\n",
                (long long)job->n, job->repeat, job->isa, (size_t)job->n_appliers, (int)job->node.valid,
                job->node.valid ? CPU_COUNT(&job->node.set) : 0);
        for (int r = 0; r < job->repeat; ++r)
            fprintf(stderr, "[ts_host_perm]   pass %2d  walk %7.3f-%7.3f  apply %7.3f-%7.3f\n", r, job->t_walk0[(size_t)r],
                    job->t_walk1[(size_t)r], job->t_app0[(size_t)r], job->t_app1[(size_t)r]);
    }
    give_back(job);
    delete job;
    return 0;
}
