// Host <-> device copies of PAGEABLE host memory for the numba_* entry points.
//
// A NumPy array handed over by @njit code is ordinary pageable memory.  cudaMemcpyAsync on such memory is staged by the
// driver through a small pinned buffer, one direction at a time and synchronously for the calling thread: measured on the
// B200 box, rfft2 of four 16384^2 images through numba_r2c reached 18 % of the throughput of the same call on pinned
// buffers (profiles/r02i_bench_1gpu.json).  This engine does the staging itself:
//   * a ring of pinned buffers per direction; a pool of host threads copies user memory <-> ring in parallel (a single
//     memcpy stream cannot feed a PCIe 5 x16 link), the DMA of one ring slot overlaps the host copy of the next;
//   * downloads are unstaged by a background thread, so that the caller goes on uploading the next chunk: both directions
//     of the link stay busy, as with pinned memory.
// (The reference has no counterpart: its arrays never leave the host.)
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "engine.h"

namespace rfb {

namespace {

class CopyPool {
  public:
    explicit CopyPool(int n) {
        for (int i = 0; i < n; ++i) workers_.emplace_back([this] { run(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    int size() const { return (int)workers_.size(); }
    // memcpy split over the workers; returns when every part is done (several callers may use the pool at once)
    void copy(char *dst, const char *src, size_t n) {
        const size_t parts = std::min<size_t>((size_t)size(), std::max<size_t>(1, n / (1u << 20)));
        if (parts <= 1) { memcpy(dst, src, n); return; }
        std::atomic<size_t> left{parts};
        std::mutex dmu;
        std::condition_variable dcv;
        const size_t per = ((n / parts) + 63) & ~(size_t)63;
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (size_t i = 0; i < parts; ++i) {
                const size_t off = i * per, len = i + 1 == parts ? n - off : per;
                q_.push_back([=, &left, &dmu, &dcv] {
                    memcpy(dst + off, src + off, len);
                    if (left.fetch_sub(1) == 1) {
                        std::lock_guard<std::mutex> l2(dmu);
                        dcv.notify_all();
                    }
                });
            }
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> lk(dmu);
        dcv.wait(lk, [&] { return left.load() == 0; });
    }

  private:
    void run() {
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                job = std::move(q_.front());
                q_.pop_front();
            }
            job();
        }
    }
    std::vector<std::thread> workers_;
    std::deque<std::function<void()>> q_;
    std::mutex mu_;
    std::condition_variable cv_;
    bool stop_ = false;
};

// Ring geometry and copy threads.  Measured on the B200 box (tools/host_copy_probe.py, profiles/r02l_host_copy_probe.log):
// memcpy pageable -> pinned 14 GB/s with one thread, 42 GB/s with 8, 49 GB/s with 12, 28 GB/s with 16 (of 16 cores); the
// link does 54 GB/s per direction.  RFB200_STAGE_THREADS / RFB200_STAGE_SLOT_MB override.
constexpr int NSLOTS = 4;
inline int env_or(const char *name, int dflt) {
    const char *e = getenv(name);
    return (e && atoi(e) > 0) ? atoi(e) : dflt;
}
const size_t SLOT_BYTES = (size_t)env_or("RFB200_STAGE_SLOT_MB", 32) << 20;

struct Slot {
    char *p = nullptr;
    cudaEvent_t ev = nullptr;
    bool dma_pending = false;  // uploads: the DMA out of this slot may still be running
    bool in_use = false;       // downloads: not yet unstaged
};

}  // namespace

struct Stager::Impl {
    Slot up[NSLOTS], down[NSLOTS];
    int up_i = 0, down_i = 0;
    CopyPool pool;
    std::thread drainer;
    std::mutex mu;
    std::condition_variable cv;
    struct Task { int slot; char *dst; size_t len; };
    std::deque<Task> tasks;
    size_t pending = 0;
    bool stop = false;
    std::atomic<bool> failed{false};

    Impl() : pool(env_or("RFB200_STAGE_THREADS", std::max(2, std::min(12, (int)std::thread::hardware_concurrency() * 3 / 4)))) {
        for (int i = 0; i < NSLOTS; ++i) {
            RFB_CUDA_CHECK(cudaHostAlloc((void **)&up[i].p, SLOT_BYTES, cudaHostAllocDefault));
            RFB_CUDA_CHECK(cudaHostAlloc((void **)&down[i].p, SLOT_BYTES, cudaHostAllocDefault));
            RFB_CUDA_CHECK(cudaEventCreateWithFlags(&up[i].ev, cudaEventDisableTiming));
            RFB_CUDA_CHECK(cudaEventCreateWithFlags(&down[i].ev, cudaEventDisableTiming));
        }
        drainer = std::thread([this] { drain(); });
    }
    void drain() {
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || !tasks.empty(); });
                if (tasks.empty()) return;
                t = tasks.front();
                tasks.pop_front();
            }
            if (cudaEventSynchronize(down[t.slot].ev) != cudaSuccess) failed = true;
            else pool.copy(t.dst, down[t.slot].p, t.len);
            {
                std::lock_guard<std::mutex> lk(mu);
                down[t.slot].in_use = false;
                --pending;
            }
            cv.notify_all();
        }
    }
};

Stager::Stager() : impl_(new Impl()) {}
Stager::~Stager() {
    // (process-lifetime object: the threads are detached from any CUDA teardown order by simply never being destroyed)
}

void Stager::upload(char *dst_dev, const char *src_host, size_t n, cudaStream_t s) {
    Impl &m = *impl_;
    for (size_t off = 0; off < n; off += SLOT_BYTES) {
        const size_t len = std::min(SLOT_BYTES, n - off);
        Slot &sl = m.up[m.up_i];
        m.up_i = (m.up_i + 1) % NSLOTS;
        if (sl.dma_pending) RFB_CUDA_CHECK(cudaEventSynchronize(sl.ev));
        m.pool.copy(sl.p, src_host + off, len);
        RFB_CUDA_CHECK(cudaMemcpyAsync(dst_dev + off, sl.p, len, cudaMemcpyHostToDevice, s));
        RFB_CUDA_CHECK(cudaEventRecord(sl.ev, s));
        sl.dma_pending = true;
    }
}

void Stager::download(char *dst_host, const char *src_dev, size_t n, cudaStream_t s) {
    Impl &m = *impl_;
    for (size_t off = 0; off < n; off += SLOT_BYTES) {
        const size_t len = std::min(SLOT_BYTES, n - off);
        const int si = m.down_i;
        m.down_i = (m.down_i + 1) % NSLOTS;
        {
            std::unique_lock<std::mutex> lk(m.mu);
            m.cv.wait(lk, [&] { return !m.down[si].in_use; });
            m.down[si].in_use = true;
            ++m.pending;
        }
        cudaError_t e = cudaMemcpyAsync(m.down[si].p, src_dev + off, len, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaEventRecord(m.down[si].ev, s);
        if (e != cudaSuccess) {
            std::lock_guard<std::mutex> lk(m.mu);
            m.down[si].in_use = false;
            --m.pending;
            RFB_CUDA_CHECK(e);
        }
        {
            std::lock_guard<std::mutex> lk(m.mu);
            m.tasks.push_back(Impl::Task{si, dst_host + off, len});
        }
        m.cv.notify_all();
    }
}

void Stager::finish() {
    Impl &m = *impl_;
    std::unique_lock<std::mutex> lk(m.mu);
    m.cv.wait(lk, [&] { return m.pending == 0; });
    for (auto &sl : m.up) sl.dma_pending = false;  // (the caller synchronises its streams before or after this)
    if (m.failed.exchange(false)) {
        set_error("device-to-host copy failed");
        throw Error();
    }
}

bool host_memory_is_pageable(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
}

}  // namespace rfb
