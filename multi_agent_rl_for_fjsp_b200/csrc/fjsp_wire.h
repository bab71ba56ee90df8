// fjsp_wire.h — HOST side of the wire format (include/fjsp_b200.h "wire rows"): decode of the compact rows the step
// kernel writes on the host-buffer path into the float32 / int8 tensors of fjsp_step, and the small thread pool that
// does it while later chunks are still crossing PCIe.  Format conversion only: no simulation logic lives here.
#ifndef FJSP_WIRE_H
#define FJSP_WIRE_H

#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include <sched.h>

#include "fjsp_core.h"

namespace fjsp {

// rows [lo, hi) of a `cells`-cell shop -> obs / masks / rewards / flags (any of them may be null).  Bit for bit what
// fjsp_step_kernel<K, false> writes: the same integers converted to float32, the same table look-ups, the same single
// IEEE division per reward.  Defined in fjsp_wire.cpp (AVX2 when the CPU has it, portable C++ otherwise).
void wire_decode(int cells, const Params& P, const u32* wire, int64_t lo, int64_t hi, float* obs, int8_t* masks, float* rewards,
                 uint8_t* flags);
const char* wire_decode_isa();  // "avx2" or "generic"
// one pass of streaming stores over `bytes` from `threads` threads (the box's ceiling for the decode's writes): seconds
double host_stream_write_seconds(void* buf, size_t bytes, int threads);

inline int usable_cpus() {
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
        const int n = CPU_COUNT(&set);
        if (n > 0) return n;
    }
    const unsigned hc = std::thread::hardware_concurrency();
    return hc ? (int)hc : 1;
}

// Decode workers of one handle.  Jobs are env ranges of the chunk whose D2H copy has completed; the submitting thread
// keeps waiting on the next chunk's event while the workers decode.  Created on the first fjsp_step_host call.
class DecodePool {
  public:
    struct Job {
        int cells;
        const Params* P;
        const u32* wire;   // row 0 of the handle's pinned staging
        int64_t lo, hi;
        float* obs;
        int8_t* masks;
        float* rewards;
        uint8_t* flags;
    };
    explicit DecodePool(int nthreads) {
        for (int i = 0; i < nthreads; i++) th_.emplace_back([this] { run(); });
    }
    ~DecodePool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_ = true;
        }
        cv_work_.notify_all();
        for (auto& t : th_) t.join();
    }
    int size() const { return (int)th_.size(); }
    void submit(const Job& j, int64_t grain) {
        {
            std::lock_guard<std::mutex> g(m_);
            for (int64_t lo = j.lo; lo < j.hi; lo += grain) {
                Job s = j;
                s.lo = lo, s.hi = lo + grain < j.hi ? lo + grain : j.hi;
                q_.push_back(s);
                pending_++;
            }
        }
        cv_work_.notify_all();
    }
    void wait() {  // the caller decodes too instead of idling
        std::unique_lock<std::mutex> g(m_);
        while (pending_ > 0) {
            if (!q_.empty()) {
                Job j = q_.front();
                q_.pop_front();
                g.unlock();
                wire_decode(j.cells, *j.P, j.wire, j.lo, j.hi, j.obs, j.masks, j.rewards, j.flags);
                g.lock();
                pending_--;
            } else {
                cv_done_.wait(g);
            }
        }
    }

  private:
    void run() {
        std::unique_lock<std::mutex> g(m_);
        for (;;) {
            cv_work_.wait(g, [this] { return stop_ || !q_.empty(); });
            if (stop_) return;
            Job j = q_.front();
            q_.pop_front();
            g.unlock();
            wire_decode(j.cells, *j.P, j.wire, j.lo, j.hi, j.obs, j.masks, j.rewards, j.flags);
            g.lock();
            if (--pending_ == 0) cv_done_.notify_all();
        }
    }
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_work_, cv_done_;
    std::deque<Job> q_;
    int64_t pending_ = 0;
    bool stop_ = false;
};

}  // namespace fjsp
#endif  // FJSP_WIRE_H
