// Host -> device staging engine for PAGEABLE host buffers (what the reference's callers pass: NumPy arrays,
// evaluation/metrics.py:322-340).  cudaMemcpy from pageable memory is staged by the driver through one bounce buffer
// with a single-threaded memcpy (13-17 GB/s measured on the B200 boxes); this engine does the staging itself:
//   * a ring of pinned slots (4 x 4 MiB, allocated once per process),
//   * a persistent pool of worker threads that claim 256 KiB pieces of a chunk and memcpy them into its slot in
//     parallel with the caller (the caller never waits for a sleeping worker to wake up),
//   * one cudaMemcpyAsync per chunk on the caller's stream, an event per slot guarding its reuse,
// so the CPU copy of chunk k+1 overlaps the PCIe transfer of chunk k and the copy rate is that of several cores.
// The call returns once every chunk is enqueued (the last transfers are still in flight; the source may be reused
// immediately -- it has been copied out).  One call at a time per process (mutex); the GIL is not held by ctypes callers.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace runia {
namespace {

constexpr size_t kSlotBytes = 4u << 20;
constexpr int kSlots = 4;
constexpr size_t kPieceBytes = 256u << 10;  // unit of work the threads claim (16 per full chunk)

struct Desc {  // one chunk to copy: pieces [i * kPieceBytes, ...) of n bytes.  `gen` brackets the fields like a seqlock:
  // a thread that overshot the piece count of an old generation may look at a slot that is being rewritten for a
  // newer one, and must not act on it
  std::atomic<char *> dst{nullptr};
  std::atomic<const char *> src{nullptr};
  std::atomic<size_t> n{0};
  std::atomic<uint32_t> npieces{0};
  std::atomic<uint64_t> gen{0};
};

static inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#else
  std::this_thread::yield();
#endif
}

// Work distribution: `ticket_` = (generation << 32) | next piece index.  Every thread -- the caller included --
// claims pieces of the current chunk with one fetch_add; a chunk is done when `done_` reaches its piece count.
// The caller never waits for a worker to WAKE UP: if the workers are asleep it simply copies the pieces itself
// (the speed of a plain cudaMemcpy from pageable memory) and they join as they arrive.  After their last piece the
// workers spin for `spin_us_` before they go back to sleep on the condition variable, so that the chunks of one
// upload and back-to-back postprocess() calls find them hot (a futex wake-up costs 100-300 us on the B200 boxes'
// virtual CPUs: with sleeping workers a 20 MB call takes 1.1-2.3 ms, with hot ones 0.6 ms; plain cudaMemcpy: 1.0 ms).
// Measured trade-off (profiles/r2_h2d_probe.jsonl, scripts/sweep_probe.py, 16 vCPUs): 4 threads reach the PCIe rate on
// a 20 MB call (51 GB/s) but only 29 GB/s on 1 GB, 8 threads 52 GB/s on both; with 8 spinning threads one 20 MB call in
// ten waits 5-15 ms for a worker that was descheduled while it held a claimed piece (the host also runs the BLAS
// threads of the callers' fits).  The caller must wait for such a straggler -- it reads the caller's buffer, which may
// be unmapped the moment the call returns -- so uploads below 32 MB enlist three workers, longer ones all seven.
class Stager {
 public:
  static Stager &get() {  // one engine per device (its events belong to that device's context)
    static Stager s[16];
    int dev = 0;
    cudaGetDevice(&dev);
    return s[dev & 15];
  }

  int upload(void *dst_dev, const void *src_host, size_t bytes, cudaStream_t st) {
    std::lock_guard<std::mutex> call_lock(call_mu_);
    int rc = ensure();
    if (rc) return rc;
    const char *src = static_cast<const char *>(src_host);
    char *dst = static_cast<char *>(dst_dev);
    // 4 threads reach the PCIe rate on a 20 MB call and keep the number of spinning threads small (fewer stragglers);
    // a long upload needs 8 to stay at 50 GB/s (4 threads: 29 GB/s on 1 GB) and can absorb a straggler in its ring
    int want = bytes >= (32u << 20) ? (int)workers_.size() : std::min<int>(3, (int)workers_.size());
    // below one slot the copy takes less time (~0.1 ms) than a descheduled helper can cost (milliseconds): alone
    if (bytes < kSlotBytes) want = 0;
    // a worker that was descheduled while it held a claimed piece made the previous call wait: the host's cores are
    // taken (BLAS threads of a fit spin for ~100 ms after their last call) -- copy alone until that has passed
    if (std::chrono::steady_clock::now() < solo_until_) want = 0;
    active_.store(want, std::memory_order_relaxed);
    for (size_t off = 0; off < bytes; off += kSlotBytes) {
      const size_t n = bytes - off < kSlotBytes ? bytes - off : kSlotBytes;
      const int s = next_slot_;
      next_slot_ = (next_slot_ + 1) % kSlots;
      if (busy_[s]) RUNIA_CUDA(cudaEventSynchronize(ev_[s]));  // the previous transfer out of this slot has finished
      fill(slot_[s], src + off, n);
      RUNIA_CUDA(cudaMemcpyAsync(dst + off, slot_[s], n, cudaMemcpyHostToDevice, st));
      RUNIA_CUDA(cudaEventRecord(ev_[s], st));
      busy_[s] = true;
    }
    return RUNIA_OK;
  }

 private:
  Stager() = default;
  ~Stager() {
    stop_.store(true);
    {
      std::lock_guard<std::mutex> lk(mu_);
    }
    cv_.notify_all();
    for (auto &t : workers_) t.join();
    // pinned slots and events are left to process teardown (the CUDA context may already be gone)
  }

  int ensure() {
    if (ready_) return RUNIA_OK;
    for (int s = 0; s < kSlots; ++s) {
      RUNIA_CUDA(cudaHostAlloc((void **)&slot_[s], kSlotBytes, cudaHostAllocPortable));
      RUNIA_CUDA(cudaEventCreateWithFlags(&ev_[s], cudaEventDisableTiming));
      busy_[s] = false;
    }
    unsigned hw = std::thread::hardware_concurrency();
    int n = hw >= 16 ? 7 : hw >= 8 ? 3 : hw >= 4 ? 1 : 0;  // + the calling thread; small uploads enlist only three of them
    if (const char *e = getenv("RUNIA_B200_STAGE_THREADS")) n = std::max(0, atoi(e) - 1);
    if (const char *e = getenv("RUNIA_B200_STAGE_SPIN_US")) spin_us_ = std::max(0, atoi(e));
    for (int i = 0; i < n; ++i) workers_.emplace_back([this, i] { worker(i); });
    ready_ = true;
    return RUNIA_OK;
  }

  bool work_available() const {
    const uint64_t t = ticket_.load(std::memory_order_seq_cst);
    const uint64_t g = t >> 32;
    const Desc &d = ring_[g & 3];
    return g != 0 && d.gen.load(std::memory_order_acquire) == g && (uint32_t)t < d.npieces.load(std::memory_order_relaxed);
  }

  // claims and copies pieces until none is left; returns true if it copied at least one
  bool claim_loop() {
    bool any = false;
    for (;;) {
      if (!work_available()) return any;
      const uint64_t t = ticket_.fetch_add(1, std::memory_order_acq_rel);
      const uint64_t g = t >> 32;
      const uint32_t i = (uint32_t)t;
      const Desc &d = ring_[g & 3];  // written before the ticket of generation g was published
      if (g == 0 || d.gen.load(std::memory_order_acquire) != g) continue;
      char *dst = d.dst.load(std::memory_order_relaxed);
      const char *src = d.src.load(std::memory_order_relaxed);
      const size_t n = d.n.load(std::memory_order_relaxed);
      const uint32_t np = d.npieces.load(std::memory_order_relaxed);
      if (d.gen.load(std::memory_order_acquire) != g) continue;  // the slot moved on: this was an overshoot of an old chunk
      if (i < np) {  // a valid claim pins the generation: the caller cannot advance before `done_` counts this piece
        const size_t off = (size_t)i * kPieceBytes;
        memcpy(dst + off, src + off, n - off < kPieceBytes ? n - off : kPieceBytes);
        done_[g & 3].fetch_add(1, std::memory_order_release);
        any = true;
      }
    }
  }

  void fill(char *dst, const char *src, size_t n) {
    if (workers_.empty() || n <= kPieceBytes || active_.load(std::memory_order_relaxed) == 0) {
      memcpy(dst, src, n);
      return;
    }
    const uint64_t g = ++gen_;
    const uint32_t np = (uint32_t)((n + kPieceBytes - 1) / kPieceBytes);
    Desc &d = ring_[g & 3];
    d.gen.store(0, std::memory_order_release);  // invalidate while the fields change
    d.dst.store(dst, std::memory_order_relaxed);
    d.src.store(src, std::memory_order_relaxed);
    d.n.store(n, std::memory_order_relaxed);
    d.npieces.store(np, std::memory_order_relaxed);
    done_[g & 3].store(0, std::memory_order_relaxed);
    d.gen.store(g, std::memory_order_release);
    ticket_.store(g << 32, std::memory_order_seq_cst);
    if (sleepers_.load(std::memory_order_seq_cst) > 0) {
      { std::lock_guard<std::mutex> lk(mu_); }
      cv_.notify_all();
    }
    claim_loop();
    if (done_[g & 3].load(std::memory_order_acquire) != np) {
      const auto t0 = std::chrono::steady_clock::now();
      while (done_[g & 3].load(std::memory_order_acquire) != np) cpu_relax();
      const auto waited = std::chrono::steady_clock::now() - t0;
      if (waited > std::chrono::microseconds(150))  // a piece takes ~15 us to copy: its holder lost its core
        solo_until_ = std::chrono::steady_clock::now() + std::chrono::milliseconds(25);
    }
  }

  void worker(int idx) {
    using clock = std::chrono::steady_clock;
    auto last = clock::now();
    for (;;) {
      if (stop_.load(std::memory_order_relaxed)) return;
      if (idx < active_.load(std::memory_order_relaxed) && claim_loop()) {
        last = clock::now();
        continue;
      }
      if (std::chrono::duration_cast<std::chrono::microseconds>(clock::now() - last).count() < spin_us_) {
        cpu_relax();
        continue;
      }
      sleepers_.fetch_add(1, std::memory_order_seq_cst);
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait_for(lk, std::chrono::milliseconds(100), [&] {
          return stop_.load() || (idx < active_.load(std::memory_order_relaxed) && work_available());
        });
      }
      sleepers_.fetch_sub(1, std::memory_order_seq_cst);
      last = clock::now();
    }
  }

  std::mutex call_mu_, mu_;
  std::condition_variable cv_;
  std::atomic<bool> stop_{false};
  bool ready_ = false;
  int spin_us_ = 1000;  // keeps the workers hot across back-to-back calls; see the measurements above
  uint64_t gen_ = 0;
  std::atomic<uint64_t> ticket_{0};
  std::atomic<uint32_t> done_[4];
  std::atomic<int> sleepers_{0};
  std::atomic<int> active_{0};  // workers [0, active_) take part in the current upload
  std::chrono::steady_clock::time_point solo_until_{};  // stragglers seen: no workers before this time
  Desc ring_[4];
  std::vector<std::thread> workers_;
  char *slot_[kSlots] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_[kSlots];
  bool busy_[kSlots] = {false, false, false, false};
  int next_slot_ = 0;
};

}  // namespace
}  // namespace runia

extern "C" int runia_stage_h2d(void *dst_dev, const void *src_host, int64_t bytes, void *stream) {
  RUNIA_NVTX();
  RUNIA_REQUIRE(bytes >= 0, RUNIA_E_BADARG, "stage_h2d: negative size");
  if (bytes == 0) return RUNIA_OK;
  RUNIA_REQUIRE(dst_dev && src_host, RUNIA_E_BADARG, "stage_h2d: null pointer");
  return runia::Stager::get().upload(dst_dev, src_host, (size_t)bytes, (cudaStream_t)stream);
}
