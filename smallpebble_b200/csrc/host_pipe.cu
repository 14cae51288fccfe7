// Host side of the input pipeline (README.md:310: `X[indices] / 255` is a
// float64 batch): page-locked float64 -> device fp32 with the NARROWING done on host cores while
// the copy is in flight, so that only half the bytes cross PCIe.
//
// The README loop's step is bound by host -> H2D -> forward -> loss read-back:
// 3.1 MB of float64 take 61 us at the 52 GB/s the link delivers.  One host core narrows at
// ~3 GB/s (numpy: 146 us for the batch), so the work is spread over a small persistent pool:
//   * the batch is cut into sub-chunks of 4096 elements which the workers claim from an atomic
//     counter and convert into a page-locked fp32 staging buffer (cvtpd2ps: the same
//     round-to-nearest-even as the device's cvt.rn.f32.f64 -- bit-identical to the old path);
//   * 16 sub-chunks make one DMA chunk (256 KB); the worker that completes a chunk's last
//     sub-chunk issues its cudaMemcpyAsync on the copy stream, so the first bytes are on the
//     wire ~5 us after the call and conversion and DMA overlap;
//   * the worker that issues the LAST chunk records the `landed` event and raises the job's
//     `issued` flag; the caller returns immediately and only checks that flag (normally long
//     set) before it orders the compute stream behind the event.
// Workers spin for a few hundred microseconds after a job (a training loop hands over the next
// batch within that window), then sleep on a condition variable.
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include <immintrin.h>
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int64_t kSub = 4096;        // elements per sub-chunk
constexpr int kSubsPerChunk = 16;     // sub-chunks per DMA chunk (256 KB of fp32)
constexpr int kMaxChunks = 4096;

struct Job {
  const double* src = nullptr;
  float* stage = nullptr;   // page-locked fp32
  float* dev = nullptr;
  int64_t n = 0;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t landed = nullptr;
  int64_t total_subs = 0;
  int total_chunks = 0;
  std::atomic<uint64_t> epoch{0};
  // ticket counter: (epoch << 32) | next sub-chunk.  start() publishes every field above, then
  // stores a fresh-epoch value here: a worker still draining the previous job either draws an
  // old-epoch ticket (beyond that job's end: it leaves) or a new-epoch one (and then sees the new
  // fields) -- no sub-chunk is ever claimed twice.
  std::atomic<uint64_t> next{0};
  std::atomic<int> chunk_left[kMaxChunks];  // sub-chunks of chunk c still to convert
  std::atomic<int> chunks_issued{0};
  std::atomic<int> issued{1};             // 1: every DMA of the job and its event are enqueued
  std::atomic<int> error{0};
};

void narrow(const double* __restrict__ s, float* __restrict__ d, int64_t n) {
  int64_t i = 0;
  for (; i + 4 <= n; i += 4) {  // SSE2: two cvtpd2ps per 4 elements (round to nearest even)
    const __m128 lo = _mm_cvtpd_ps(_mm_loadu_pd(s + i));
    const __m128 hi = _mm_cvtpd_ps(_mm_loadu_pd(s + i + 2));
    _mm_storeu_ps(d + i, _mm_movelh_ps(lo, hi));
  }
  for (; i < n; ++i) d[i] = (float)s[i];
}

class Pool {
 public:
  static Pool& get() {
    static Pool* p = new Pool();  // leaked on purpose: detached workers outlive static destructors
    return *p;
  }
  int threads() const { return (int)n_workers_; }

  // Claims and converts ONE sub-chunk; the thread that completes a DMA chunk puts it on the
  // wire, the one that issues the last chunk records the event.  False: nothing left to claim.
  bool work_one() {
    Job& j = job_;
    const uint64_t ticket = j.next.fetch_add(1, std::memory_order_acq_rel);
    if ((ticket >> 32) != (j.epoch.load(std::memory_order_acquire) & 0xffffffffu)) return false;
    const int64_t sub = (int64_t)(ticket & 0xffffffffu);
    if (sub >= j.total_subs) return false;
    const int64_t e0 = sub * kSub, e1 = std::min(j.n, e0 + kSub);
    narrow(j.src + e0, j.stage + e0, e1 - e0);
    const int c = (int)(sub / kSubsPerChunk);
    if (j.chunk_left[c].fetch_sub(1, std::memory_order_acq_rel) != 1) return true;
    static thread_local int bound = -1;  // this thread's CUDA device
    if (bound != j.device) {
      if (cudaSetDevice(j.device) != cudaSuccess) j.error.store(1);
      bound = j.device;
    }
    const int64_t c0 = (int64_t)c * kSubsPerChunk * kSub;
    const int64_t c1 = std::min(j.n, c0 + (int64_t)kSubsPerChunk * kSub);
    if (cudaMemcpyAsync(j.dev + c0, j.stage + c0, (size_t)(c1 - c0) * sizeof(float),
                        cudaMemcpyHostToDevice, j.stream) != cudaSuccess)
      j.error.store(1);
    if (j.chunks_issued.fetch_add(1, std::memory_order_acq_rel) + 1 == j.total_chunks) {
      if (cudaEventRecord(j.landed, j.stream) != cudaSuccess) j.error.store(1);
      j.issued.store(1, std::memory_order_release);
    }
    return true;
  }

  // Starts the job and returns; the previous job must have been waited for (wait_issued)
  void start(const double* src, float* stage, float* dev, int64_t n, int device, cudaStream_t st,
             cudaEvent_t landed) {
    Job& j = job_;
    // the epoch moves FIRST: from here on every ticket of the previous job is stale (a worker
    // still drawing one leaves without looking at the fields below), and no ticket of the new
    // epoch exists until `next` is stored at the end
    const uint64_t e = (j.epoch.load(std::memory_order_relaxed) + 1) & 0xffffffffu;
    j.epoch.store(e, std::memory_order_release);
    j.src = src;
    j.stage = stage;
    j.dev = dev;
    j.n = n;
    j.device = device;
    j.stream = st;
    j.landed = landed;
    j.total_subs = (n + kSub - 1) / kSub;
    j.total_chunks = (int)((j.total_subs + kSubsPerChunk - 1) / kSubsPerChunk);
    for (int c = 0; c < j.total_chunks; ++c) {
      const int64_t first = (int64_t)c * kSubsPerChunk;
      j.chunk_left[c].store((int)std::min<int64_t>(kSubsPerChunk, j.total_subs - first),
                            std::memory_order_relaxed);
    }
    j.chunks_issued.store(0, std::memory_order_relaxed);
    j.error.store(0, std::memory_order_relaxed);
    j.issued.store(0, std::memory_order_relaxed);
    j.next.store(e << 32, std::memory_order_release);
    seq_.fetch_add(1, std::memory_order_release);
    if (sleepers_.load(std::memory_order_acquire) > 0) {
      std::lock_guard<std::mutex> lk(m_);
      cv_.notify_all();
    }
  }

  // Spins (normally not at all) until every copy of the job and its event are enqueued; pitches
  // in when the pool makes no progress (workers asleep, or none could be started).
  int wait_issued() {
    Job& j = job_;
    int spins = 0;
    while (!j.issued.load(std::memory_order_acquire)) {
      if (n_workers_ == 0 || ++spins > 2000) {
        while (work_one()) {
        }
        spins = 0;
      }
      _mm_pause();
    }
    return j.error.load() ? 1 : 0;
  }

 private:
  Pool() {
    int want = 0;
    if (const char* e = getenv("SP_HOST_THREADS")) want = atoi(e);
    if (want <= 0) {
      const unsigned hw = std::thread::hardware_concurrency();
      int ranks = 1;
      if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, atoi(e));
      want = (int)hw / (2 * ranks);
    }
    want = std::max(1, std::min(want, 8));
    for (int i = 0; i < want; ++i) {
      try {
        std::thread(&Pool::loop, this).detach();
        ++n_workers_;
      } catch (...) {
        break;
      }
    }
  }

  void loop() {
    uint64_t seen = 0;
    for (;;) {
      // spin for the next job (~300 us: a training loop hands over its next batch within that
      // window), then sleep
      int spins = 0;
      while (seq_.load(std::memory_order_acquire) == seen) {
        if (++spins < 60000) {
          _mm_pause();
          continue;
        }
        std::unique_lock<std::mutex> lk(m_);
        sleepers_.fetch_add(1, std::memory_order_acq_rel);
        cv_.wait(lk, [&] { return seq_.load(std::memory_order_acquire) != seen; });
        sleepers_.fetch_sub(1, std::memory_order_acq_rel);
        break;
      }
      seen = seq_.load(std::memory_order_acquire);
      while (work_one()) {
      }
    }
  }

  Job job_;
  std::atomic<uint64_t> seq_{0};
  std::atomic<int> sleepers_{0};
  std::mutex m_;
  std::condition_variable cv_;
  size_t n_workers_ = 0;
};

}  // namespace

extern "C" {

// Worker threads of the host pipeline (started on first use; SP_HOST_THREADS overrides the
// default of min(8, cores / (2 * LOCAL_WORLD_SIZE))).
int sp_host_pipe_threads(int* threads) {
  SP_REQUIRE(threads != nullptr, "null output");
  *threads = Pool::get().threads();
  return 0;
}

// dev[0, n) (fp32, device) = src[0, n) (float64, host), narrowed on the host pool into the
// page-locked `stage` and copied chunk by chunk on `stream`; `landed` (an sp_event_create event)
// is recorded behind the last chunk.  RETURNS AT ONCE: call sp_h2d_f64_as_f32_finish() before
// ordering another stream behind `landed`, before starting another upload, and before touching
// `src` or `stage` again.
int sp_h2d_f64_as_f32_begin(const double* src, float* stage, float* dev, int64_t n, void* landed,
                            sp_stream_t stream) {
  SP_REQUIRE(src && stage && dev && landed, "null operand");
  SP_REQUIRE(n > 0 && (n + kSub - 1) / kSub <= (int64_t)kMaxChunks * kSubsPerChunk, "bad size");
  int device = 0;
  SP_CHECK_CUDA(cudaGetDevice(&device));
  Pool& pool = Pool::get();
  if (pool.wait_issued()) return sp::set_error(SP_ERR_DRIVER, "host pipe: previous upload failed");
  pool.start(src, stage, dev, n, device, sp::as_stream(stream), static_cast<cudaEvent_t>(landed));
  pool.work_one();  // the caller converts one sub-chunk itself (workers may still be waking up)
  return 0;
}

// Blocks (spinning, normally not at all) until every copy of the upload begun last and its event
// are enqueued.
int sp_h2d_f64_as_f32_finish(void) {
  if (Pool::get().wait_issued()) return sp::set_error(SP_ERR_DRIVER, "host pipe: upload failed");
  return 0;
}

}  // extern "C"
