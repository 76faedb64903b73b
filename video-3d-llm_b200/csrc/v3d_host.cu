// v3d_host.cu -- host-side loader helpers (C ABI: include/v3d.h).  No device code.
//
// The model-side front end runs at microseconds per frame, so the end-to-end rate is set by how fast a training
// sample's depth frames get from the page cache (packed scenes, scene_pack.py) into pinned memory.  Python threads
// doing that with numpy copies hand the interpreter lock back and forth once per frame; this gathers a whole batch
// in one call on a small persistent pool of native threads (the caller's interpreter lock is released by ctypes
// for the duration).
#include <emmintrin.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <unistd.h>
#include <vector>

#include "v3d_common.cuh"

namespace v3d {

class CopyPool {
 public:
  static CopyPool& get() {
    static CopyPool* p = new CopyPool;  // never destroyed: its detached workers wait on its condition variable at exit
    return *p;
  }
  // run fn(i) for i in [0, n) on up to n_threads workers (the caller participates)
  template <typename F>
  void run(int n, int n_threads, F&& fn) {
    if (n <= 0) return;
    std::unique_lock<std::mutex> call_lock(call_mu_);  // one gather at a time per process
    ensure(n_threads - 1);
    std::function<void(int)> f = fn;
    {
      std::lock_guard<std::mutex> lk(mu_);
      job_ = &f;
      next_.store(0);
      total_ = n;
      done_workers_ = 0;
      active_ = std::min<int>((int)workers_.size(), std::max(0, n_threads - 1));
      ++epoch_;
    }
    cv_.notify_all();
    drain(f);
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return done_workers_ == active_; });
    job_ = nullptr;
  }

 private:
  void drain(std::function<void(int)>& f) {
    for (;;) {
      const int i = next_.fetch_add(1);
      if (i >= total_) break;
      f(i);
    }
  }
  void ensure(int n) {
    n = std::min(n, 63);
    if (pid_ != getpid()) {  // forked child (e.g. a DataLoader worker): the parent's threads do not exist here
      workers_.clear();
      pid_ = getpid();
    }
    while ((int)workers_.size() < n) {
      const int id = (int)workers_.size();
      unsigned long born;
      {
        std::lock_guard<std::mutex> lk(mu_);
        born = epoch_;  // a worker only reacts to jobs posted after it was created
      }
      workers_.emplace_back([this, id, born] { loop(id, born); });
      workers_.back().detach();
    }
  }
  void loop(int id, unsigned long seen) {
    for (;;) {
      std::function<void(int)>* f = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (id >= active_) continue;
        f = job_;
      }
      if (f) drain(*f);
      {
        std::lock_guard<std::mutex> lk(mu_);
        ++done_workers_;
      }
      cv_done_.notify_one();
    }
  }
  std::mutex call_mu_, mu_;
  std::condition_variable cv_, cv_done_;
  std::vector<std::thread> workers_;
  std::function<void(int)>* job_ = nullptr;
  std::atomic<int> next_{0};
  int total_ = 0, active_ = 0, done_workers_ = 0;
  pid_t pid_ = 0;
  unsigned long epoch_ = 0;
};

}  // namespace v3d

using namespace v3d;

// memcpy whose stores bypass the cache (movntdq): the destination is a pinned staging buffer that only the DMA engine reads
// next, so pulling its lines into the cache first (read-for-ownership) would be a third pass over host DRAM per byte --
// and host DRAM is what bounds the loader when several ranks gather at once.  dst 16-byte aligned; any src.
static void copy_streaming(unsigned char* dst, const unsigned char* src, size_t len) {
  if ((reinterpret_cast<uintptr_t>(dst) & 15) != 0 || len < 256) {
    memcpy(dst, src, len);
    return;
  }
  size_t i = 0;
  for (; i + 64 <= len; i += 64) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 16));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 32));
    const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 48));
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 32), c);
    _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 48), d);
  }
  if (i < len) memcpy(dst + i, src + i, len - i);
  _mm_sfence();
}

extern "C" {

int v3d_host_gather(void* dst_host, const void* const* src_host, size_t bytes_each, int n_blocks, int n_threads) {
  if (n_blocks == 0) return V3D_OK;
  V3D_REQUIRE(dst_host && src_host, V3D_E_ARG, "v3d_host_gather: null pointer");
  V3D_REQUIRE(n_blocks > 0 && bytes_each > 0, V3D_E_ARG, "v3d_host_gather: bad sizes");
  if (n_threads < 1) n_threads = 1;
  // split large blocks so that every thread has work even for a handful of frames
  const size_t piece = 256 * 1024;
  const int per_block = (int)((bytes_each + piece - 1) / piece);
  unsigned char* dst = static_cast<unsigned char*>(dst_host);
  CopyPool::get().run(n_blocks * per_block, n_threads, [&](int i) {
    const int b = i / per_block, p = i - b * per_block;
    const size_t off = (size_t)p * piece;
    const size_t len = off + piece <= bytes_each ? piece : bytes_each - off;
    copy_streaming(dst + (size_t)b * bytes_each + off, static_cast<const unsigned char*>(src_host[b]) + off, len);
  });
  return V3D_OK;
}


int v3d_h2d_gather(void* dst_dev, const void* const* src_host, size_t bytes_each, int n_blocks, v3d_stream_t stream) {
  if (n_blocks == 0) return V3D_OK;
  V3D_REQUIRE(dst_dev && src_host, V3D_E_ARG, "v3d_h2d_gather: null pointer");
  V3D_REQUIRE(n_blocks > 0 && bytes_each > 0, V3D_E_ARG, "v3d_h2d_gather: bad sizes");
  unsigned char* dst = static_cast<unsigned char*>(dst_dev);
  cudaStream_t s = (cudaStream_t)stream;
  // consecutive blocks that are also consecutive in host memory go out as one DMA
  int i = 0;
  while (i < n_blocks) {
    int j = i + 1;
    while (j < n_blocks && static_cast<const unsigned char*>(src_host[j]) == static_cast<const unsigned char*>(src_host[j - 1]) + bytes_each) ++j;
    V3D_CUDA(cudaMemcpyAsync(dst + (size_t)i * bytes_each, src_host[i], (size_t)(j - i) * bytes_each, cudaMemcpyHostToDevice, s));
    i = j;
  }
  return V3D_OK;
}

}  // extern "C"
