// ls2d_stage.h -- uploads from PAGEABLE caller buffers (host side only).
//
// cudaMemcpyAsync from pageable memory is a synchronous, single-threaded staged copy (~11 GB/s measured here): a
// caller that hands plain malloc'ed clouds to ls2d_align_pairs_host spends 12 ms on 142 MB that the link moves in
// 2.7 ms.  stage_pool copies the caller's bytes into a small ring of pinned slots with a few threads and sends every
// slot on with an asynchronous copy, so the host copy and the DMA overlap and run at several threads' worth of memory
// bandwidth.  Pinned or registered buffers bypass it.
#pragma once

#include <cuda_runtime.h>

#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace ls2d {

class stage_pool {
 public:
  static constexpr int SLOTS        = 4;
  static constexpr size_t SLOT_SIZE = 4u << 20;

  stage_pool() = default;
  stage_pool(const stage_pool&) = delete;
  stage_pool& operator=(const stage_pool&) = delete;
  ~stage_pool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      quit_ = true;
    }
    cv_work_.notify_all();
    for (std::thread& t : workers_) t.join();
    for (cudaEvent_t e : sent_)
      if (e) cudaEventDestroy(e);
    if (ring_) cudaFreeHost(ring_);
  }

  // dst (device) <- src (host, pinned or not), asynchronous on `stream` for pinned sources; for pageable sources the
  // call returns when the last slot has been handed to the copy engine
  cudaError_t upload(void* dst, const void* src, size_t bytes, cudaStream_t stream) {
    if (bytes == 0) return cudaSuccess;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, src) != cudaSuccess) {
      cudaGetLastError();  // very old drivers report unregistered memory as an error
      at.type = cudaMemoryTypeUnregistered;
    }
    if (at.type != cudaMemoryTypeUnregistered || bytes < SLOT_SIZE / 4)
      return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream);
    if (cudaError_t e = start()) return e;
    for (size_t off = 0; off < bytes; off += SLOT_SIZE) {
      const size_t len = bytes - off < SLOT_SIZE ? bytes - off : SLOT_SIZE;
      const int s      = next_++ % SLOTS;
      if (cudaError_t e = cudaEventSynchronize(sent_[s])) return e;  // the slot's previous contents are on the device
      char* slot = ring_ + (size_t) s * SLOT_SIZE;
      parallel_copy(slot, (const char*) src + off, len);
      if (cudaError_t e = cudaMemcpyAsync((char*) dst + off, slot, len, cudaMemcpyHostToDevice, stream)) return e;
      if (cudaError_t e = cudaEventRecord(sent_[s], stream)) return e;
    }
    return cudaSuccess;
  }

 private:
  cudaError_t start() {
    if (ring_) return cudaSuccess;
    if (cudaError_t e = cudaHostAlloc((void**) &ring_, SLOTS * SLOT_SIZE, cudaHostAllocDefault)) return e;
    for (cudaEvent_t& ev : sent_)
      if (cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) return e;
    unsigned hw = std::thread::hardware_concurrency();
    const int n = hw >= 16 ? 6 : (hw >= 8 ? 4 : 2);
    for (int i = 0; i < n - 1; ++i) workers_.emplace_back([this, i] { work(i + 1); });
    parts_ = n;
    return cudaSuccess;
  }

  void parallel_copy(char* dst, const char* src, size_t len) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      dst_ = dst, src_ = src, len_ = len, pending_ = parts_ - 1;
      ++job_;
    }
    cv_work_.notify_all();
    copy_part(0);
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [this] { return pending_ == 0; });
  }

  void copy_part(int part) {
    const size_t per = ((len_ + parts_ - 1) / parts_ + 63) & ~(size_t) 63;
    const size_t lo = per * part, hi = lo + per < len_ ? lo + per : len_;
    if (lo < hi) std::memcpy(dst_ + lo, src_ + lo, hi - lo);
  }

  void work(int part) {
    unsigned long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_work_.wait(lk, [&] { return quit_ || job_ != seen; });
        if (quit_) return;
        seen = job_;
      }
      copy_part(part);
      bool last;
      {
        std::lock_guard<std::mutex> lk(mu_);
        last = --pending_ == 0;
      }
      if (last) cv_done_.notify_one();
    }
  }

  char* ring_ = nullptr;
  cudaEvent_t sent_[SLOTS] = {};
  unsigned next_ = 0;
  std::vector<std::thread> workers_;
  int parts_ = 1;
  std::mutex mu_;
  std::condition_variable cv_work_, cv_done_;
  bool quit_ = false;
  unsigned long job_ = 0;
  int pending_ = 0;
  char* dst_ = nullptr;
  const char* src_ = nullptr;
  size_t len_ = 0;
};

}  // namespace ls2d
