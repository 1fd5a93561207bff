// TEST INFRASTRUCTURE (oracle/): stand-in for Intel TBB's <tbb/pipeline.h>, which is absent
// from this image.  It exists only so that the reference's own sources (/root/reference/src,
// compiled where they lie, see oracle/Makefile) build into oracle/_ref/.  Not product code;
// nothing under experimental-mf_b200/ may include it.
//
// Semantics kept (the subset the reference uses: main.cc:45-50, mf.h:18,56,75):
//   * tbb::filter(mode) with virtual void* operator()(void*);
//   * tbb::pipeline::add_filter / run(max_number_of_live_tokens) / clear;
//   * the first filter is serial_in_order and is the token source: it is called with NULL and
//     ends the stream by returning NULL; later filters are `parallel`;
//   * at most `max_number_of_live_tokens` tokens are in flight.  With 1 token execution is
//     strictly serial in file order (the "--fly 1" single-thread update order).
// Scheduling model: N worker threads; each takes the source lock, pulls one token from filter 0,
// releases the lock and carries the token through the remaining filters itself.  That gives
// exactly "serial in-order source, parallel tail, <= N live tokens".
#ifndef ORACLE_SHIM_TBB_PIPELINE_H
#define ORACLE_SHIM_TBB_PIPELINE_H

#include <cstddef>
#include <mutex>
#include <thread>
#include <vector>

namespace tbb {

class filter {
 public:
  enum mode { parallel = 0, serial_in_order = 1, serial_out_of_order = 2, serial = 1 };
  explicit filter(mode m) : mode_(m) {}
  virtual ~filter() {}
  virtual void* operator()(void* item) = 0;
  bool is_serial() const { return mode_ != parallel; }

 private:
  mode mode_;
};

class pipeline {
 public:
  pipeline() {}
  ~pipeline() {}
  void add_filter(filter& f) { stages_.push_back(&f); }
  void clear() { stages_.clear(); }

  void run(size_t max_number_of_live_tokens) {
    if (stages_.empty()) return;
    size_t n = max_number_of_live_tokens < 1 ? 1 : max_number_of_live_tokens;
    done_ = false;
    if (n == 1) {
      worker();
      return;
    }
    std::vector<std::thread> pool;
    pool.reserve(n - 1);
    for (size_t i = 0; i + 1 < n; i++) pool.emplace_back([this] { worker(); });
    worker();
    for (auto& t : pool) t.join();
  }

 private:
  void worker() {
    for (;;) {
      void* tok;
      {
        std::lock_guard<std::mutex> g(source_lock_);
        if (done_) return;
        tok = (*stages_[0])(NULL);
        if (tok == NULL) {
          done_ = true;
          return;
        }
      }
      for (size_t s = 1; s < stages_.size(); s++) {
        if (stages_[s]->is_serial()) {
          std::lock_guard<std::mutex> g(serial_lock_);
          tok = (*stages_[s])(tok);
        } else {
          tok = (*stages_[s])(tok);
        }
      }
    }
  }

  std::vector<filter*> stages_;
  std::mutex source_lock_, serial_lock_;
  bool done_ = false;
};

}  // namespace tbb

#endif
