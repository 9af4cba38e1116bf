/* group_workers.hpp -- the enqueueing threads of a device group (hexed_b200_group_*, group.cu). Plain C++ (no CUDA) so that the dispatch logic is
 * unit-tested on the CPU (tests/test_group_workers.py). */
#ifndef HEXED_B200_GROUP_WORKERS_HPP_
#define HEXED_B200_GROUP_WORKERS_HPP_

#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace hb {

/* One enqueueing thread per rank. A stage is ~8 kernel launches per rank; issued by ONE host thread rank after rank, the last of eight devices
 * starts a stage ~0.3 ms after the first, every stage, because nothing can be enqueued ahead of the admissibility answer Solver::update waits
 * for. run(f) executes f(r) for every rank -- rank 0 on the caller, the others on their workers -- and returns the
 * first non-zero code. NCCL group calls stay on the calling thread. The host-thread emulation build keeps the serial loop (its launches share
 * global state). HEXED_B200_GROUP_THREADS=0 in the environment keeps the serial loop as well. */
struct Workers
{
  int n = 0;
  bool threaded = false;
  std::vector<std::thread> th;
  std::mutex m;
  std::condition_variable cv_go, cv_done;
  const std::function<int(int)>* job = nullptr;
  unsigned long long epoch = 0;
  int pending = 0;
  bool stop = false;
  std::vector<int> rc;

  void start(int n_)
  {
    n = n_; rc.assign(n, 0);
#ifndef HB_EMULATE
    const char* env = std::getenv("HEXED_B200_GROUP_THREADS");
    threaded = n > 1 && !(env && env[0] == '0');
#endif
    if (!threaded) return;
    for (int r = 1; r < n; ++r) th.emplace_back([this, r]() {
      unsigned long long seen = 0;
      for (;;) {
        const std::function<int(int)>* f;
        {
          std::unique_lock<std::mutex> lk(m);
          cv_go.wait(lk, [&]() {return stop || epoch != seen;});
          if (stop) return;
          seen = epoch; f = job;
        }
        const int code = (*f)(r);
        {
          std::lock_guard<std::mutex> lk(m);
          rc[r] = code;
          if (--pending == 0) cv_done.notify_one();
        }
      }
    });
  }
  int run(const std::function<int(int)>& f)
  {
    if (!threaded) { for (int r = 0; r < n; ++r) { const int code = f(r); if (code) return code; } return 0; }
    {
      std::lock_guard<std::mutex> lk(m);
      job = &f; pending = n - 1; ++epoch;
    }
    cv_go.notify_all();
    rc[0] = f(0);
    {
      std::unique_lock<std::mutex> lk(m);
      cv_done.wait(lk, [&]() {return pending == 0;});
    }
    for (int r = 0; r < n; ++r) if (rc[r]) return rc[r];
    return 0;
  }
  ~Workers()
  {
    { std::lock_guard<std::mutex> lk(m); stop = true; }
    cv_go.notify_all();
    for (std::thread& t : th) t.join();
  }
};


} // namespace hb
#endif
