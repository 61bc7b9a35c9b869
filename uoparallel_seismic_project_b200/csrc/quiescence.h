// quiescence.h -- termination detection of the asynchronous multi-device relaxation (solver.cu, sweeptt_solve_slabs).
// Plain C++ (no CUDA): tests/test_quiescence.py drives it against a randomised model of parts that wake each other up.
#pragma once
#include <mutex>
#include <vector>

namespace sweeptt {

// Termination of the asynchronous multi-device relaxation (no barrier): every part publishes after each batch how many
// tiles it has relaxed in total and whether anything is pending; the job is finished when all parts are idle over a
// window in which every part completed at least two further batches (so at least one started after the window opened
// and saw every wake-up sent before it) without relaxing a single tile.
struct Quiescence {
  std::mutex mu;
  int n = 0;
  std::vector<long long> seq, visits, seq0, visits0;
  std::vector<char> idle;
  bool window = false, done = false, failed = false;
  explicit Quiescence(int parts) : n(parts), seq(parts, 0), visits(parts, 0), seq0(parts, 0), visits0(parts, 0), idle(parts, 0) {}
  bool report(int p, long long total_visits, bool is_idle) {  // returns true when the job is finished
    std::lock_guard<std::mutex> lk(mu);
    if (done || failed) return true;
    seq[p] += 1; visits[p] = total_visits; idle[p] = is_idle;
    bool all_idle = true;
    for (int q = 0; q < n; ++q) all_idle = all_idle && idle[q] && seq[q] > 0;
    if (!all_idle) { window = false; return false; }
    if (!window) { seq0 = seq; visits0 = visits; window = true; return false; }
    for (int q = 0; q < n; ++q) {
      if (visits[q] != visits0[q]) { seq0 = seq; visits0 = visits; return false; }  // (cannot happen while idle; restart)
      if (seq[q] < seq0[q] + 2) return false;
    }
    done = true;
    return true;
  }
  void abort() { std::lock_guard<std::mutex> lk(mu); failed = true; }
};

}  // namespace sweeptt
