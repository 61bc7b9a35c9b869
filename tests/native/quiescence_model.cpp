// Randomised model check of csrc/quiescence.h (test infrastructure; built and run by tests/test_quiescence.py).
//
// N parts run batches asynchronously.  A batch that starts at time t0 processes the wake-ups that were delivered to
// the part before a cut-off inside the batch, still SEES as pending those delivered until a second point (its last
// compaction) and knows nothing of later ones (marks that land late in a batch are only picked up by the next one --
// the situation the real rounds are in); processing a wake-up is work and sends new wake-ups (delivered at once) to
// random parts while a budget lasts.  After every batch the part reports (total work, idle) exactly like
// solve_slabs_impl does.  The detector must never announce the end while a wake-up is still undelivered/unprocessed,
// and must announce it once everything is quiet.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "quiescence.h"

struct Batch { bool running = false; double cutoff = 0, seen = 0, end = 0; };

int main(int argc, char** argv) {
  const int runs = argc > 1 ? atoi(argv[1]) : 2000;
  long long false_ends = 0, never_ended = 0, total_reports = 0;
  for (int run = 0; run < runs; ++run) {
    std::mt19937_64 rng(1234567 + run);
    auto uni = [&](double a, double b) { return std::uniform_real_distribution<double>(a, b)(rng); };
    const int n = 1 + (int)(rng() % 8);
    sweeptt::Quiescence q(n);
    std::vector<std::vector<double>> inbox(n);  // delivery times of unprocessed wake-ups
    std::vector<long long> work(n, 0);
    std::vector<Batch> b(n);
    long long budget = 5 + (long long)(rng() % 400);  // wake-ups that may still be generated
    inbox[rng() % n].push_back(0.0);                  // the start point
    double now = 0;
    bool ended = false;
    for (long long step = 0; step < 200000 && !ended; ++step) {
      // next event: the earliest batch end, or start a batch on an idle part
      int p = (int)(rng() % n);
      if (!b[p].running) {
        b[p].running = true;
        const double len = uni(0.1, 3.0) * (rng() % 5 == 0 ? 10.0 : 1.0);  // parts run at very different speeds
        b[p].cutoff = now + uni(0.0, 1.0) * len;                         // wake-ups delivered until here are processed
        b[p].seen = b[p].cutoff + uni(0.0, 1.0) * (now + len - b[p].cutoff);  // ... until here: seen as "pending" (last compaction)
        b[p].end = now + len;                                            // later ones are invisible to this batch's report
        now += uni(0.0, 0.2);
        continue;
      }
      // finish p's batch (time jumps to its end if later)
      if (b[p].end > now) now = b[p].end;
      long long before = work[p];
      std::vector<double> keep;
      for (double t : inbox[p]) {
        if (t <= b[p].cutoff) {
          work[p] += 1;
          const int fan = budget > 0 ? (int)(rng() % 4) : 0;
          for (int k = 0; k < fan && budget > 0; ++k, --budget) {
            // wake-ups sent by this batch are delivered somewhere inside the batch's lifetime
            inbox[rng() % n].push_back(uni(b[p].cutoff, b[p].end));
          }
        } else {
          keep.push_back(t);
        }
      }
      // (a wake-up p sent to itself during this batch stays pending)
      std::vector<double> mine;
      for (double t : inbox[p]) if (t > b[p].cutoff) mine.push_back(t);
      inbox[p].swap(mine);
      (void)keep;
      bool pending = false;
      for (double t : inbox[p]) pending = pending || t <= b[p].seen;
      const bool worked = work[p] != before;
      b[p].running = false;
      ++total_reports;
      if (q.report(p, work[p], !pending && !worked)) {
        ended = true;
        for (int r = 0; r < n; ++r)
          if (!inbox[r].empty()) { ++false_ends; break; }
      }
    }
    if (!ended) ++never_ended;
  }
  printf("runs %d reports %lld false_ends %lld never_ended %lld\n", runs, total_reports, false_ends, never_ended);
  return (false_ends || never_ended) ? 1 : 0;
}
