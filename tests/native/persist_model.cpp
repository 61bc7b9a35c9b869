// Randomised interleaving model of the single-launch scheduler's termination protocol (kernels.cu: build_generation,
// the finisher's epilogue, SolveState::inflight / busy / done).  Test infrastructure; built and run by
// tests/test_persist_model.py.
//
// Tiles form a ring; relaxing a tile may wake (set the key of) itself and its neighbours a bounded number of times.
// CTAs pop tiles from the published list, relax them (several steps) and finish them in the kernel's order:
//     marks of the neighbours' keys -> busy[t] = 0 -> inflight -= 1
// One builder at a time (claimed like S->builder) turns the keys into the next list in the kernel's order:
//     read inflight  ->  scan keys (a key whose tile is busy is left pending, unless nothing was in flight)
//     ->  per selected tile: list entry, key = clean, busy = 1  ->  inflight += count, publish
//     or, nothing selected: done = 1 only if inflight read 0 BEFORE the scan (an early builder gives its claim back).
// Every line above is its own atomic step; a scheduler picks the next actor at random.  Properties:
//   * a tile is never relaxed by two CTAs at once;
//   * `done` is only ever announced when no key is pending and no tile is in flight or being relaxed;
//   * the run always ends.
// A mutant (`late`) reads inflight AFTER the scan instead of before it -- the ordering the kernel comments insist on --
// and must be caught announcing `done` too early.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

int main(int argc, char** argv) {
  const bool late = argc > 1 && !strcmp(argv[1], "late");
  const int runs = argc > 2 ? atoi(argv[2]) : 2000;
  long long early_done = 0, overlaps = 0, hung = 0;
  for (int run = 0; run < runs; ++run) {
    std::mt19937_64 rng(4242 + run);
    const int T = 4 + (int)(rng() % 20), ncta = 1 + (int)(rng() % 6);
    std::vector<char> key(T, 0), busy(T, 0);
    std::vector<int> running(T, 0);
    long long budget = 10 + (long long)(rng() % 300);  // wake-ups that may still be produced
    int inflight = 0, done = 0;
    std::vector<int> list;       // published entries not yet popped
    bool builder_claimed = false;
    // builder state machine
    int b_phase = 0, b_pos = 0, b_inflight = 0, b_owner = -1, b_first = -1;
    bool b_early = false;
    std::vector<int> b_sel;
    // CTA state: 0 idle, 1 relaxing (steps left), 2 marking (marks left), 3 clear busy, 4 decrement inflight
    struct Cta { int st = 0, tile = -1, left = 0; std::vector<int> marks; };
    std::vector<Cta> cta(ncta);
    key[rng() % T] = 1;  // the start point's tile
    bool ended = false;
    for (long long step = 0; step < 300000 && !ended; ++step) {
      const int who = (int)(rng() % ncta);
      Cta& c = cta[who];
      if (b_owner == who) {  // this CTA is building
        switch (b_phase) {
          case 0: if (!late) b_inflight = inflight; b_pos = 0; b_first = -1; b_sel.clear(); b_phase = 1; break;
          case 1:  // scan one key per step
            if (b_pos < T) {
              int k = key[b_pos];
              if (k && (late ? inflight != 0 : b_inflight != 0) && busy[b_pos]) k = 0;  // busy tiles stay pending
              if (k && b_first < 0) b_first = b_pos;
              if (k && (rng() % 4)) b_sel.push_back(b_pos);  // (bucket: not every pending key is within the threshold ...
              ++b_pos;
            } else {
              if (b_sel.empty() && b_first >= 0) b_sel.push_back(b_first);  // ... but the smallest one always is)
              if (late) b_inflight = inflight;  // MUTANT: inflight read after the scan
              b_phase = 2; b_pos = 0;
            }
            break;
          case 2:  // per selected tile: list entry, key = clean, busy = 1 (three steps folded into two)
            if (b_pos < (int)b_sel.size()) { key[b_sel[b_pos]] = 0; busy[b_sel[b_pos]] = 1; ++b_pos; }
            else b_phase = 3;
            break;
          case 3:
            if (!b_sel.empty()) { inflight += (int)b_sel.size(); for (int t : b_sel) list.push_back(t); }
            else if (!b_early && b_inflight == 0) {
              done = 1;
              bool clean = inflight == 0;
              for (int t = 0; t < T; ++t) clean = clean && !key[t] && !running[t];
              for (auto& o : cta) clean = clean && o.st == 0;
              if (!clean) ++early_done;
              ended = true;
            }
            builder_claimed = false; b_owner = -1; b_phase = 0;
            break;
        }
        continue;
      }
      switch (c.st) {
        case 0:
          if (!list.empty()) {
            c.tile = list.back(); list.pop_back();
            if (++running[c.tile] > 1) ++overlaps;
            c.st = 1; c.left = 1 + (int)(rng() % (rng() % 6 == 0 ? 40 : 5));
          } else if (!builder_claimed && (rng() % 2)) {
            builder_claimed = true; b_owner = who; b_phase = 0; b_early = false;
          }
          break;
        case 1:
          if (--c.left <= 0) {
            c.marks.clear();
            if (budget > 0 && (rng() % 3)) {
              const int n = 1 + (int)(rng() % 3);
              for (int k = 0; k < n && budget > 0; ++k, --budget) c.marks.push_back((c.tile + (int)(rng() % 3) - 1 + T) % T);
            }
            // an early build: this CTA popped the trigger entry and builds the next list before going on
            if (!builder_claimed && !list.empty() && (rng() % 5 == 0)) { /* early builder claimed below, after the finish */ }
            c.st = 2;
          }
          break;
        case 2:
          if (!c.marks.empty()) { key[c.marks.back()] = 1; c.marks.pop_back(); }
          else c.st = 3;
          break;
        case 3: busy[c.tile] = 0; --running[c.tile]; c.st = 4; break;
        case 4:
          --inflight; c.st = 0; c.tile = -1;
          if (!builder_claimed && !list.empty() && (rng() % 6 == 0)) {  // early builder (nobody waits for it)
            builder_claimed = true; b_owner = who; b_phase = 0; b_early = true;
          }
          break;
      }
    }
    if (!ended) ++hung;
  }
  printf("mode %s runs %d early_done %lld overlaps %lld hung %lld\n", late ? "late" : "shipped", runs, early_done, overlaps, hung);
  return (early_done || overlaps || hung) ? 1 : 0;
}
