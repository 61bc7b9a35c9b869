// Randomised interleaving model of the single-launch scheduler's work-list slots (kernels.cu: SolveState::gslot,
// pop_now / the feeder's pops, build_generation).  Test infrastructure; built and run by tests/test_genslot_model.py.
//
// Generation g's list lives in slot g & 3.  CTAs consume generations in order; a CTA that was busy with a long tile may
// still be at generation g-4 when the builder recycles its slot for generation g.
//   protocol "word"  (shipped):  one 64-bit word {count:32 | cursor:32} per slot; pop = ONE fetch_add that returns
//                                cursor and count of the same generation; the builder publishes with ONE exchange.
//   protocol "split" (round 1):  separate count and cursor words; pop = fetch_add(cursor) then load(count);
//                                the builder stores count, then cursor = 0 (two steps).
// Every step below is one atomic memory operation; a scheduler picks the next actor at random.  Property: every entry
// of every generation is handed out exactly once.  "word" must hold it, "split" must be caught violating it.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include <vector>

struct Cta {
  unsigned gen = 0;      // s_gen
  int busy = 0;          // remaining steps of the tile it is relaxing
  int phase = 0;         // split protocol: 0 = before fetch_add, 1 = cursor taken, count not yet read
  unsigned taken_i = 0;
};

int main(int argc, char** argv) {
  const bool split = argc > 1 && !strcmp(argv[1], "split");
  const int runs = argc > 2 ? atoi(argv[2]) : 2000;
  long long violations = 0, lost = 0;
  for (int run = 0; run < runs; ++run) {
    std::mt19937_64 rng(99 + run);
    const int ncta = 2 + (int)(rng() % 6);
    const unsigned ngen = 12 + (unsigned)(rng() % 20);
    std::vector<unsigned> count_of_gen(ngen);
    for (auto& c : count_of_gen) c = 1 + (unsigned)(rng() % 6);  // short lists: slots recycle quickly
    uint64_t word[4] = {0, 0, 0, 0};     // protocol "word"
    unsigned cnt[4] = {0, 0, 0, 0}, cur[4] = {0, 0, 0, 0};  // protocol "split"
    unsigned slot_gen[4] = {0, 0, 0, 0};  // which generation's entries a slot currently holds (model bookkeeping)
    unsigned published = 0;               // S->gen
    int build_phase = 0;                  // split: 0 idle, 1 count stored / cursor not yet reset
    unsigned building = 0;
    std::map<std::pair<unsigned, unsigned>, int> handed;
    // generation 0 is published before the launch
    word[0] = (uint64_t)count_of_gen[0] << 32; cnt[0] = count_of_gen[0]; cur[0] = 0; slot_gen[0] = 0;
    std::vector<Cta> cta(ncta);
    for (long long step = 0; step < 400000; ++step) {
      const int who = (int)(rng() % (ncta + 1));
      if (who == ncta) {  // ---- builder: a CTA that is AT the newest generation publishes its successor, early or late ----
        if (published + 1 >= ngen) continue;
        // (kernels.cu: "only its successor is ours to build" -- such a CTA has found every older list handed out,
        //  so the slot that generation published+1 recycles holds a list that is completely handed out)
        if (build_phase == 0 && cta[rng() % ncta].gen != published) continue;
        if (!split) {
          if (rng() % 3) continue;
          const unsigned g = published + 1;
          slot_gen[g & 3] = g;
          word[g & 3] = (uint64_t)count_of_gen[g] << 32;  // ONE exchange
          published = g;
        } else if (build_phase == 0) {
          if (rng() % 3) continue;
          building = published + 1;
          slot_gen[building & 3] = building;
          cnt[building & 3] = count_of_gen[building];     // store 1
          build_phase = 1;
        } else {
          cur[building & 3] = 0;                          // store 2
          published = building;
          build_phase = 0;
        }
        continue;
      }
      Cta& c = cta[who];
      if (c.busy > 0) { --c.busy; continue; }
      const unsigned s = c.gen & 3;
      if (!split) {
        const uint64_t w = word[s]; word[s] = w + 1;      // ONE fetch_add
        const unsigned i = (unsigned)w, n = (unsigned)(w >> 32);
        if (i < n) { handed[{slot_gen[s], i}] += 1; c.busy = (int)(rng() % (rng() % 8 == 0 ? 60 : 6)); }
        else if (published != c.gen) c.gen += 1;
      } else if (c.phase == 0) {
        c.taken_i = cur[s]; cur[s] += 1;                  // fetch_add(cursor)
        c.phase = 1;
      } else {
        const unsigned n = cnt[s];                        // load(count)
        c.phase = 0;
        if (c.taken_i < n) { handed[{slot_gen[s], c.taken_i}] += 1; c.busy = (int)(rng() % (rng() % 8 == 0 ? 60 : 6)); }
        else if (published != c.gen) c.gen += 1;
      }
    }
    for (auto& kv : handed) if (kv.second != 1) ++violations;
    for (unsigned g = 0; g <= published; ++g)
      for (unsigned i = 0; i < count_of_gen[g]; ++i) if (!handed.count({g, i})) ++lost;
  }
  printf("protocol %s runs %d double_handouts %lld never_handed_out %lld\n", split ? "split" : "word", runs, violations, lost);
  return (violations || lost) ? 1 : 0;
}
