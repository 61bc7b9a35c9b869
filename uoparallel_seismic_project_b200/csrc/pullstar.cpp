// pullstar.cpp -- see pullstar.h.  Pure host code, no CUDA.
#include "pullstar.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>

namespace sweeptt {

namespace {
struct Key {
  int i, j, k;
  uint32_t dbits;
  bool operator<(const Key& o) const {
    return std::tie(i, j, k, dbits) < std::tie(o.i, o.j, o.k, o.dbits);
  }
};
uint32_t bits_of(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return u;
}
}  // namespace

PullStar build_pull_star(const FS* fs, int starsize, int star_used) {
  PullStar ps;
  if (star_used <= 0 || star_used > starsize) star_used = starsize - 1;
  // (offset, distance) -> is the pull supplied unconditionally (rule 1) / only by the
  // neighbour's own visit (rule 2)?
  //   rule 1: centre n, star entry l: n may take delay + tt[n+o_l]       (:233-243), n != start
  //   rule 2: centre m, star entry l: neighbour n = m+o_l may take delay + tt[m] (:228-232,
  //           :245-248), i.e. n pulls over offset -o_l, valid only while m != start (:219-221)
  std::map<Key, int> have;  // bit0 = rule 1, bit1 = rule 2
  for (int l = 0; l < star_used; ++l) {
    const FS& e = fs[l];
    if (e.i == 0 && e.j == 0 && e.k == 0) continue;  // self edge: can never change anything
    have[Key{e.i, e.j, e.k, bits_of(e.d)}] |= 1;
    have[Key{-e.i, -e.j, -e.k, bits_of(e.d)}] |= 2;
  }
  std::vector<PullOffset> plain;
  for (const auto& kv : have) {
    PullOffset p;
    p.i = kv.first.i; p.j = kv.first.j; p.k = kv.first.k;
    float d;
    std::memcpy(&d, &kv.first.dbits, 4);
    p.hd = d * 0.5f;
    p.guarded = (kv.second & 1) ? 0 : 1;  // rule 1 subsumes rule 2 (tt[start] is never updated)
    ps.rx = std::max(ps.rx, std::abs(p.i));
    ps.ry = std::max(ps.ry, std::abs(p.j));
    ps.rz = std::max(ps.rz, std::abs(p.k));
    if (p.guarded) ps.extra.push_back(p); else plain.push_back(p);
  }
  // Column grouping of the plain pulls.  A second plain pull with the same (i,j,k) but a
  // different distance (only possible for hand-made stars) goes to the one-at-a-time list.
  std::map<std::pair<int, int>, std::vector<PullOffset>> cols;
  std::vector<PullOffset> grouped;
  for (const auto& p : plain) {
    auto& col = cols[{p.i, p.j}];
    bool dup = false;
    for (const auto& q : col) dup |= (q.k == p.k);
    if (dup || std::abs(p.k) > KHALO) ps.extra.push_back(p); else col.push_back(p);
  }
  for (auto& kv : cols) {
    auto& col = kv.second;
    if (col.empty()) continue;
    std::sort(col.begin(), col.end(), [](const PullOffset& a, const PullOffset& b) { return a.k < b.k; });
    PullColumn c;
    c.i = kv.first.first; c.j = kv.first.second; c.kmask = 0;
    c.hd_begin = (int)ps.col_hd.size();
    for (const auto& p : col) {
      c.kmask |= 1u << (p.k + KHALO);
      ps.col_hd.push_back(p.hd);
      grouped.push_back(p);
    }
    ps.columns.push_back(c);
  }
  ps.all = grouped;
  ps.all.insert(ps.all.end(), ps.extra.begin(), ps.extra.end());
  return ps;
}

long long count_pulls(const PullStar& ps, int nx, int ny, int nz, int x0, int x1, int y0, int y1,
                      int z0, int z1) {
  auto span = [](int lo, int hi, int o, int n) -> long long {
    // #{c in [lo,hi) : 0 <= c+o < n}
    int a = std::max(lo, -o), b = std::min(hi, n - o);
    return b > a ? (long long)(b - a) : 0;
  };
  x1 = std::min(x1, nx); y1 = std::min(y1, ny); z1 = std::min(z1, nz);
  long long total = 0;
  for (const auto& p : ps.all)
    total += span(x0, x1, p.i, nx) * span(y0, y1, p.j, ny) * span(z0, z1, p.k, nz);
  return total;
}

}  // namespace sweeptt
