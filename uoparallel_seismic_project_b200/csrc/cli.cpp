// cli.cpp -- `sweep-tt-multistart vfile fsfile startfile`: the reference's command line
// (serial_new/sweep-tt-multistart.c:12,70-195) on top of the C ABI.
//
// Same positional arguments, same stdout lines in the same order, same `output.tt` in the
// current directory.  Differences, all deliberate:
//   * the sweep loop runs to convergence on the GPU(s) (the reference's `break` after one
//     sweep at :169 is marked TEMPORARY); the per-sweep ">>> start s: changed == n" lines
//     are replaced by one summary per source because change counts are visiting-order
//     dependent;
//   * vfile may be a .vbox (serial_new) or either text dialect (cuda/, old/wavefront-openmp);
//   * no STARTMAX / FSMAX / MODELMAX limits; start points are range-checked.
// Optional environment: SWEEPTT_DEVICES=n (shard sources over n GPUs), SWEEPTT_SLABS=n (one grid in n
// slabs over the GPUs, single start point files only), SWEEPTT_TT_BIN=path
// (also dump raw float32 fields), SWEEPTT_NO_OUTPUT=1 (skip output.tt), SWEEPTT_KERNEL=simple.
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/sweeptt.h"

int main(int argc, char* argv[]) {
  if (argc < 4) {
    std::printf("usage: %s vfile fsfile startfile\n", argv[0]);
    return 1;
  }
  const char* vfile = argv[1];
  std::printf("Loading velocity model file: %s...", vfile);
  std::fflush(stdout);
  float* slow = nullptr;
  int origin[3], dims[3];
  bool loaded = false;
  {
    char magic[4] = {0, 0, 0, 0};
    if (FILE* f = std::fopen(vfile, "rb")) {
      size_t got = std::fread(magic, 1, 4, f);
      std::fclose(f);
      if (got == 4 && std::memcmp(magic, "vbox", 4) == 0)
        loaded = sweeptt_vbox_load(vfile, &slow, origin, dims) != 0;
      else
        loaded = sweeptt_text_load(vfile, &slow, origin, dims) != 0;
    }
  }
  if (!loaded) {
    std::fprintf(stderr, "%s\n", sweeptt_last_error());
    std::printf("Cannot open velocity model file: %s\n", vfile);
    return 1;
  }
  const int nx = dims[0], ny = dims[1], nz = dims[2];
  std::printf(" done.\n");
  std::printf("Velocity model dimensions: %d x %d x %d\n", nx, ny, nz);

  const float delta = 10.0;
  struct FS* fs = nullptr;
  int starsize = 0;
  if (!sweeptt_star_load(argv[2], delta, &fs, &starsize)) {
    std::fprintf(stderr, "%s\n", sweeptt_last_error());
    std::printf("Cannot open forward star offset file: %s\n", argv[2]);
    return 1;
  }
  std::printf("Forward star offset file: %s\n", argv[2]);
  struct START* starts = nullptr;
  int numstart = 0;
  if (!sweeptt_starts_load(argv[3], &starts, &numstart)) {
    std::fprintf(stderr, "%s\n", sweeptt_last_error());
    std::printf("Cannot open starting points file: %s\n", argv[3]);
    return 1;
  }
  std::printf("Starting points file: %s\n", argv[3]);
  std::printf("Delta: %f\n", delta);
  std::printf("Forward star size: %d\n", starsize);
  {  // the numradius / fsindex bookkeeping main() prints (:116-132)
    const int FSRADIUSMAX = 7;
    int fsindex[FSRADIUSMAX] = {0, 0, 0, 0, 0, 0, 0}, numradius = 0;
    for (int l = 0; l < starsize; ++l) {
      const float d = fs[l].d / delta;
      if ((numradius + 1) < d) {
        if (numradius < FSRADIUSMAX) fsindex[numradius] = l;
        numradius++;
      }
    }
    std::printf("Forward star offsets read\n");
    for (int r = 0; r < FSRADIUSMAX; ++r) std::printf("numradius: %d, fsindex[%d]: %d\n", numradius, r, fsindex[r]);
  }
  for (int s = 0; s < numstart; ++s) std::printf("starting point %d: %d %d %d\n", s, starts[s].i, starts[s].j, starts[s].k);
  std::printf("Starting points read\n");

  const size_t vol = (size_t)nx * ny * nz;
  std::vector<float*> tt(numstart);
  std::vector<char> pinned(numstart, 0);
  for (int s = 0; s < numstart; ++s) {
    // boxalloc (include/floatbox.h:122-123) -- page-locked when possible, so the copies back overlap the solve
    // (page-locking costs ~1 s per GB on the test box: worth it for a few boxes, not for config 3's 1.3 GB)
    tt[s] = (vol * sizeof(float) * (size_t)numstart <= (size_t(512) << 20)) ? static_cast<float*>(sweeptt_host_alloc(vol * sizeof(float))) : nullptr;
    pinned[s] = tt[s] != nullptr;
    if (!tt[s]) tt[s] = static_cast<float*>(std::malloc(vol * sizeof(float)));
    if (!tt[s]) {
      std::printf("out of memory for travel time volume %d\n", s);
      return 1;
    }
  }
  sweeptt_opts opts;
  std::memset(&opts, 0, sizeof opts);
  opts.struct_size = sizeof opts;
  opts.device = -1;
  if (const char* e = std::getenv("SWEEPTT_DEVICES")) opts.num_devices = std::atoi(e);
  if (const char* e = std::getenv("SWEEPTT_KERNEL")) opts.kernel = !std::strcmp(e, "simple") ? SWEEPTT_KERNEL_SIMPLE : SWEEPTT_KERNEL_AUTO;
  sweeptt_stats st;
  const auto t_solve0 = std::chrono::steady_clock::now();
  std::printf("sweep 1 begin\n");
  std::fflush(stdout);
  int ok;
  const char* slabs = std::getenv("SWEEPTT_SLABS");
  if (slabs && std::atoi(slabs) > 1) {
    // one huge grid: slab decomposition, one start point at a time (mpi/16partsmpi.c scheme)
    opts.num_devices = std::atoi(slabs);
    if (const char* ax = std::getenv("SWEEPTT_SLAB_AXIS")) opts.slab_axis = std::atoi(ax);
    ok = 1;
    for (int s = 0; s < numstart && ok; ++s) ok = sweeptt_solve_slabs(slow, nx, ny, nz, fs, starsize, starts[s], tt[s], &opts, &st);
  } else {
    ok = sweeptt_solve(slow, nx, ny, nz, fs, starsize, starts, numstart, tt.data(), &opts, &st);
  }
  if (!ok) {
    std::printf("sweep failed: %s\n", sweeptt_last_error());
    return 1;
  }
  for (int s = 0; s < numstart; ++s) std::printf(">>> start %d: converged\n", s);
  std::printf("sweep %d finished: anychange = 0\n", st.rounds);
  std::printf("[sweeptt] %d GPU(s), %d rounds, %.3f ms, %.3f GRelax, %.1f GRelax/s\n", st.devices_used, st.rounds,
              st.solve_ms, st.relaxations * 1e-9, st.solve_ms > 0 ? st.relaxations * 1e-6 / st.solve_ms : 0.0);

  if (const char* bin = std::getenv("SWEEPTT_TT_BIN")) {
    if (FILE* f = std::fopen(bin, "wb")) {
      for (int s = 0; s < numstart; ++s) std::fwrite(tt[s], sizeof(float), vol, f);
      std::fclose(f);
    }
  }
  const auto t_solve1 = std::chrono::steady_clock::now();
  if (!std::getenv("SWEEPTT_NO_OUTPUT")) {
    if (!sweeptt_write_output_tt("output.tt", tt.data(), numstart, nx, ny, nz)) {
      std::printf("Can not open travel time output file: %s\n", "output.tt");
      return 1;
    }
  }
  if (std::getenv("SWEEPTT_TIMING")) {  // phase times of the whole command (stderr: stdout keeps the reference's lines)
    const auto t_end = std::chrono::steady_clock::now();
    auto sec = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
      return std::chrono::duration<double>(b - a).count();
    };
    std::fprintf(stderr, "[sweeptt-timing] solve_call_s=%.4f output_tt_s=%.4f sources=%d nodes=%zu\n", sec(t_solve0, t_solve1),
                 sec(t_solve1, t_end), numstart, vol);
  }
  for (int s = 0; s < numstart; ++s) {
    if (pinned[s]) sweeptt_host_free(tt[s]); else std::free(tt[s]);
  }
  sweeptt_free(slow);
  sweeptt_free(fs);
  sweeptt_free(starts);
  sweeptt_release_cache();
  return 0;
}
