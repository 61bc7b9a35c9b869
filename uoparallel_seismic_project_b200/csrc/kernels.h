// kernels.h -- host-callable launchers of the sm_100a kernels (kernels.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "device_types.h"

namespace sweeptt {

struct TiledLaunch {
  int rxy;                   // halo template: 2, 4 or 7
  int stock_id;              // 0 = generic runtime-mask kernel, else a stock-star instantiation
  int nw;                    // compute warps per CTA (the star's columns of a tile are shared out between them)
  int grid;                  // persistent CTAs
  int grid_persistent;       // ... of the single-launch variant
  size_t smem_bytes;
};

// Shared-memory halo variant for a star with max(|i|,|j|) = r (2, 4 or 7), or 0 if none fits.
int tiled_variant_for_radius(int r);
// smem row/plane pitches of a variant (host needs them to precompute column offsets).
void tiled_variant_dims(int rxy, int* sxd, int* syd, int* szd);
// Occupancy-derived persistent grid + opt-in shared memory; returns cudaSuccess or an error.
cudaError_t tiled_prepare(int rxy, int stock_id, int device, TiledLaunch* out);
// Stock-star instantiation whose compile-time pattern list equals `masks` (ascending), or 0.
int tiled_stock_star_for(const uint32_t* masks_ascending, int n, int rxy_needed);

// Upload the column tables into __constant__ memory (stream ordered).
cudaError_t upload_star_constants(const ColumnDev* cols, int ncols, const float* col_hd, int nhd,
                                  const ExtraDev* extra, int nextra, const unsigned* pdesc, int npdesc,
                                  cudaStream_t stream);

struct RelaxArgs {
  BoxGeom g;
  const float* slow;           // padded slowness box
  float* tt;                   // nsrc padded travel-time boxes, contiguous
  int nsrc;
  const int* src_xyz;          // 3 ints per source (logical coords)
  SolveState* st;
  unsigned* worklist;          // 4 * cap entries: one list per generation in flight (round-based scheduling uses two)
  unsigned cap;
  unsigned* key;               // nsrc * ntiles activation keys: float bits of the smallest travel time
                               // that changed next to the tile since it was last relaxed; INF = clean
  unsigned* tmax;              // nsrc * ntiles: float bits of an upper bound of each tile's largest in-grid travel
                               // time (INF until the tile was relaxed with every node reached); nullptr = no filter
  float dmin;                  // lower bound (>= 0) of every edge delay fl(hd*fl(v_n+v_m)) of this model and star
  unsigned* busy;              // nsrc * ntiles: 1 while a tile is on a published list or being relaxed (single-launch
                               // scheduling only: such a tile is never put on a second list)
  unsigned* keysnap;           // nsrc * ntiles scratch words: key snapshot of the generation builder (large problems)
  float bin_scale;             // 32 / bucket (0 when bucket < 0): key -> sort bin of the generation builder
  float bucket;                // only tiles with key <= (smallest key) + bucket run in a round; <0 = all
  const unsigned long long* tile_pulls;  // per tile position: in-bounds pulls of one visit
  int ncols, nextra;
  int max_inner;               // in-tile relaxation passes per visit (>= 1)
  float neg_zero;              // -0.0f passed at run time (see mul2_exact in kernels.cu)
  unsigned lookahead;          // single launch: the list entry this far before the end triggers the next build
  unsigned trig_q8;            // ... but not before this fraction (in 1/256) of the list is handed out
  // ---- one grid spread over several devices (shared_grid.cu): the boxes are ONE virtual address range whose pages
  // live block-cyclically on the devices; every part relaxes the tiles of the x blocks it owns, reads halos from
  // peer memory with the same TMA loads and wakes the owner of a neighbour tile through that part's key array ----
  int nparts;                  // 0 or 1: a single context
  int part;                    // index of this context
  const unsigned char* tx_owner;    // [ntx] owner part of every tile column along x (blocks of a few tiles, dealt round-robin)
  const int* tx_slow0;              // [ntx] owned tiles: first plane of the tile's staged box in this part's local slowness copy
  int slow_pb;                 // != 0: `slow` / the slowness tensor map hold this part's blocks only, each with its halo
                               // planes (slow_pb planes per block, blocks in ascending order)
  unsigned* part_key[MAX_PARTS];    // every part's activation keys (peer memory)
  unsigned* part_tmax[MAX_PARTS];   // every part's per-tile upper bounds (peer memory; nullptr = no filter)
  unsigned* part_kmin[MAX_PARTS];   // every part's published smallest pending key: the activation bucket follows
                                    // the smallest key on ANY device, so no device runs far ahead of the front
  float front_slack;                // ... by more than this (travel-time units; a multiple of the bucket)
  int npat;                    // pattern groups (generic kernel; the stock kernels know theirs at compile time)
  int pat_begin[MAX_PATTERNS + 1];  // stock-star kernels: columns [pat_begin[p], pat_begin[p+1]) share pattern p
};

cudaError_t launch_relax_tiled(const TiledLaunch& tl, const CUtensorMap& tm_slow,
                               const CUtensorMap& tm_tt, const RelaxArgs& a, cudaStream_t stream);
// Single-launch solve: ONE persistent launch relaxes to the fixed point; the CTAs build the next work
// list ("generation") themselves as soon as the current one is handed out, so there is no round barrier,
// no per-round launch and no tail.  Needs launch_reset + launch_persist_begin on the same stream before it.
cudaError_t launch_relax_persistent(const TiledLaunch& tl, const CUtensorMap& tm_slow, const CUtensorMap& tm_tt,
                                    const RelaxArgs& a, cudaStream_t stream);
cudaError_t launch_persist_begin(const RelaxArgs& a, cudaStream_t stream);
cudaError_t launch_persist_check(const RelaxArgs& a, cudaStream_t stream);
size_t tiled_persistent_max_keys(int rxy);
// Two launches: (1) min-reduce the activation keys, (2) move every tile whose key is within the
// bucket of that minimum to the next work list (clearing its key), flip parity, advance the
// round and (when cond != 0) set the CUDA-graph WHILE condition to "work list not empty".
cudaError_t launch_compact(const RelaxArgs& a, unsigned long long cond, cudaStream_t stream);

// Simple (verification / fallback) path: one thread per node and source, global memory.
cudaError_t launch_relax_simple(const RelaxArgs& a, const StarDev* star, int nstar,
                                unsigned long long pulls_per_round, cudaStream_t stream);
cudaError_t launch_advance_simple(SolveState* st, unsigned long long cond, cudaStream_t stream);

// One grid over several devices: the tiles of this part around the start point go on its first work list (reset of
// the shared box itself is done per owned block by the host: launch_fill on plane ranges).
cudaError_t launch_reset_part(const RelaxArgs& a, int max_rounds, cudaStream_t stream);
cudaError_t launch_init_sources(const RelaxArgs& a, cudaStream_t stream);

// Fixed-point verifier.
cudaError_t launch_count_violations(const RelaxArgs& a, int source, const StarDev* star, int nstar,
                                    unsigned long long* out, cudaStream_t stream);

// Box utilities (device float-box pool).
cudaError_t launch_fill(float* p, long long n, float value, cudaStream_t stream);
cudaError_t launch_pad_box(const float* dense, float* padded, BoxGeom g, cudaStream_t stream);
cudaError_t launch_unpad_box(const float* padded, float* dense, BoxGeom g, cudaStream_t stream);
// tt := INF everywhere, 0 at each start; state, flags and the first work list reset.
cudaError_t launch_reset(const RelaxArgs& a, int max_rounds, cudaStream_t stream);
// Model statistics, 24 bytes (8-byte aligned): out6[0] = float bits of the smallest slowness of the dense staged
// model, out6[1] != 0 if any value is negative/NaN, then sum (double) and count (u64) of the finite values
cudaError_t launch_min_slowness(const float* dense, long long n, unsigned* out6, cudaStream_t stream);
// tmax[] := INF (call whenever travel times were overwritten from outside, e.g. sweeptt_put_tt)
cudaError_t launch_fill_tmax(const RelaxArgs& a, cudaStream_t stream);
cudaError_t launch_reset_state_only(SolveState* st, int max_rounds, cudaStream_t stream);

}  // namespace sweeptt
