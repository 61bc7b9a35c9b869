/*
 * sweeptt.h -- C ABI of the B200-native multi-start forward-star sweep.
 *
 * This is the drop-in boundary for ONE hot path of
 * scrasmussen/uoparallel-seismic-project: "relax every grid node's travel time
 * against all offsets of a forward star until nothing changes, for every start
 * point".  Each entry point cites the reference interface it replaces
 * (paths relative to the reference checkout).
 *
 * Conventions (same polarity as the reference's loaders/allocators,
 * include/floatbox.h:118-120, include/velocityboxfiler.h:105-106):
 *   - functions returning int return NON-ZERO on success and 0 on failure;
 *     sweeptt_last_error() then holds a message (thread-local);
 *   - all boxes are FLOATBOX-ordered float32: index = (x*ny + y)*nz + z, z fastest
 *     (include/floatbox.h:127-129,160);
 *   - the caller owns every host buffer; the library owns all device memory;
 *   - there is NO CPU fallback: every compute entry point fails loudly when no
 *     CUDA device / sm_100 kernel image is available.
 *
 * Threading: a context is used by one host thread at a time.  Different contexts may be driven from different
 * threads; the forward-star tables live in per-device __constant__ memory, so contexts on ONE device whose stars
 * differ take turns (each compute call holds a lease on the device's tables for its duration; identical stars
 * share it).  Concurrent sweeptt_solve() calls for the same device serialise on that device's cached context.
 *
 * Plain C: no C++ or torch types cross this boundary.
 */
#ifndef SWEEPTT_H
#define SWEEPTT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* serial_new/sweep-tt-multistart.c:46-49 (and cuda/cudasweep-tt-multistart.cu:58-61):
 * forward-star offset and its distance d = delta * sqrt(i*i+j*j+k*k). Layout kept.
 * A translation unit that already defines these two structs itself -- the reference's own
 * programs do -- defines SWEEPTT_NO_STRUCTS before including this header (after its own
 * definitions) and passes its arrays as they are. */
#ifndef SWEEPTT_NO_STRUCTS
struct FS {
  int i, j, k;
  float d;
};

/* serial_new/sweep-tt-multistart.c:56-58: 0-based start point. Layout kept. */
struct START {
  int i, j, k;
};
#endif

/* which relaxation kernel runs (the default is the tiled sm_100a kernel) */
enum {
  SWEEPTT_KERNEL_AUTO = 0,   /* tiled TMA kernel when the star fits its halo, else simple */
  SWEEPTT_KERNEL_SIMPLE = 1, /* one thread per node, global memory (verification path)    */
  SWEEPTT_KERNEL_TILED = 2   /* force the tiled kernel (fails if the star does not fit)   */
};

/* convergence driver */
enum {
  SWEEPTT_LOOP_AUTO = 0,
  SWEEPTT_LOOP_BATCHED = 1, /* K rounds enqueued per host poll of the device flag          */
  SWEEPTT_LOOP_GRAPH = 2    /* device-resident loop (no host poll): ONE persistent launch that builds
                               its own work lists while the activation keys fit its shared memory,
                               else a CUDA graph with a device-evaluated WHILE node              */
};

typedef struct sweeptt_opts {
  int struct_size;     /* = sizeof(sweeptt_opts); lets the struct grow compatibly           */
  int device;          /* CUDA ordinal for single-device calls; -1 = current device         */
  int num_devices;     /* sweeptt_solve only: shard sources over devices [0,num_devices);
                          0 or 1 = one device                                               */
  int kernel;          /* SWEEPTT_KERNEL_*                                                  */
  int loop;            /* SWEEPTT_LOOP_*                                                    */
  int rounds_per_poll; /* batched loop: rounds enqueued between flag reads (0 = default 8)  */
  int max_rounds;      /* safety cap on relaxation rounds (0 = none)                        */
  int star_used;       /* number of leading star entries used as sweep centres' offsets;
                          0 = starsize-1, which is what the reference passes
                          (serial_new/sweep-tt-multistart.c:160)                            */
  int verbose;         /* >0: progress lines on stderr                                      */
  int slab_axis;       /* sweeptt_solve_slabs only: 0 = x (contiguous halos), 2 = z         */
  int profile_kernels; /* !=0: bracket every relaxation launch with CUDA events (batched
                          loop) so stats.relax_kernel_ms is measured, at ~1 us per round    */
} sweeptt_opts;

typedef struct sweeptt_stats {
  int struct_size;
  int rounds;                /* relaxation rounds executed (max over sources/devices)      */
  int kernel_used;           /* SWEEPTT_KERNEL_SIMPLE or _TILED                            */
  int devices_used;
  long long kernel_launches; /* launches of OUR kernels inside the solve                   */
  long long tile_visits;     /* tiles relaxed (tiled kernel)                               */
  long long relaxations;     /* in-bounds (node, offset, source) pull evaluations executed */
  double solve_ms;           /* device time of the solve (CUDA events on the solve stream) */
  double relax_kernel_ms;    /* of which: inside the relaxation kernel, summed over its
                                launches (only measured with opts.profile_kernels)         */
  long long relax_launches;  /* launches of the relaxation kernel                          */
  double h2d_ms, d2h_ms;     /* host<->device copies performed by the call                 */
  long long h2d_bytes, d2h_bytes;
  long long units_run;       /* (warp, tile) work units executed by the tiled kernel          */
  long long units_changed;   /* ... of which lowered at least one travel time                 */
} sweeptt_stats;

/* ---- discovery / errors ------------------------------------------------- */

/* cuda/cudasweep-tt-multistart.cu:187-201 (cudaGetDeviceCount + property print). */
int sweeptt_device_count(void);
/* Writes "name, SMs, clock kHz, smem/block optin" for `device`; returns non-zero on success. */
int sweeptt_device_info(int device, char *name, int name_len, int *sm_count, int *clock_khz,
                        size_t *smem_optin);
const char *sweeptt_last_error(void);
const char *sweeptt_version(void);

/* ---- forward star helpers ----------------------------------------------- */

/* serial_new/sweep-tt-multistart.c:120-128: fills fs[l].d = delta*(float)sqrt(i*i+j*j+k*k). */
void sweeptt_star_fill_distances(struct FS *fs, int starsize, float delta);

/* Host-side edge-set analysis (no GPU needed): turns the reference's two-sided
 * "centre relaxes itself and its neighbour" rule with its two quirks
 * (serial_new/...c:160 last offset unused; :219-221 centre==start skipped) into
 * the equivalent pull star.  Outputs (each may be NULL): offsets as (i,j,k)
 * triples, half distances, and a guard flag (1 = this pull is invalid when the
 * neighbour is the start point).  Returns the number of pull offsets, or -1. */
int sweeptt_build_pull_star(const struct FS *fs, int starsize, int star_used, int32_t *ijk_out,
                            float *half_d_out, int32_t *guard_out, int capacity);

/* Test hook (no device needed): how the tiled kernel shares the star's (i,j) columns out between `nw` warps.
 * Columns are grouped by k pattern; kmasks_out[c] is column c's pattern in group order; cuts holds six tables
 * of cut points, cuts[(t * ngroups + g) * (nw + 1) + p] = first column of part p's contiguous piece of group
 * g (entry nw = end of the group).  Returns the number of columns, or -1 if a capacity is too small. */
int sweeptt_debug_column_split(const struct FS *fs, int starsize, int nw, int *ngroups_out,
                               unsigned short *cuts, int cuts_capacity, unsigned *kmasks_out,
                               int kmasks_capacity);

/* ---- one-shot solve: host buffers in, host buffers out -------------------- */

/* Replaces `void cudaRun(int numstart, int starsize)` (cuda/cudasweep-tt-multistart.cu:80,227)
 * and the CPU loop `while(anychange) for s: sweepXYZ(...)` (serial_new/...c:151-170), with
 * explicit arguments instead of the file-scope globals fs[], start[], vbox, ttboxes[].
 * tt_out[s] (nx*ny*nz floats each) receives the converged field of source s; the library
 * initialises travel times itself (INF, 0 at the start; serial_new/...c:139-144).
 * With opts->num_devices > 1 the sources are sharded over the devices with no
 * inter-device communication (the mpi/backup.c:351-363 scheme).
 * More sources than one single-launch solve holds run as consecutive waves; the fields of a finished
 * wave are copied to tt_out[] while the next wave is relaxed.  tt_out[] may be page-locked
 * (sweeptt_host_alloc) or plain malloc memory like the reference's boxalloc; the latter is served
 * through a page-locked stop-over inside the library (the first box decides which way all boxes of a
 * call go; either way is correct for any mix). */
int sweeptt_solve(const float *slowness, int nx, int ny, int nz, const struct FS *fs, int starsize,
                  const struct START *starts, int numstart, float *const *tt_out,
                  const sweeptt_opts *opts, sweeptt_stats *stats);

/* Page-locked host memory for the boxes handed to sweeptt_solve (the counterpart of boxalloc's malloc,
 * include/floatbox.h:122-123): with such buffers the device->host copies of finished sources run at PCIe speed
 * BEHIND the relaxation of the remaining ones; pageable memory works too, only slower (the driver stages it).
 * Returns NULL on failure.  Release with sweeptt_host_free. */
void *sweeptt_host_alloc(size_t bytes);
void sweeptt_host_free(void *p);

/* sweeptt_solve keeps one cached context (device boxes, star tables) per device between
 * calls -- the device-resident float-box pool; this releases them. */
void sweeptt_release_cache(void);

/* ---- device-resident context (float-box pool) ---------------------------- */

/* A context owns one device's padded slowness box, a pool of padded travel-time
 * boxes, the star tables and the convergence state.  It is the device-resident
 * counterpart of the globals vbox/ttboxes[] plus boxalloc/boxsetall/boxput
 * (include/floatbox.h:114-199; device precedent cuda/floatbox.h:194-244). */
typedef struct sweeptt_ctx sweeptt_ctx;

sweeptt_ctx *sweeptt_create(const sweeptt_opts *opts);
void sweeptt_destroy(sweeptt_ctx *ctx);

/* Optional: run all of the context's work on a caller-provided cudaStream_t
 * (passed as void*), e.g. torch's current stream, so the caller's CUDA events see it. */
int sweeptt_set_stream(sweeptt_ctx *ctx, void *cuda_stream);

/* H2D of the slowness box into the padded device box (cuda/...cu:288-294). */
int sweeptt_set_model(sweeptt_ctx *ctx, const float *slowness, int nx, int ny, int nz);
/* Star tables into __constant__ memory (cuda/...cu:74,285). */
int sweeptt_set_star(sweeptt_ctx *ctx, const struct FS *fs, int starsize);
/* Takes numstart boxes from the pool (grows it if needed) and records the start points
 * (serial_new/...c:135-147). Out-of-range starts are an error (the reference would write
 * out of bounds). */
int sweeptt_set_sources(sweeptt_ctx *ctx, const struct START *starts, int numstart);

/* (Re)initialise the travel times on the device and relax to convergence.  Entirely
 * device-resident: no host buffer is touched.  Asynchronous errors surface here. */
int sweeptt_run(sweeptt_ctx *ctx, sweeptt_stats *stats);

/* Exactly `rounds` relaxation rounds without re-initialising (rounds >= 1); *changed
 * receives 0 when the last round changed nothing.  Used by tests and by profiling. */
int sweeptt_step(sweeptt_ctx *ctx, int rounds, int *changed, sweeptt_stats *stats);
int sweeptt_reset(sweeptt_ctx *ctx); /* INF / 0-at-start re-initialisation only */

/* D2H of one converged box, un-padded into FLOATBOX order (cuda/...cu:393-397). */
int sweeptt_get_tt(sweeptt_ctx *ctx, int source, float *tt_out);
/* H2D of a caller-provided state for source `source` (tests: arbitrary valid upper bounds). */
int sweeptt_put_tt(sweeptt_ctx *ctx, int source, const float *tt_in);

/* Device-side fixed-point verifier: counts (node, pull offset) pairs that would still
 * lower a travel time (the `testconvergence` invariant,
 * old/wavefront-openmp/wave-multistart.c:300-347, for this edge set). 0 <=> converged. */
int sweeptt_count_violations(sweeptt_ctx *ctx, int source, long long *violations);

/* In-bounds pull evaluations of one full round of one source (analytic). */
long long sweeptt_relaxations_per_round(sweeptt_ctx *ctx);
/* Bytes of device memory currently held by the context's pool. */
size_t sweeptt_pool_bytes(sweeptt_ctx *ctx);
/* Tiles (= activation keys) per source of the current model; the single-launch scheduler takes as many
 * sources per launch as fit its shared-memory key snapshot, more sources run as consecutive waves. */
long long sweeptt_tiles_per_source(sweeptt_ctx *ctx);

/* ---- single huge grid over the devices of one box ----------------------------------------- */

/* Replaces the MPI ghost-cell decomposition (mpi/16partsmpi.c:740-909,
 * mpi/sweep-tt-multistart.c:422-556; ghost width = star radius) for ONE source on ONE grid.
 * opts->num_devices = number of parts (more parts than visible devices share devices round-robin);
 * opts->slab_axis = the caller axis the grid is cut along.  The slowness box and the travel-time box
 * are each one virtual address range whose pages are dealt block-cyclically (blocks of a few tiles
 * along slab_axis) to the devices; every device relaxes the tiles of its own blocks, its TMA loads
 * read halo planes straight from the owner's memory over NVLink and a changed tile wakes its
 * neighbours through the owning device's activation keys.  There are no ghost copies and no exchange
 * step; the host only detects quiescence (opts->rounds_per_poll rounds per look, default 4).
 * The result is bit-identical to the single-device field.  stats->solve_ms is the host wall clock
 * of the relaxation phase, h2d_ms / d2h_ms those of upload (incl. set-up) and gather. */
int sweeptt_solve_slabs(const float *slowness, int nx, int ny, int nz, const struct FS *fs,
                        int starsize, struct START start, float *tt_out, const sweeptt_opts *opts,
                        sweeptt_stats *stats);

/* Same, but every part loads only the planes of its own blocks from the .vbox file with the subset
 * reader (include/velocityboxfiler.h:741) -- what mpi/16partsmpi.c would need for models that do not
 * fit one node's memory.  Model dimensions come from the file header. */
int sweeptt_solve_slabs_vbox(const char *vbox_path, const struct FS *fs, int starsize, struct START start,
                             float *tt_out, const sweeptt_opts *opts, sweeptt_stats *stats);

/* ---- file formats (drop-in surface of the CLI) ---------------------------- */

/* include/velocityboxfiler.h:631 vbfileloadbinary: *slowness is malloc'd (free with
 * sweeptt_free). Fails on bad magic, short file or checksum mismatch (:540-547,:727-732). */
int sweeptt_vbox_load(const char *path, float **slowness, int origin[3], int dims[3]);
/* include/velocityboxfiler.h:511 vbfileopenbinary: header only (dimensions of the stored box). */
int sweeptt_vbox_dims(const char *path, int dims[3]);
/* include/velocityboxfiler.h:310 vbfilestorebinary (byte-identical output). */
int sweeptt_vbox_store(const char *path, const float *slowness, const int origin[3],
                       const int dims[3]);
/* include/velocityboxfiler.h:741 vbfileloadbinarysubset via :511 vbfileopenbinary: reads the
 * sub-box of size sub_dims without reading the rest of the file (checksum not verified, :798-799).
 * sub_origin is FILE-RELATIVE: a 0-based index into the stored box, NOT the absolute coordinates
 * that the reference's bounds check compares with the file's min corner (:761-779, ox >= min);
 * a caller holding absolute coordinates subtracts the origin returned by sweeptt_vbox_load. */
int sweeptt_vbox_load_subset(const char *path, const int sub_origin[3], const int sub_dims[3],
                             float **slowness);
/* Text dialect A "x,y,z,v" per line (include/velocityboxfiler.h:91 vbfileloadtext) and
 * dialect B "nx ny nz" + bare floats (old/wavefront-openmp/wave-multistart.c:151-161);
 * sniffed by the comma on the first line. */
int sweeptt_text_load(const char *path, float **slowness, int origin[3], int dims[3]);
/* serial_new/sweep-tt-multistart.c:111-128: count, then "oi oj ok" rows; d filled with delta. */
int sweeptt_star_load(const char *path, float delta, struct FS **fs, int *starsize);
/* serial_new/sweep-tt-multistart.c:135-147: count, then "si sj sk" rows (0-based). */
int sweeptt_starts_load(const char *path, struct START **starts, int *numstart);
/* serial_new/sweep-tt-multistart.c:176-194: the output.tt text, byte-identical formatting. */
int sweeptt_write_output_tt(const char *path, const float *const *tt, int numstart, int nx, int ny,
                            int nz);
void sweeptt_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* SWEEPTT_H */
