/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (oracle/_ref).
 *
 * Builds the reference's own serial program *unmodified* into a shared library by
 * textually including it from where it lies under /root/reference (never copied
 * into this repo) with its main() renamed.  The harness then drives the
 * reference's sweepXYZ() (serial_new/sweep-tt-multistart.c:198-256) through the
 * reference's own file-scope globals fs[], start[], vbox, ttboxes[] (:60-66).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference leg may load the resulting oracle/_ref/libref_sweep.so.  The product
 * path (uoparallel_seismic_project_b200/) never touches it.
 *
 * Build: see oracle/Makefile (flags follow serial_new/Makefile:2: -O3, no -march,
 * no -ffast-math; -ffp-contract=off added defensively).
 */
#define main ref_serial_main
#include REF_SERIAL_SOURCE /* "/root/reference/serial_new/sweep-tt-multistart.c" */
#undef main

#include <string.h>

/* the reference's fixed capacities (serial_new/sweep-tt-multistart.c:41-44) */
int refh_fsmax(void) { return FSMAX; }
int refh_startmax(void) { return STARTMAX; }

/* Fill fs[] exactly as main() does at :120-128 (sqrt in double on an int, stored
 * to float, then multiplied by float delta = 10.0). */
int refh_set_star(const int *ijk, int starsize) {
  int i;
  float delta = 10.0;
  if (starsize > FSMAX) return 0;
  for (i = 0; i < starsize; i++) {
    fs[i].i = ijk[3 * i + 0];
    fs[i].j = ijk[3 * i + 1];
    fs[i].k = ijk[3 * i + 2];
    fs[i].d = sqrt(fs[i].i * fs[i].i + fs[i].j * fs[i].j + fs[i].k * fs[i].k);
    fs[i].d = delta * fs[i].d;
  }
  return 1;
}

float refh_star_distance(int l) { return fs[l].d; }

/* Velocity box: allocate through the reference's own vboxalloc and copy values. */
int refh_set_velocity(const float *v, int nx, int ny, int nz) {
  vboxfree(&vbox);
  if (!vboxalloc(&vbox, 0, 0, 0, nx, ny, nz)) return 0;
  memcpy(vbox.box.flat, v, sizeof(float) * (size_t)nx * ny * nz);
  return 1;
}

/* Load a .vbox through the reference's loader (include/velocityboxfiler.h:631). */
int refh_load_vbox(const char *path, int *dims, int *origin) {
  vboxfree(&vbox);
  if (!vbfileloadbinary(&vbox, path)) return 0;
  dims[0] = vbox.box.size.x; dims[1] = vbox.box.size.y; dims[2] = vbox.box.size.z;
  origin[0] = vbox.min.x; origin[1] = vbox.min.y; origin[2] = vbox.min.z;
  return 1;
}

/* Text dialect A through the reference's loader (velocityboxfiler.h:91). */
int refh_load_text(const char *path, int *dims, int *origin) {
  vboxfree(&vbox);
  if (!vbfileloadtext(&vbox, path)) return 0;
  dims[0] = vbox.box.size.x; dims[1] = vbox.box.size.y; dims[2] = vbox.box.size.z;
  origin[0] = vbox.min.x; origin[1] = vbox.min.y; origin[2] = vbox.min.z;
  return 1;
}

int refh_store_vbox(const char *path, int ox, int oy, int oz) {
  struct VELOCITYBOX out = vbox;
  out.min.x = ox; out.min.y = oy; out.min.z = oz;
  return vbfilestorebinary(path, out);
}

const float *refh_velocity_ptr(void) { return vbox.box.flat; }

/* Slot `s`: fresh travel-time box, INF everywhere, 0 at the start (:139-146). */
int refh_init_source(int s, int si, int sj, int sk) {
  int nx = vbox.box.size.x, ny = vbox.box.size.y, nz = vbox.box.size.z;
  if (s < 0 || s >= STARTMAX) return 0;
  boxfree(&ttboxes[s]);
  if (!boxalloc(&ttboxes[s], nx, ny, nz)) return 0;
  boxsetall(ttboxes[s], INFINITY);
  boxput(ttboxes[s], si, sj, sk, 0);
  start[s].i = si; start[s].j = sj; start[s].k = sk;
  return 1;
}

/* One call of the reference's sweepXYZ with the shipped arguments (:160). */
int refh_sweep_once(int s, int starsize) {
  return sweepXYZ(vbox.box.size.x, vbox.box.size.y, vbox.box.size.z, s, 0, starsize - 1);
}

/* Loop to convergence -- the documented intent of :151-170 (the shipped `break`
 * at :169 is marked TEMPORARY).  Returns the number of sweeps, including the
 * confirming one that changes nothing. */
int refh_solve(int s, int starsize, int maxsweeps, long long *total_changes) {
  int sweeps = 0, c;
  long long tot = 0;
  do {
    c = refh_sweep_once(s, starsize);
    tot += c;
    sweeps++;
  } while (c != 0 && (maxsweeps <= 0 || sweeps < maxsweeps));
  if (total_changes) *total_changes = tot;
  return sweeps;
}

void refh_copy_tt(int s, float *out) {
  memcpy(out, ttboxes[s].flat, sizeof(float) * boxvolume(ttboxes[s]));
}

void refh_set_tt(int s, const float *in) {
  memcpy(ttboxes[s].flat, in, sizeof(float) * boxvolume(ttboxes[s]));
}
