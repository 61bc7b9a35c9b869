// worksim.c -- CPU model of the tile scheduler (activation keys, delta-stepping buckets, downwind filter) used to
// evaluate work-reduction policies OFFLINE (no GPU): how many (unit, column) evaluations a visit really needs.
//
// Policy under test ("useful-source window"): a tile is visited with activation key K = the smallest travel time that
// changed next to it since its key was last cleared.  A source node m can only matter for this visit if
//     tt[m] >= K                    (else it has not changed since the tile's previous visit)
// and fl(tt[m] + dmin) < umax       (else it cannot improve any node of the unit, umax = the unit's largest value).
// A star column (i,j) of a 4x8x8 unit is SKIPPED when none of the 32 x 24 staged source values behind it is useful.
// The model runs the solve twice (with and without skipping), checks that both fields are bit-identical and prints
// the fraction of column evaluations saved.  Scheduling is sequential (one tile at a time, in key order), which is a
// close stand-in for 148 concurrent CTAs.
//
// build: gcc -O3 -ffp-contract=off -o worksim worksim.c -lm
// usage: worksim slowness.f32 nx ny nz star.txt sx sy sz [bucket_factor]
// Three runs: all columns; column skipping by the useful-source window; per-unit dirty flags (all columns of the
// dirty units).  Every run must end in the same field.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef TX
#define TX 8
#endif
#define TY 8
#define TZ 8
#define R 7
#define XR ((7 + TX - 1) / TX)
#define ZH 8
static int nx, ny, nz, ntx, nty, ntz, nstar;
static float *slow, *tt;
static int *oi, *oj, *ok;
static float *hd;
static float *key, *tmaxv;
static unsigned char *udirty; static int use_units;
static float dmin_, bucket;
static unsigned *nstamp, *tvis, *tchg; static unsigned vclock;
static unsigned long long ex_col_need, ex_col_all, ex_pull_need, ex_pull_all, nb_col_need;

typedef struct { int i, j, nk, k[17]; float h[17]; } Col;
static Col cols[320];
static int ncols;

static inline size_t IDX(int x, int y, int z) { return ((size_t)x * ny + y) * nz + z; }
static inline float TT(int x, int y, int z) {
  if (x < 0 || y < 0 || z < 0 || x >= nx || y >= ny || z >= nz) return INFINITY;
  return tt[IDX(x, y, z)];
}
static inline float SL(int x, int y, int z) {
  if (x < 0 || y < 0 || z < 0 || x >= nx || y >= ny || z >= nz) return INFINITY;
  return slow[IDX(x, y, z)];
}

static unsigned long long cols_run, cols_all, unit_visits, unit_visits_empty, unit_changed, tile_visits, pulls_run, pulls_all;
static unsigned long long hist_frac[11];
static unsigned long long fine_pull_useful, fine_pull_all, fine_off_useful, fine_off_all, improving_pulls;

// returns 1 if anything changed; tmin_out = smallest lowered value
static int g_chmask;
static int visit(int tx, int ty, int tz, float K, int skip, float* tmin_out, int sxp, int syp, int szp) {
  static float sv[TX + 2 * R][TY + 2 * R][TZ + 2 * ZH], st[TX + 2 * R][TY + 2 * R][TZ + 2 * ZH];
  const int x0 = tx * TX, y0 = ty * TY, z0 = tz * TZ;
  for (int a = 0; a < TX + 2 * R; ++a)
    for (int b = 0; b < TY + 2 * R; ++b)
      for (int c = 0; c < TZ + 2 * ZH; ++c) {
        sv[a][b][c] = SL(x0 + a - R, y0 + b - R, z0 + c - ZH);
        st[a][b][c] = TT(x0 + a - R, y0 + b - R, z0 + c - ZH);
      }
  float out[TX][TY][TZ];
  int changed = 0;
  ++vclock;
  const int tself = (tx * nty + ty) * ntz + tz;
  const unsigned since = tvis[tself];
  // which neighbour tiles changed since our last visit
  int nbch[2 * XR + 1][3][3];
  for (int dx = -XR; dx <= XR; ++dx) for (int dy = -1; dy <= 1; ++dy) for (int dz = -1; dz <= 1; ++dz) {
    const int ux = tx + dx, uy = ty + dy, uz = tz + dz;
    nbch[dx + XR][dy + 1][dz + 1] = (ux >= 0 && ux < ntx && uy >= 0 && uy < nty && uz >= 0 && uz < ntz) ? tchg[(ux * nty + uy) * ntz + uz] > since : 0;
  }
  for (int u = 0; u < TX / 4; ++u) {
    if (x0 + 4 * u >= nx) continue;
    for (int ci = 0; ci < ncols; ++ci) {
      const Col* C = &cols[ci];
      int need = 0, nbneed = 0;
      for (int a = 0; a < 4; ++a) for (int b = 0; b < TY; ++b) for (int q = 0; q < C->nk; ++q) for (int c = 0; c < TZ; ++c) {
        const int lx = 4 * u + a + C->i, ly = b + C->j, lz = c + C->k[q];
        const int gx = x0 + lx, gy = y0 + ly, gz = z0 + lz;
        ex_pull_all++;
        if (gx < 0 || gy < 0 || gz < 0 || gx >= nx || gy >= ny || gz >= nz) continue;
        if (nstamp[IDX(gx, gy, gz)] > since) { need = 1; ex_pull_need++; }
        const int ddx = (lx + XR * TX) / TX, ddy = ly < 0 ? 0 : ly >= TY ? 2 : 1, ddz = lz < 0 ? 0 : lz >= TZ ? 2 : 1;
        if (nbch[ddx][ddy][ddz]) nbneed = 1;
      }
      ex_col_all++; ex_col_need += need; nb_col_need += nbneed;
    }
  }
  tvis[tself] = vclock;
  float tmin = INFINITY, tmx = 0.f;
  ++tile_visits;
  g_chmask = 0;
  const int tself2 = (tx * nty + ty) * ntz + tz;
  for (int u = 0; u < TX / 4; ++u) {
    if (x0 + 4 * u >= nx) continue;
    if (use_units && !udirty[2 * tself2 + u]) continue;
    udirty[2 * tself2 + u] = 0;
    ++unit_visits;
    // unit max over in-grid nodes
    float umax = 0.f;
    for (int a = 0; a < 4; ++a) for (int b = 0; b < TY; ++b) for (int c = 0; c < TZ; ++c) {
      const int gx = x0 + 4 * u + a, gy = y0 + b, gz = z0 + c;
      if (gx < nx && gy < ny && gz < nz) umax = fmaxf(umax, st[4 * u + a + R][b + R][c + ZH]);
    }
    // useful (x,y) bitmap of the staged box (full 24-value z window)
    unsigned rows[TX + 2 * R];
    for (int a = 0; a < TX + 2 * R; ++a) {
      unsigned w = 0;
      for (int b = 0; b < TY + 2 * R; ++b) {
        int useful = 0;
        for (int c = 0; c < TZ + 2 * ZH; ++c) {
          const float t = st[a][b][c];
          if (t >= K && (t + dmin_) < umax) { useful = 1; break; }
        }
        w |= (unsigned)useful << b;
      }
      rows[a] = w;
    }
    float acc[4][TY][TZ];
    for (int a = 0; a < 4; ++a) for (int b = 0; b < TY; ++b) for (int c = 0; c < TZ; ++c) acc[a][b][c] = st[4 * u + a + R][b + R][c + ZH];
    int ran = 0;
    for (int ci = 0; ci < ncols; ++ci) {
      const Col* C = &cols[ci];
      ++cols_all;
      pulls_all += 256ull * C->nk;
      if (skip) {
        unsigned w = 0;
        for (int a = 0; a < 4; ++a) w |= rows[4 * u + a + C->i + R];
        if (((w >> (C->j + R)) & 0xffu) == 0) continue;
      }
      ++cols_run; ++ran;
      pulls_run += 256ull * C->nk;
      if (skip) {
        for (int q = 0; q < C->nk; ++q) {
          int any = 0;
          for (int a = 0; a < 4; ++a) for (int b = 0; b < TY; ++b) for (int c = 0; c < TZ; ++c) {
            const float t = st[4 * u + a + R + C->i][b + R + C->j][c + ZH + C->k[q]];
            const int us = (t >= K && (t + dmin_) < umax);
            fine_pull_useful += us; any |= us;
          }
          fine_pull_all += 256; fine_off_all += 1; fine_off_useful += any;
        }
      }
      for (int a = 0; a < 4; ++a) for (int b = 0; b < TY; ++b) {
        const int xa = 4 * u + a + R, yb = b + R;
        for (int q = 0; q < C->nk; ++q) {
          const int k = C->k[q];
          const float h = C->h[q];
          for (int c = 0; c < TZ; ++c) {
            const float vn = sv[xa][yb][c + ZH], vm = sv[xa + C->i][yb + C->j][c + ZH + k], tm = st[xa + C->i][yb + C->j][c + ZH + k];
            const float cand = h * (vn + vm) + tm;
            if (cand < acc[a][b][c]) acc[a][b][c] = cand;
          }
        }
      }
    }
    if (ran == 0) ++unit_visits_empty;
    hist_frac[(ran * 10) / ncols]++;
    int uch = 0;
    for (int a = 0; a < 4; ++a) for (int b = 0; b < TY; ++b) for (int c = 0; c < TZ; ++c) {
      const int gx = x0 + 4 * u + a, gy = y0 + b, gz = z0 + c;
      if (gx >= nx || gy >= ny || gz >= nz) continue;
      float v = acc[a][b][c];
      if (gx == sxp && gy == syp && gz == szp) v = st[4 * u + a + R][b + R][c + ZH];
      if (v < st[4 * u + a + R][b + R][c + ZH]) { uch = 1; tmin = fminf(tmin, v); tt[IDX(gx, gy, gz)] = v; nstamp[IDX(gx, gy, gz)] = vclock; tchg[tself] = vclock; }
      tmx = fmaxf(tmx, v);
    }
    if (uch) { ++unit_changed; changed = 1; g_chmask |= 1 << u; }
  }
  tmaxv[(tx * nty + ty) * ntz + tz] = tmx;
  *tmin_out = tmin;
  return changed;
}

static int cmpf(const void* a, const void* b) {
  const float x = key[*(const int*)a], y = key[*(const int*)b];
  return (x > y) - (x < y);
}

static void solve(int skip, int sxp, int syp, int szp) {
  const int ntiles = ntx * nty * ntz;
  for (size_t i = 0; i < (size_t)nx * ny * nz; ++i) tt[i] = INFINITY;
  tt[IDX(sxp, syp, szp)] = 0.f;
  for (int i = 0; i < ntiles; ++i) { key[i] = INFINITY; tmaxv[i] = INFINITY; udirty[2 * i] = udirty[2 * i + 1] = 0; }
  for (int dx = -XR; dx <= XR; ++dx) for (int dy = -1; dy <= 1; ++dy) for (int dz = -1; dz <= 1; ++dz) {
    const int ux = sxp / TX + dx, uy = syp / TY + dy, uz = szp / TZ + dz;
    if (ux >= 0 && ux < ntx && uy >= 0 && uy < nty && uz >= 0 && uz < ntz) { key[(ux * nty + uy) * ntz + uz] = 0.f; udirty[2 * ((ux * nty + uy) * ntz + uz)] = udirty[2 * ((ux * nty + uy) * ntz + uz) + 1] = 1; }
  }
  memset(nstamp, 0, 4 * (size_t)nx * ny * nz); memset(tvis, 0, 4 * ntiles); memset(tchg, 0, 4 * ntiles); vclock = 1;
  nstamp[IDX(sxp, syp, szp)] = 1;
  for (int i = 0; i < ntiles; ++i) tchg[i] = 0;
  tchg[((sxp / TX) * nty + syp / TY) * ntz + szp / TZ] = 1;
  ex_col_need = ex_col_all = ex_pull_need = ex_pull_all = nb_col_need = 0;
  cols_run = cols_all = unit_visits = unit_visits_empty = unit_changed = tile_visits = pulls_run = pulls_all = 0;
  memset(hist_frac, 0, sizeof hist_frac);
  int* list = malloc(sizeof(int) * ntiles);
  float* kv = malloc(sizeof(float) * ntiles);
  int gens = 0;
  for (;;) {
    float kmin = INFINITY;
    for (int i = 0; i < ntiles; ++i) kmin = fminf(kmin, key[i]);
    if (kmin == INFINITY) break;
    ++gens;
    const float thr = kmin + bucket;
    int n = 0;
    for (int i = 0; i < ntiles; ++i) if (key[i] <= thr) list[n++] = i;
    qsort(list, n, sizeof(int), cmpf);
    for (int q = 0; q < n; ++q) { kv[q] = key[list[q]]; key[list[q]] = INFINITY; }
    for (int q = 0; q < n; ++q) {
      const int t = list[q];
      const int tz = t % ntz, ty = (t / ntz) % nty, tx = t / (ntz * nty);
      float tmin;
      if (!visit(tx, ty, tz, kv[q], skip, &tmin, sxp, syp, szp)) continue;
      for (int dx = -XR; dx <= XR; ++dx) for (int dy = -1; dy <= 1; ++dy) for (int dz = -1; dz <= 1; ++dz) {
        const int ux = tx + dx, uy = ty + dy, uz = tz + dz;
        if ((abs(dx) - 1) * TX >= R) continue;
        if (ux < 0 || ux >= ntx || uy < 0 || uy >= nty || uz < 0 || uz >= ntz) continue;
        const int u = (ux * nty + uy) * ntz + uz;
        const int self = !dx && !dy && !dz;
        if (!self && !((tmin + dmin_) < tmaxv[u])) continue;
        if (tmin < key[u]) key[u] = tmin;
        for (int cu = 0; cu < 2; ++cu) if (g_chmask & (1 << cu)) for (int vu = 0; vu < 2; ++vu) {
          const int U = 2 * tx + cu, V = 2 * ux + vu;
          if (abs(U - V) <= 2) udirty[2 * u + vu] = 1;
        }
      }
    }
  }
  free(list); free(kv);
  printf("skip=%d: generations %d, tile visits %llu (%.2f per tile), unit visits %llu (changed %llu, nothing to run %llu), "
         "columns run %llu of %llu = %.3f, pulls run %.3f G = %.2f grid-equivalents (all columns: %.2f)\n",
         skip, gens, tile_visits, (double)tile_visits / ntiles, unit_visits, unit_changed, unit_visits_empty, cols_run, cols_all,
         (double)cols_run / cols_all, pulls_run / 1e9, (double)pulls_run / ((double)nx * ny * nz * nstar),
         (double)pulls_all / ((double)nx * ny * nz * nstar));
  if (skip) printf("   within the columns that ran: useful (column,k) offsets %.3f, useful single pulls %.3f\n",
         (double)fine_off_useful / fine_off_all, (double)fine_pull_useful / fine_pull_all);
  printf("   ideal (sources changed since the previous visit): columns needed %.3f, single pulls needed %.3f; neighbour-tile mask: columns needed %.3f\n",
         (double)ex_col_need / ex_col_all, (double)ex_pull_need / ex_pull_all, (double)nb_col_need / ex_col_all);
  printf("   unit visits by fraction of columns run [0-10%%, ..., 100%%]:");
  for (int i = 0; i < 11; ++i) printf(" %llu", hist_frac[i]);
  printf("\n");
}

int main(int argc, char** argv) {
  if (argc < 9) { fprintf(stderr, "usage\n"); return 1; }
  nx = atoi(argv[2]); ny = atoi(argv[3]); nz = atoi(argv[4]);
  const size_t vol = (size_t)nx * ny * nz;
  slow = malloc(vol * 4); tt = malloc(vol * 4);
  FILE* f = fopen(argv[1], "rb");
  if (!f || fread(slow, 4, vol, f) != vol) { fprintf(stderr, "bad slowness file\n"); return 1; }
  fclose(f);
  f = fopen(argv[5], "r");
  if (!f || fscanf(f, "%d", &nstar) != 1) return 1;
  oi = malloc(4 * nstar); oj = malloc(4 * nstar); ok = malloc(4 * nstar); hd = malloc(4 * nstar);
  float hdmin = INFINITY, hdmax = 0;
  for (int l = 0; l < nstar; ++l) {
    if (fscanf(f, "%d %d %d", &oi[l], &oj[l], &ok[l]) != 3) return 1;
    hd[l] = 0.5f * (10.0f * (float)sqrt((double)(oi[l] * oi[l] + oj[l] * oj[l] + ok[l] * ok[l])));
    hdmin = fminf(hdmin, hd[l]); hdmax = fmaxf(hdmax, hd[l]);
  }
  fclose(f);
  // columns
  ncols = 0;
  for (int i = -R; i <= R; ++i) for (int j = -R; j <= R; ++j) {
    Col c; c.i = i; c.j = j; c.nk = 0;
    for (int k = -ZH; k <= ZH; ++k)
      for (int l = 0; l < nstar; ++l) if (oi[l] == i && oj[l] == j && ok[l] == k) { c.k[c.nk] = k; c.h[c.nk] = hd[l]; ++c.nk; break; }
    if (c.nk) cols[ncols++] = c;
  }
  const int sxp = atoi(argv[6]), syp = atoi(argv[7]), szp = atoi(argv[8]);
  const double factor = argc > 9 ? atof(argv[9]) : 2.0;
  double mean = 0; float vmin = INFINITY;
  for (size_t i = 0; i < vol; ++i) { mean += slow[i]; vmin = fminf(vmin, slow[i]); }
  mean /= (double)vol;
  bucket = (float)(factor * 2.0 * hdmax * mean);
  dmin_ = hdmin * (vmin + vmin);
  ntx = (nx + TX - 1) / TX; nty = (ny + TY - 1) / TY; ntz = (nz + TZ - 1) / TZ;
  key = malloc(4 * ntx * nty * ntz); tmaxv = malloc(4 * ntx * nty * ntz);
  udirty = malloc(2 * ntx * nty * ntz);
  nstamp = malloc(4 * vol); tvis = malloc(4 * ntx * nty * ntz); tchg = malloc(4 * ntx * nty * ntz);
  printf("%d x %d x %d, %d offsets in %d columns, bucket %.2f, dmin %.3f\n", nx, ny, nz, nstar, ncols, bucket, dmin_);
  use_units = 0;
  solve(0, sxp, syp, szp);
  float* ref = malloc(vol * 4);
  memcpy(ref, tt, vol * 4);
  solve(1, sxp, syp, szp);
  size_t nd0 = 0;
  for (size_t i = 0; i < vol; ++i) nd0 += memcmp(&ref[i], &tt[i], 4) != 0;
  // third policy: tile-level keys as before, plus a dirty flag per 4x8x8 UNIT (set by a changed unit within reach,
  // |unit index difference| <= 2); a visit relaxes only the dirty units
  use_units = 1;
  printf("per-unit dirty flags:\n");
  solve(0, sxp, syp, szp);
  if (nd0) { printf("column skipping changed %zu floats\n", nd0); return 1; }
  size_t nd = 0;
  for (size_t i = 0; i < vol; ++i) nd += memcmp(&ref[i], &tt[i], 4) != 0;
  printf("fields differ in %zu of %zu floats\n", nd, vol);
  return nd != 0;
}
