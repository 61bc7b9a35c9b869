// hostio.cpp -- file formats of the drop-in surface (no CUDA here).
//
//   .vbox binary     formats/VBOXFORMAT.txt:26-44; include/velocityboxfiler.h:310-452 (store),
//                    :511-737 (load), :741-864 (subset load)
//   text dialect A   "x,y,z,v" per line, include/velocityboxfiler.h:91-219
//   text dialect B   "nx ny nz" + bare floats, old/wavefront-openmp/wave-multistart.c:151-161
//   forward star     serial_new/sweep-tt-multistart.c:111-128
//   start points     serial_new/sweep-tt-multistart.c:135-147
//   output.tt        serial_new/sweep-tt-multistart.c:176-194
#include <cerrno>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sweeptt.h"

// error plumbing shared with solver.cu
extern "C" const char* sweeptt_last_error(void);
namespace sweeptt { int set_error(const char* fmt, ...); }

namespace {

// The reference sums every 4-byte word with each byte taken as a SIGNED char before it is
// shifted into place (include/velocityboxfiler.h:78-83,241-252).  In two's complement that is
// word - 0x100*[b0<0] - 0x10000*[b1<0] - 0x1000000*[b2<0]  (mod 2^32).
inline uint32_t vbox_word_sum(uint32_t w) {
  uint32_t s = w;
  if (w & 0x00000080u) s -= 0x00000100u;
  if (w & 0x00008000u) s -= 0x00010000u;
  if (w & 0x00800000u) s -= 0x01000000u;
  return s;
}
uint32_t vbox_checksum(const void* data, size_t nwords, uint32_t seed) {
  const unsigned char* p = static_cast<const unsigned char*>(data);
  uint32_t sum = seed;
  for (size_t i = 0; i < nwords; ++i) {
    uint32_t w;
    std::memcpy(&w, p + 4 * i, 4);
    sum += vbox_word_sum(w);
  }
  return sum;
}
bool host_is_little_endian() {
  const uint32_t one = 1;
  return *reinterpret_cast<const unsigned char*>(&one) == 1;
}
void byteswap_words(void* data, size_t nwords) {
  unsigned char* p = static_cast<unsigned char*>(data);
  for (size_t i = 0; i < nwords; ++i, p += 4) {
    std::swap(p[0], p[3]);
    std::swap(p[1], p[2]);
  }
}

struct VboxHeader {
  int32_t origin[3];
  int32_t dims[3];
  uint32_t checksum_so_far;
  long datapos;
};

// include/velocityboxfiler.h:511-616 (vbfileopenbinary)
FILE* open_vbox(const char* path, VboxHeader* h) {
  FILE* f = std::fopen(path, "rb");
  if (!f) {
    sweeptt::set_error("error opening file %s", path);
    return nullptr;
  }
  unsigned char raw[28];
  if (std::fread(raw, 1, 28, f) != 28 || std::memcmp(raw, "vbox", 4) != 0) {
    sweeptt::set_error("input file %s is not a vbox binary file, or is corrupted", path);
    std::fclose(f);
    return nullptr;
  }
  h->checksum_so_far = vbox_checksum(raw, 7, 0);  // checksum runs over the bytes as stored
  if (!host_is_little_endian()) byteswap_words(raw + 4, 6);
  std::memcpy(h->origin, raw + 4, 12);
  std::memcpy(h->dims, raw + 16, 12);
  h->datapos = 28;
  if (h->dims[0] <= 0 || h->dims[1] <= 0 || h->dims[2] <= 0) {
    sweeptt::set_error("error reading header in %s: suspect corruption", path);
    std::fclose(f);
    return nullptr;
  }
  return f;
}

}  // namespace

extern "C" void sweeptt_free(void* p) { std::free(p); }

extern "C" int sweeptt_vbox_load(const char* path, float** slowness, int origin[3], int dims[3]) {
  if (!path || !slowness) return sweeptt::set_error("sweeptt_vbox_load: null argument");
  VboxHeader h;
  FILE* f = open_vbox(path, &h);
  if (!f) return 0;
  const size_t vol = (size_t)h.dims[0] * h.dims[1] * h.dims[2];
  float* v = static_cast<float*>(std::malloc(vol * 4));
  if (!v) {
    std::fclose(f);
    return sweeptt::set_error("unable to allocate memory for a VELOCITYBOX with dimension: %d x %d x %d", h.dims[0],
                              h.dims[1], h.dims[2]);
  }
  uint32_t stored = 0;
  const bool ok = std::fread(v, 4, vol, f) == vol && std::fread(&stored, 4, 1, f) == 1;
  std::fclose(f);
  if (!ok) {
    std::free(v);
    return sweeptt::set_error("error reading value from %s (file too short)", path);
  }
  const uint32_t sum = vbox_checksum(v, vol, h.checksum_so_far);
  if (!host_is_little_endian()) {
    byteswap_words(v, vol);
    byteswap_words(&stored, 1);
  }
  if (sum != stored) {  // include/velocityboxfiler.h:727-732
    std::free(v);
    return sweeptt::set_error("checksum mismatch in input file %s: suspect corruption", path);
  }
  *slowness = v;
  for (int a = 0; a < 3; ++a) {
    if (origin) origin[a] = h.origin[a];
    if (dims) dims[a] = h.dims[a];
  }
  return 1;
}

// header only: include/velocityboxfiler.h:511-616 (vbfileopenbinary)
extern "C" int sweeptt_vbox_dims(const char* path, int dims[3]) {
  if (!path || !dims) return sweeptt::set_error("sweeptt_vbox_dims: null argument");
  VboxHeader h;
  FILE* f = open_vbox(path, &h);
  if (!f) return 0;
  std::fclose(f);
  for (int a = 0; a < 3; ++a) dims[a] = h.dims[a];
  return 1;
}

extern "C" int sweeptt_vbox_store(const char* path, const float* slowness, const int origin[3], const int dims[3]) {
  if (!path || !slowness || !dims) return sweeptt::set_error("sweeptt_vbox_store: null argument");
  FILE* f = std::fopen(path, "wb");
  if (!f) return sweeptt::set_error("error creating file %s", path);
  int32_t head[7];
  std::memcpy(&head[0], "vbox", 4);
  for (int a = 0; a < 3; ++a) {
    head[1 + a] = origin ? origin[a] : 0;
    head[4 + a] = dims[a];
  }
  const size_t vol = (size_t)dims[0] * dims[1] * dims[2];
  bool ok;
  uint32_t sum;
  if (host_is_little_endian()) {
    sum = vbox_checksum(head, 7, 0);
    sum = vbox_checksum(slowness, vol, sum);
    ok = std::fwrite(head, 4, 7, f) == 7 && std::fwrite(slowness, 4, vol, f) == vol && std::fwrite(&sum, 4, 1, f) == 1;
  } else {  // include/velocityboxfiler.h:401-446: reverse every word after the magic
    byteswap_words(head + 1, 6);
    std::vector<float> tmp(slowness, slowness + vol);
    byteswap_words(tmp.data(), vol);
    sum = vbox_checksum(head, 7, 0);
    sum = vbox_checksum(tmp.data(), vol, sum);
    uint32_t s2 = sum;
    byteswap_words(&s2, 1);
    ok = std::fwrite(head, 4, 7, f) == 7 && std::fwrite(tmp.data(), 4, vol, f) == vol && std::fwrite(&s2, 4, 1, f) == 1;
  }
  ok = (std::fclose(f) == 0) && ok;
  return ok ? 1 : sweeptt::set_error("error writing to file %s", path);
}

extern "C" int sweeptt_vbox_load_subset(const char* path, const int so[3], const int sd[3], float** slowness) {
  if (!path || !so || !sd || !slowness) return sweeptt::set_error("sweeptt_vbox_load_subset: null argument");
  VboxHeader h;
  FILE* f = open_vbox(path, &h);
  if (!f) return 0;
  for (int a = 0; a < 3; ++a)
    if (so[a] < 0 || sd[a] <= 0 || sd[a] > h.dims[a] - so[a]) {  // include/velocityboxfiler.h:761-779
      std::fclose(f);
      return sweeptt::set_error("file %s doesn't contain the requested subset", path);
    }
  const size_t vol = (size_t)sd[0] * sd[1] * sd[2];
  float* v = static_cast<float*>(std::malloc(vol * 4));
  if (!v) {
    std::fclose(f);
    return sweeptt::set_error("unable to allocate memory for the subset");
  }
  // One seek + one read per contiguous run; checksum not verified (include/velocityboxfiler.h:798-826 reads z strip
  // by z strip).  Strips that span the stored z range are contiguous over y, planes that also span the stored y range
  // are contiguous over x: a slab of whole planes -- what one part of a multi-device grid loads -- is ONE read.
  bool ok = true;
  const bool full_z = so[2] == 0 && sd[2] == h.dims[2];
  const bool full_y = full_z && so[1] == 0 && sd[1] == h.dims[1];
  const size_t run = full_y ? vol : full_z ? (size_t)sd[1] * sd[2] : (size_t)sd[2];   // floats per read
  const int nx_runs = full_y ? 1 : sd[0], ny_runs = full_z ? 1 : sd[1];
  for (int x = 0; x < nx_runs && ok; ++x)
    for (int y = 0; y < ny_runs && ok; ++y) {
      const long long pos = h.datapos + 4LL * (((long long)(x + so[0]) * h.dims[1] + (y + so[1])) * h.dims[2] + so[2]);
      ok = fseeko(f, (off_t)pos, SEEK_SET) == 0 &&
           std::fread(v + ((size_t)x * sd[1] + y) * sd[2], 4, run, f) == run;
    }
  std::fclose(f);
  if (!ok) {
    std::free(v);
    return sweeptt::set_error("error reading subset from %s", path);
  }
  if (!host_is_little_endian()) byteswap_words(v, vol);
  *slowness = v;
  return 1;
}

extern "C" int sweeptt_text_load(const char* path, float** slowness, int origin[3], int dims[3]) {
  if (!path || !slowness) return sweeptt::set_error("sweeptt_text_load: null argument");
  FILE* f = std::fopen(path, "rb");
  if (!f) return sweeptt::set_error("error opening file %s", path);
  std::fseek(f, 0, SEEK_END);
  const long size = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::string buf((size_t)std::max(0L, size), '\0');
  const bool rd = size > 0 && std::fread(&buf[0], 1, (size_t)size, f) == (size_t)size;
  std::fclose(f);
  if (!rd) return sweeptt::set_error("error reading file %s", path);
  const size_t eol = buf.find('\n');
  const bool dialect_a = buf.substr(0, eol).find(',') != std::string::npos;
  int o[3] = {0, 0, 0}, n[3];
  const char* p = buf.c_str();
  char* end = nullptr;
  float* v = nullptr;
  size_t vol = 0;
  if (dialect_a) {
    // first line gives the origin, last line the far corner (include/velocityboxfiler.h:126,136-175)
    if (std::sscanf(p, "%d,%d,%d", &o[0], &o[1], &o[2]) != 3)
      return sweeptt::set_error("error reading first line from file %s", path);
    size_t last = buf.find_last_not_of("\r\n \t");
    if (last == std::string::npos) return sweeptt::set_error("error scanning for last line in file %s", path);
    size_t bol = buf.find_last_of("\r\n", last);
    bol = (bol == std::string::npos) ? 0 : bol + 1;
    int x, y, z;
    if (std::sscanf(p + bol, "%d,%d,%d", &x, &y, &z) != 3)
      return sweeptt::set_error("error reading last line from file %s", path);
    n[0] = x - o[0] + 1; n[1] = y - o[1] + 1; n[2] = z - o[2] + 1;
    if (n[0] <= 0 || n[1] <= 0 || n[2] <= 0) return sweeptt::set_error("nonsense coordinates in file %s", path);
    vol = (size_t)n[0] * n[1] * n[2];
    v = static_cast<float*>(std::malloc(vol * 4));
    if (!v) return sweeptt::set_error("unable to allocate memory for %d x %d x %d", n[0], n[1], n[2]);
    // values are taken in file order; the coordinates on each line are not re-checked (:196-212)
    for (size_t l = 0; l < vol; ++l) {
      bool ok = true;
      for (int k = 0; k < 3 && ok; ++k) {
        std::strtol(p, &end, 10);
        ok = end != p && *end == ',';
        p = end + 1;
      }
      if (ok) {
        v[l] = std::strtof(p, &end);
        ok = end != p;
        p = end;
      }
      if (!ok) {
        std::free(v);
        return sweeptt::set_error("I am confused by line %zu in %s", l + 1, path);
      }
    }
  } else {
    for (int k = 0; k < 3; ++k) {
      n[k] = (int)std::strtol(p, &end, 10);
      if (end == p || n[k] <= 0) return sweeptt::set_error("bad 'nx ny nz' header in %s", path);
      p = end;
    }
    vol = (size_t)n[0] * n[1] * n[2];
    v = static_cast<float*>(std::malloc(vol * 4));
    if (!v) return sweeptt::set_error("unable to allocate memory for %d x %d x %d", n[0], n[1], n[2]);
    for (size_t l = 0; l < vol; ++l) {
      v[l] = std::strtof(p, &end);
      if (end == p) {
        std::free(v);
        return sweeptt::set_error("value %zu missing in %s", l + 1, path);
      }
      p = end;
    }
  }
  *slowness = v;
  for (int a = 0; a < 3; ++a) {
    if (origin) origin[a] = o[a];
    if (dims) dims[a] = n[a];
  }
  return 1;
}

// whitespace-separated ints with C "%i" semantics (decimal, 0x.., 0..), like the reference's fscanf
static bool read_ints(const char* path, std::vector<long>& out) {
  FILE* f = std::fopen(path, "r");
  if (!f) return false;
  int v;
  while (std::fscanf(f, "%i", &v) == 1) out.push_back(v);
  std::fclose(f);
  return true;
}

extern "C" int sweeptt_star_load(const char* path, float delta, struct FS** fs, int* starsize) {
  if (!path || !fs || !starsize) return sweeptt::set_error("sweeptt_star_load: null argument");
  std::vector<long> t;
  if (!read_ints(path, t)) return sweeptt::set_error("Cannot open forward star offset file: %s", path);
  if (t.empty() || t[0] <= 0 || (long)t.size() < 1 + 3 * t[0])
    return sweeptt::set_error("forward star file %s: expected a count followed by that many 'oi oj ok' rows", path);
  const int n = (int)t[0];
  FS* out = static_cast<FS*>(std::malloc(sizeof(FS) * n));
  if (!out) return sweeptt::set_error("out of memory");
  for (int l = 0; l < n; ++l) {
    out[l].i = (int)t[1 + 3 * l]; out[l].j = (int)t[2 + 3 * l]; out[l].k = (int)t[3 + 3 * l];
  }
  sweeptt_star_fill_distances(out, n, delta);
  *fs = out;
  *starsize = n;
  return 1;
}

extern "C" int sweeptt_starts_load(const char* path, struct START** starts, int* numstart) {
  if (!path || !starts || !numstart) return sweeptt::set_error("sweeptt_starts_load: null argument");
  std::vector<long> t;
  if (!read_ints(path, t)) return sweeptt::set_error("Cannot open starting points file: %s", path);
  if (t.empty() || t[0] <= 0 || (long)t.size() < 1 + 3 * t[0])
    return sweeptt::set_error("start file %s: expected a count followed by that many 'si sj sk' rows", path);
  const int n = (int)t[0];
  START* out = static_cast<START*>(std::malloc(sizeof(START) * n));
  if (!out) return sweeptt::set_error("out of memory");
  for (int s = 0; s < n; ++s) {
    out[s].i = (int)t[1 + 3 * s]; out[s].j = (int)t[2 + 3 * s]; out[s].k = (int)t[3 + 3 * s];
  }
  *starts = out;
  *numstart = n;
  return 1;
}

// ---------------------------------------------------------------------------------------
// output.tt
// ---------------------------------------------------------------------------------------
namespace {

// "00".."99"
struct Digits2 {
  char t[200];
  Digits2() { for (int i = 0; i < 100; ++i) { t[2 * i] = (char)('0' + i / 10); t[2 * i + 1] = (char)('0' + i % 10); } }
};
const Digits2 kDigits2;

inline char* put_uint(char* p, unsigned v) {
  char tmp[12];
  int n = 0;
  while (v >= 100) { const unsigned r = v % 100; v /= 100; tmp[n++] = kDigits2.t[2 * r + 1]; tmp[n++] = kDigits2.t[2 * r]; }
  if (v >= 10) { tmp[n++] = kDigits2.t[2 * v + 1]; tmp[n++] = kDigits2.t[2 * v]; } else tmp[n++] = (char)('0' + v);
  while (n) *p++ = tmp[--n];
  return p;
}
inline char* put_u64(char* p, uint64_t v) {
  char tmp[24];
  int n = 0;
  while (v >= 100) { const unsigned r = (unsigned)(v % 100); v /= 100; tmp[n++] = kDigits2.t[2 * r + 1]; tmp[n++] = kDigits2.t[2 * r]; }
  if (v >= 10) { tmp[n++] = kDigits2.t[2 * v + 1]; tmp[n++] = kDigits2.t[2 * v]; } else tmp[n++] = (char)('0' + v);
  while (n) *p++ = tmp[--n];
  return p;
}

// printf("%f") of a float, byte-identical to glibc for finite values below 2^39: the binary value m*2^e is scaled by
// 10^6 in exact integer arithmetic (m * 10^6 < 2^44, shifted left by at most 15 or right with the remainder kept) and
// rounded half-to-even, which is what glibc does in the default rounding mode.  Everything else goes through snprintf.
inline char* put_float_f(char* p, float x) {
  uint32_t bits;
  std::memcpy(&bits, &x, 4);
  const uint32_t expo = (bits >> 23) & 0xff;
  if (expo == 0xff || expo >= 127 + 39) return p + std::snprintf(p, 64, "%f", (double)x);
  if (bits >> 31) *p++ = '-';
  uint64_t m = bits & 0x7fffffu;
  int e;
  if (expo == 0) e = -149; else { m |= 0x800000u; e = (int)expo - 150; }
  const uint64_t scaled = m * 1000000u;  // < 2^44
  uint64_t q;
  if (e >= 0) {
    q = scaled << e;  // e <= 15 by the guard above: < 2^59
  } else if (-e >= 46) {
    q = 0;            // scaled < 2^44 <= half a unit of the last decimal
  } else {
    const int sh = -e;
    const uint64_t one = (uint64_t)1 << sh;
    const uint64_t rem = scaled & (one - 1);
    q = scaled >> sh;
    const uint64_t half = one >> 1;
    if (rem > half || (rem == half && (q & 1))) ++q;
  }
  const uint64_t ip = q / 1000000u;
  const unsigned fp = (unsigned)(q - ip * 1000000u);
  p = ip < 0x100000000ull ? put_uint(p, (unsigned)ip) : put_u64(p, ip);
  *p++ = '.';
  const unsigned a = fp / 10000u, bc = fp - a * 10000u, b = bc / 100u, c = bc - b * 100u;
  std::memcpy(p, &kDigits2.t[2 * a], 2);
  std::memcpy(p + 2, &kDigits2.t[2 * b], 2);
  std::memcpy(p + 4, &kDigits2.t[2 * c], 2);
  return p + 6;
}

// worst-case bytes of one line: prefix 17 + three ints (<= 10 digits) + 2 commas + "): " + %f (<= 48) + " 0 0 0\n"
constexpr size_t kLineMax = 17 + 30 + 2 + 3 + 48 + 7;

// formats rows [i0,i1) of one source's box into buf (capacity: rows * ny * nz * kLineMax); returns the bytes written
size_t format_slab(const float* tt, int i0, int i1, int ny, int nz, char* buf) {
  char* p = buf;
  const float* v = tt + (size_t)i0 * ny * nz;
  char prefix[64];
  for (int i = i0; i < i1; ++i)
    for (int j = 0; j < ny; ++j) {
      char* q = prefix;
      std::memcpy(q, "travel time for (", 17); q += 17;
      q = put_uint(q, (unsigned)i); *q++ = ',';
      q = put_uint(q, (unsigned)j); *q++ = ',';
      const size_t plen = (size_t)(q - prefix);
      for (int k = 0; k < nz; ++k) {
        std::memcpy(p, prefix, plen); p += plen;
        p = put_uint(p, (unsigned)k);
        std::memcpy(p, "): ", 3); p += 3;
        p = put_float_f(p, *v++);
        std::memcpy(p, " 0 0 0\n", 7); p += 7;
      }
    }
  return (size_t)(p - buf);
}

}  // namespace

extern "C" int sweeptt_write_output_tt(const char* path, const float* const* tt, int numstart, int nx, int ny, int nz) {
  if (!path || !tt) return sweeptt::set_error("sweeptt_write_output_tt: null argument");
  FILE* f = std::fopen(path, "w");
  if (!f) return sweeptt::set_error("Can not open travel time output file: %s", path);
  std::fprintf(f, "%d %d %d\n", nx, ny, nz);
  unsigned nthreads = std::thread::hardware_concurrency();
  if (nthreads == 0) nthreads = 1;
  nthreads = std::min<unsigned>(nthreads, 32);
  if (const char* env = std::getenv("SWEEPTT_IO_THREADS")) nthreads = std::max(1, std::atoi(env));
  nthreads = std::min<unsigned>(nthreads, (unsigned)nx);
  // The text of one source is formatted in x-slabs by worker threads while the text of the source before it is being
  // written: two sets of slab buffers (sized for the worst case; only the pages actually written become resident).
  struct Slab { std::unique_ptr<char[]> buf; size_t cap = 0, len = 0; };
  std::vector<Slab> sets[2];
  for (auto& set : sets) set.resize(nthreads);
  auto slab_rows = [&](unsigned t, int* i0, int* i1) {
    *i0 = (int)((long long)nx * t / nthreads); *i1 = (int)((long long)nx * (t + 1) / nthreads);
  };
  for (unsigned t = 0; t < nthreads; ++t) {
    int i0, i1;
    slab_rows(t, &i0, &i1);
    const size_t cap = (size_t)(i1 - i0) * ny * nz * kLineMax + 64;
    for (auto& set : sets) { set[t].buf.reset(new (std::nothrow) char[cap]); set[t].cap = cap; }
    if (!sets[0][t].buf || !sets[1][t].buf) { std::fclose(f); return sweeptt::set_error("out of memory for the output.tt buffers"); }
  }
  auto launch = [&](int s, std::vector<std::thread>& th) {
    th.clear();
    for (unsigned t = 0; t < nthreads; ++t) {
      int i0, i1;
      slab_rows(t, &i0, &i1);
      Slab* sl = &sets[s & 1][t];
      const float* src = tt[s];
      th.emplace_back([=] { sl->len = format_slab(src, i0, i1, ny, nz, sl->buf.get()); });
    }
  };
  bool ok = true;
  std::vector<std::thread> cur, next;
  if (numstart > 0) launch(0, cur);
  for (int s = 0; s < numstart; ++s) {
    for (auto& t : cur) t.join();
    if (s + 1 < numstart) launch(s + 1, next);  // formats into the other set while this one is written
    if (ok) {
      std::fprintf(f, "starting point: %d\n", s);
      for (unsigned t = 0; t < nthreads && ok; ++t)
        ok = std::fwrite(sets[s & 1][t].buf.get(), 1, sets[s & 1][t].len, f) == sets[s & 1][t].len;
    }
    cur.swap(next);
  }
  ok = (std::fclose(f) == 0) && ok;
  return ok ? 1 : sweeptt::set_error("error writing %s", path);
}
