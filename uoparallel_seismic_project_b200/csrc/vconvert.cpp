// vconvert.cpp -- text velocity file -> .vbox, the tools/vconvert.c command line
// (`vconvert <in:oldfile.txt> <out:newfile.vbox>`, tools/vconvert.c:15-35) on top of the C ABI.
// Accepts both text dialects; output is byte-identical to the reference tool's for dialect A.
#include <cstdio>

#include "../../include/sweeptt.h"

int main(int argc, char* argv[]) {
  if (argc != 3) {
    std::printf("vconvert: velocity file converter\n");
    std::printf("usage: %s <in:oldfile.txt> <out:newfile.vbox>\n", argv[0]);
    return 0;
  }
  float* v = nullptr;
  int origin[3], dims[3];
  std::printf("reading old velocity model %s...", argv[1]);
  std::fflush(stdout);
  if (!sweeptt_text_load(argv[1], &v, origin, dims)) {
    std::fprintf(stderr, "%s\n", sweeptt_last_error());
    return 1;
  }
  std::printf(" done.\n");
  std::printf("writing new velocity model %s...", argv[2]);
  std::fflush(stdout);
  if (!sweeptt_vbox_store(argv[2], v, origin, dims)) {
    std::fprintf(stderr, "%s\n", sweeptt_last_error());
    return 1;
  }
  std::printf(" done.\n");
  sweeptt_free(v);
  return 0;
}
