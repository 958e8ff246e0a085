/* oracle/_ref/libsrand_shim.so -- TEST INFRASTRUCTURE ONLY.
 * LD_PRELOADed into the reference's own protein2datapoints binary so that its
 * srand(time(NULL)) calls (protein2datapoints.cpp:38,85; protein.hpp:45) seed the
 * value of HS_SEED instead: the random window stride (30 + rand() % 20, :57,70) then
 * repeats, and the multi-window / duplicate-k-mer branch can be compared byte for
 * byte with the drop-in (which honours HS_SEED itself).  No reference file is edited. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdlib.h>

void srand(unsigned int seed) {
  static void (*real)(unsigned int) = 0;
  if (!real) real = (void (*)(unsigned int))dlsym(RTLD_NEXT, "srand");
  const char *e = getenv("HS_SEED");
  real(e ? (unsigned int)strtoul(e, 0, 10) : seed);
}
