// g16_version(): library version + digest of the sources it was built from.  build.py recompiles this unit on
// every build with -DG16_SOURCE_HASH=<sha256 of csrc/*.cu, csrc/*.cuh, include/*.h>, so bench.py and the tests can
// tell a stale prebuilt lib/libg16cuda.so from a fresh one.
#include "../../include/g16_cuda.h"

#ifndef G16_SOURCE_HASH
#define G16_SOURCE_HASH "unhashed"
#endif
#ifdef G16_EMU
#define G16_TARGET "host-emulation (tests only)"
#else
#define G16_TARGET "sm_100a"
#endif

extern "C" const char *g16_version(void) { return "groth16-cuda 0.2 (" G16_TARGET ") src:" G16_SOURCE_HASH; }
