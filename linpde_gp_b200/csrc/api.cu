// Library information entry points of liblpgp.
#include "common.cuh"

extern "C" int lpgp_version(void) { return 100; }

extern "C" const char* lpgp_build_arch(void) { return "sm_100a"; }

extern "C" const char* lpgp_error_string(int code) {
  if (code == 0) return "success";
  if (code > 0) return "leading minor is not positive definite (LAPACK info > 0)";
  if (code <= -1000) return cudaGetErrorString((cudaError_t)(-code - 1000));
  return "invalid argument";
}
