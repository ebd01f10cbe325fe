// Library-level entry points: version, build info, thread-local error string.
#include <stdio.h>
#include <string.h>

#include "../../include/vilbert_b200.h"
#include "common.cuh"

static thread_local char g_err[512] = "";

extern "C" void vb_set_last_error(const char* what, const char* detail) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what ? what : "?", detail ? detail : "");
}
extern "C" const char* vb_last_error(void) { return g_err; }
extern "C" int vb_abi_version(void) { return 4; }
extern "C" const char* vb_build_info(void) {
  return "libvilbert_b200 sm_100a; CUDA " VB_STR2(__CUDACC_VER_MAJOR__) "." VB_STR2(__CUDACC_VER_MINOR__) "; " __DATE__;
}
