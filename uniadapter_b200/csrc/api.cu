// Host-side plumbing of libua_b200.so: error string, launch counter, tuning knobs.
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include "common.cuh"

namespace ua {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

extern int g_fps_threads;
extern int g_dota_ksplit;
extern int g_dota_staged;
extern int g_dota_ka;
extern int g_dota_pdl;
extern int g_fps_cluster;
extern int g_knn_warps;
extern int g_knn_hist;
extern int g_modedota_threads;
extern int g_modedota_v;
extern int g_gemm_bn;
extern int g_resid_cb;
extern int g_resid_dbl;
extern int g_modedota_groups;
extern int g_modedota_logprod;
extern int g_modedota_batch;
extern int g_p2p_timeout_ms;
extern int g_sample_v;
extern int g_sample_g;
extern int g_sample_per;
extern int g_sample_skip;
extern int g_sample_trace_on;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int check_launch(const char* what) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return UA_ERR_CUDA;
  }
  return UA_OK;
}

}  // namespace ua

extern "C" int ua_abi_version(void) { return UA_ABI_VERSION; }
extern "C" const char* ua_last_error(void) { return ua::g_err; }
extern "C" int64_t ua_launch_count(void) { return ua::g_launches.load(); }
extern "C" void ua_reset_launch_count(void) { ua::g_launches.store(0); }

extern "C" int ua_set_tuning(const char* key, int value) {
  if (!key) return UA_ERR_INVALID_ARG;
  if (!strcmp(key, "fps_threads")) { ua::g_fps_threads = value; return UA_OK; }
  if (!strcmp(key, "dota_staged")) { ua::g_dota_staged = value; return UA_OK; }
  if (!strcmp(key, "dota_ka")) { ua::g_dota_ka = value; return UA_OK; }
  if (!strcmp(key, "dota_pdl")) { ua::g_dota_pdl = value; return UA_OK; }
  if (!strcmp(key, "dota_ksplit")) { ua::g_dota_ksplit = value; return UA_OK; }
  if (!strcmp(key, "fps_cluster")) { ua::g_fps_cluster = value; return UA_OK; }
  if (!strcmp(key, "knn_warps")) { ua::g_knn_warps = value; return UA_OK; }
  if (!strcmp(key, "knn_hist")) { ua::g_knn_hist = value; return UA_OK; }
  if (!strcmp(key, "modedota_threads")) { ua::g_modedota_threads = value; return UA_OK; }
  if (!strcmp(key, "modedota_groups")) { ua::g_modedota_groups = value; return UA_OK; }
  if (!strcmp(key, "resid_cb")) { ua::g_resid_cb = value; return UA_OK; }
  if (!strcmp(key, "resid_dbl")) { ua::g_resid_dbl = value; return UA_OK; }
  if (!strcmp(key, "gemm_bn")) { ua::g_gemm_bn = value; return UA_OK; }
  if (!strcmp(key, "modedota_v")) { ua::g_modedota_v = value; return UA_OK; }
  if (!strcmp(key, "modedota_logprod")) { ua::g_modedota_logprod = value; return UA_OK; }
  if (!strcmp(key, "modedota_batch")) { ua::g_modedota_batch = value; return UA_OK; }
  if (!strcmp(key, "sample_v")) { ua::g_sample_v = value; return UA_OK; }
  if (!strcmp(key, "sample_trace")) { ua::g_sample_trace_on = value; return UA_OK; }
  if (!strcmp(key, "sample_skip")) { ua::g_sample_skip = value; return UA_OK; }
  if (!strcmp(key, "sample_per")) { ua::g_sample_per = value; return UA_OK; }
  if (!strcmp(key, "sample_g")) { ua::g_sample_g = value; return UA_OK; }
  if (!strcmp(key, "p2p_timeout_ms")) { ua::g_p2p_timeout_ms = value; return UA_OK; }
  ua::set_error("ua_set_tuning: unknown key '%s'", key);
  return UA_ERR_INVALID_ARG;
}
