// Library-wide pieces of the C ABI: version, error string, launch counter.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace dmvs {

std::atomic<unsigned long long> g_launches{0};
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern int g_tc2_max_ctas;
extern int g_head_px;
extern int g_pb_td8;
extern int g_tc2_pdl;
extern int g_regnet_streams;
extern int g_kf;
extern int g_kf_dbg;
extern int g_kf_pdl;
extern long long* g_kf_trace;
extern int g_kf_wide;
extern int g_tc2_skip_prefetch;
extern int g_kf_mw;

}  // namespace dmvs

// tuning knobs for experiments on the GPU box (not part of the reference-facing surface; values persist per process)
extern "C" int dmvs_debug_set(const char* key, int value) {
  if (key && !strcmp(key, "tc2_max_ctas") && value >= 1) {
    dmvs::g_tc2_max_ctas = value;
    return DMVS_OK;
  }
  if (key && !strcmp(key, "pb_td8") && (value == 0 || value == 1)) {
    dmvs::g_pb_td8 = value;
    return DMVS_OK;
  }
  if (key && !strcmp(key, "tc2_pdl") && (value == 0 || value == 1)) {
    dmvs::g_tc2_pdl = value;
    return DMVS_OK;
  }
  if (key && !strcmp(key, "regnet_streams") && (value == 0 || value == 1)) {
    dmvs::g_regnet_streams = value;
    return DMVS_OK;
  }
  if (key && !strcmp(key, "kf") && value >= 0 && value <= 2) {
    dmvs::g_kf = value;
    return DMVS_OK;
  }
  if (key && !strcmp(key, "tc2_skip_prefetch") && (value == 0 || value == 1)) {
    dmvs::g_tc2_skip_prefetch = value;
    return DMVS_OK;
  }
  if (key && !strcmp(key, "kf_dbg") && value >= 0 && value <= 15) {
    dmvs::g_kf_dbg = value;
    return DMVS_OK;
  }
  if (key && !strcmp(key, "kf_pdl") && (value == 0 || value == 1)) {
    dmvs::g_kf_pdl = value;
    return DMVS_OK;
  }
  if (key && !strcmp(key, "kf_wide") && (value == 0 || value == 1)) {
    dmvs::g_kf_wide = value;
    return DMVS_OK;
  }
  if (key && !strcmp(key, "kf_mw") && (value == 0 || value == 1 || value == 2)) {
    dmvs::g_kf_mw = value;
    return DMVS_OK;
  }
  if (key && !strcmp(key, "head_px") && (value == 32 || value == 64 || value == 128)) {
    dmvs::g_head_px = value;
    return DMVS_OK;
  }
  dmvs::set_error("dmvs_debug_set: unknown key or bad value");
  return DMVS_ERR_BAD_SHAPE;
}

// experiments: device buffers handed to instrumented kernels ("kf_trace": >= 512 x 16 int64 clock stamps of CTA 0, conv_kf.cu)
extern "C" int dmvs_debug_set_ptr(const char* key, void* ptr) {
  if (key && !strcmp(key, "kf_trace")) {
    dmvs::g_kf_trace = static_cast<long long*>(ptr);
    return DMVS_OK;
  }
  dmvs::set_error("dmvs_debug_set_ptr: unknown key");
  return DMVS_ERR_BAD_SHAPE;
}

extern "C" int dmvs_abi_version(void) { return DMVS_ABI_VERSION; }
extern "C" const char* dmvs_last_error(void) { return dmvs::g_err; }
extern "C" unsigned long long dmvs_launch_count(void) { return dmvs::g_launches.load(); }
