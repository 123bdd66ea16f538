// Library-level entry points of libnudf.so: error reporting, ABI version, engine selection.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/nudf.h"
#include "common.cuh"

namespace nudf {

static thread_local char g_err[512] = "";
static int g_engine = -1;
static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// NUDF_TC_MASK: which chains may run on the tensor engine (bits: 1 fwd value, 2 reverse sweep, 4 tangent, 8 backward,
// 16 weight gradients, 32 colour-net backward, 64 NeRF++ backward, 128 colour / NeRF++ forward).  Default (126): everything
// except the forward value chains (gemm_engine.cuh explains why).
static int g_tc_mask = -1;
int tc_mask() {
  if (g_tc_mask < 0) {
    const char* e = getenv("NUDF_TC_MASK");
    g_tc_mask = e ? atoi(e) : (2 | 4 | 8 | 16 | 32 | 64);
  }
  return g_tc_mask;
}

// NUDF_PLANES / nudf_set_chain_planes: plane-fed reverse-sweep and tangent chains (gemm_engine.cuh)
static int g_chain_planes = -1;
int chain_planes_flag() {
  if (g_chain_planes < 0) {
    const char* e = getenv("NUDF_PLANES");
    g_chain_planes = (e && atoi(e) != 0) ? 1 : 0;
  }
  return g_chain_planes;
}

namespace tc {
int tc_debug() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("NUDF_TC_DEBUG"); v = e ? atoi(e) : 0; }
  return v;
}
}  // namespace tc

int get_engine() {
  if (g_engine < 0) {
    const char* e = getenv("NUDF_ENGINE");
    g_engine = (e && e[0] == '0') ? 0 : 1;
  }
  return g_engine;
}

}  // namespace nudf

extern "C" {

int nudf_abi_version(void) { return NUDF_ABI_VERSION; }
const char* nudf_last_error(void) { return nudf::g_err; }
int nudf_set_engine(int engine) {
  if (engine != 0 && engine != 1) {
    nudf::set_error("nudf_set_engine: engine must be 0 (fp32 FFMA) or 1 (tcgen05 3xBF16)");
    return -1;
  }
  nudf::g_engine = engine;
  return 0;
}
int nudf_get_engine(void) { return nudf::get_engine(); }
int nudf_set_tc_mask(int mask) { nudf::g_tc_mask = mask & 255; return 0; }
int nudf_get_tc_mask(void) { return nudf::tc_mask(); }
int nudf_set_chain_planes(int on) { nudf::g_chain_planes = on ? 1 : 0; return 0; }
int nudf_get_chain_planes(void) { return nudf::chain_planes_flag(); }
int64_t nudf_launch_count(void) { return (int64_t)__atomic_load_n(&nudf::g_launches, __ATOMIC_RELAXED); }

}  // extern "C"
