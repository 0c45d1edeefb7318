#include "../../include/deephall_b200.h"
#include "kernels.h"
struct dh_plan;
size_t vjp_ws_floats(const dh_plan*, int64_t) { return 0; }
extern "C" int dh_logpsi_vjp(dh_plan*, const float*, const float*, int64_t, const float*, float*, float*, void*, size_t, void*) { return DH_E_UNSUPPORTED; }
