// tc_path.cu -- tensor-core (tcgen05) forward path.  Placeholder until the kernels land.
#include "common.cuh"
namespace kprn {
bool tc_supported(const kprn_ctx*) { return false; }
int tc_forward(kprn_ctx*, const int32_t*, const int32_t*, int64_t, int64_t, bool) {
    set_error("tensor-core path not built");
    return KPRN_EINVAL;
}
void tc_destroy(kprn_ctx*) {}
}  // namespace kprn
