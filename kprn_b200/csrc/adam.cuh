// adam.cuh -- the one Adam element update every optimizer kernel uses (dense optim.adam of SP/model/OneModel.lua:340-355 and
// its lazy, row-sparse evaluation in train.cu).  Sharing the expression keeps the two bit-identical: with g = 0 the update
// degenerates to m *= b1, v *= b2, x -= step * m / (sqrt(v) + eps), which the lazy path replays step by step.
#pragma once
namespace kprn {
__device__ __forceinline__ void adam_elem(float& x, float& m, float& v, float g, float step, float b1, float b2, float eps) {
    const float mi = fmaf(b1, m, (1.f - b1) * g);
    const float vi = fmaf(b2, v, (1.f - b2) * g * g);
    m = mi; v = vi;
    x -= step * mi / (sqrtf(vi) + eps);       // eps outside the square root, as optim.adam
}
}  // namespace kprn
