#include "pow2_launch.cuh"
namespace rfb {
bool launch_pow2_f32(const LineJob &job, const std::vector<Dim> &dims, bool load_lf, bool store_lf, cudaStream_t s) {
    return launch_pow2_any<float, 14>(job, dims, load_lf, store_lf, s);
}
}  // namespace rfb
