// Measurement aid (not on any product path): a pure COPY with the memory access pattern of the four-step column passes --
// every CTA moves a tile of 128 rows x 32 complex64 columns (256-byte row segments), 16 loads in flight per thread, the
// rows either `n2` array rows apart (the strided side of a pass) or dense (the scratch side of the fused variant).
// Its bandwidth is the ceiling of that access pattern at the occupancy of the transform kernels: if the copy is no faster
// than the transform pass, the pass is bound by the pattern (DRAM pages / TLB reach), not by its arithmetic.
// tools/probe_strided_copy.py drives it.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rocketfft_b200.h"

namespace {

// mode 0: strided -> strided (same coordinates), 1: strided -> dense, 2: dense -> strided
__global__ void __launch_bounds__(256, 4) tile_copy_kernel(const char *__restrict__ in, char *__restrict__ out, int64_t pitch,
                                                         uint32_t n2, uint32_t cols, uint32_t tiles0, int mode) {
    const uint32_t tile = blockIdx.x, t0 = tile % tiles0, j0 = tile / tiles0;
    const uint32_t w = threadIdx.x & 31u, r = threadIdx.x >> 5;
    const uint32_t col = t0 * 32u + w;
    if (col >= cols) return;
    float2 v[16];
    const int64_t dense_base = ((int64_t)tile * 128) * 256 + (int64_t)w * 8;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t rr = r + 8u * (uint32_t)i;
        const char *p = (mode == 2) ? in + dense_base + (int64_t)rr * 256
                                    : in + ((int64_t)rr * n2 + j0) * pitch + (int64_t)col * 8;
        v[i] = *reinterpret_cast<const float2 *>(p);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t rr = r + 8u * (uint32_t)i;
        char *p = (mode == 1) ? out + dense_base + (int64_t)rr * 256 : out + ((int64_t)rr * n2 + j0) * pitch + (int64_t)col * 8;
        *reinterpret_cast<float2 *>(p) = v[i];
    }
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int rfb200_debug_tile_copy(const void *in, void *out, uint64_t rows,
                                                                             uint64_t cols, int64_t pitch_bytes, int mode,
                                                                             void *stream) {
    if (rows % 128 || rows / 128 == 0 || cols == 0 || mode < 0 || mode > 2) return 1;
    const uint32_t n2 = (uint32_t)(rows / 128), tiles0 = (uint32_t)((cols + 31) / 32);
    tile_copy_kernel<<<tiles0 * n2, 256, 0, (cudaStream_t)stream>>>((const char *)in, (char *)out, pitch_bytes, n2, (uint32_t)cols,
                                                                   tiles0, mode);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}
