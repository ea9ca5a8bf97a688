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

// General form: every CTA copies `nrows` segments of `seg` bytes (8-byte items, all loads issued before the first store);
// segment r of tile (a, b) lies at  a*seg + b*outer + r*row_stride  (a < tiles0).
template <int PER>
__global__ void __launch_bounds__(1024) seg_copy_kernel(const char *__restrict__ in, char *__restrict__ out, uint32_t nrows, uint32_t seg8,
                                                      int64_t row_stride, uint32_t tiles0, int64_t outer) {
    const uint32_t a = blockIdx.x % tiles0, b = blockIdx.x / tiles0;
    const int64_t base = (int64_t)a * seg8 * 8 + (int64_t)b * outer;
    const uint32_t total = nrows * seg8;
    float2 v[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const uint32_t idx = threadIdx.x + (uint32_t)i * blockDim.x;
        if (idx < total) v[i] = *reinterpret_cast<const float2 *>(in + base + (int64_t)(idx / seg8) * row_stride + (int64_t)(idx % seg8) * 8);
    }
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const uint32_t idx = threadIdx.x + (uint32_t)i * blockDim.x;
        if (idx < total) *reinterpret_cast<float2 *>(out + base + (int64_t)(idx / seg8) * row_stride + (int64_t)(idx % seg8) * 8) = v[i];
    }
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int rfb200_debug_seg_copy(const void *in, void *out, uint32_t nrows, uint32_t seg_bytes,
                                                                            int64_t row_stride, uint32_t tiles0, uint32_t tiles1,
                                                                            int64_t outer_stride, uint32_t threads, uint32_t smem_bytes,
                                                                            void *stream) {
    // smem_bytes: dynamic shared memory requested per CTA (unused by the kernel) to reproduce a transform kernel's occupancy
    if (seg_bytes % 8 || threads == 0 || threads > 1024 || tiles0 == 0 || tiles1 == 0 || smem_bytes > 227 * 1024) return 1;
    if (cudaFuncSetAttribute(seg_copy_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) return 2;
    const uint32_t seg8 = seg_bytes / 8;
    const uint64_t per = ((uint64_t)nrows * seg8 + threads - 1) / threads;
    if (per > 16) return 1;
    seg_copy_kernel<16><<<tiles0 * tiles1, threads, smem_bytes, (cudaStream_t)stream>>>((const char *)in, (char *)out, nrows, seg8, row_stride, tiles0,
                                                                            outer_stride);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

extern "C" __attribute__((visibility("default"))) int rfb200_debug_tile_copy(const void *in, void *out, uint64_t rows,
                                                                             uint64_t cols, int64_t pitch_bytes, int mode,
                                                                             void *stream) {
    if (rows % 128 || rows / 128 == 0 || cols == 0 || mode < 0 || mode > 2) return 1;
    const uint32_t n2 = (uint32_t)(rows / 128), tiles0 = (uint32_t)((cols + 31) / 32);
    tile_copy_kernel<<<tiles0 * n2, 256, 0, (cudaStream_t)stream>>>((const char *)in, (char *)out, pitch_bytes, n2, (uint32_t)cols,
                                                                   tiles0, mode);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}
