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

// ---- round 2 probes: bandwidth of an L2-resident working set, and of distributed shared memory inside a cluster -----------
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
namespace {

// mode 0: read only (16-byte loads, summed), 1: copy in -> out.  Every CTA sweeps the whole buffer `reps` times, starting
// at a CTA-dependent offset so that the CTAs do not walk in lockstep.
__global__ void __launch_bounds__(512, 2) l2_sweep_kernel(const float4 *__restrict__ in, float4 *__restrict__ out, uint64_t n16, int reps,
                                                        int mode, float *sink) {
    float acc = 0.f;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; ++r) {
        uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n16; i += 4 * stride) {
            float4 a = __ldcg(in + i), b = __ldcg(in + i + stride), c = __ldcg(in + i + 2 * stride), d = __ldcg(in + i + 3 * stride);
            if (mode) { __stcg(out + i, a); __stcg(out + i + stride, b); __stcg(out + i + 2 * stride, c); __stcg(out + i + 3 * stride, d); }
            else acc += a.x + b.y + c.z + d.w;
        }
        for (; i < n16; i += stride) {
            float4 a = __ldcg(in + i);
            if (mode) __stcg(out + i, a);
            else acc += a.x;
        }
    }
    if (acc == 12345.678f) *sink = acc;
}

// Every CTA of a cluster writes (mode 0) / reads (mode 1) `kb` KiB to / from EACH other CTA's shared memory, `reps` times.
// cycles[cta] = SM cycles the exchange took.
__global__ void __launch_bounds__(256, 1) dsmem_kernel(uint32_t kb, int reps, int mode, unsigned long long *cycles, float *sink) {
    extern __shared__ __align__(16) unsigned char smem_probe[];
    cg::cluster_group cl = cg::this_cluster();
    const unsigned nr = cl.num_blocks(), me = cl.block_rank();
    float4 *mine = reinterpret_cast<float4 *>(smem_probe);
    const uint32_t n16 = kb * 64u;  // 16-byte items per peer region
    for (uint32_t i = threadIdx.x; i < n16 * nr; i += blockDim.x) mine[i] = make_float4((float)i, 1.f, 2.f, 3.f);
    cl.sync();
    float acc = 0.f;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        for (unsigned d = 1; d < nr; ++d) {
            const unsigned peer = (me + d) % nr;
            float4 *remote = cl.map_shared_rank(mine, peer) + (size_t)me * n16;  // my slot in the peer's buffer
            if (mode == 0) {
                for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) remote[i] = make_float4(acc, (float)i, 0.f, 1.f);
            } else {
                for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) { const float4 v = remote[i]; acc += v.x + v.w; }
            }
        }
    }
    cl.sync();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 12345.678f) *sink = acc;
}

}  // namespace

extern "C" __attribute__((visibility("default"))) int rfb200_debug_l2_sweep(const void *in, void *out, uint64_t bytes, int reps, int mode,
                                                                            uint32_t ctas, void *sink, void *stream) {
    l2_sweep_kernel<<<ctas, 512, 0, (cudaStream_t)stream>>>((const float4 *)in, (float4 *)out, bytes / 16, reps, mode, (float *)sink);
    return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

extern "C" __attribute__((visibility("default"))) int rfb200_debug_dsmem(uint32_t cluster, uint32_t nclusters, uint32_t kb, int reps, int mode,
                                                                         void *cycles, void *sink, void *stream) {
    const size_t smem = (size_t)kb * 1024 * cluster;
    if (cluster < 2 || cluster > 16 || smem > 227 * 1024) return 1;
    if (cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 2;
    if (cluster > 8 && cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cluster * nclusters);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, dsmem_kernel, kb, reps, mode, (unsigned long long *)cycles, (float *)sink);
    return e == cudaSuccess ? 0 : 2;
}
