// Host side of the persistent TMA-fed kernel for strided 512 / 1024-point complex64 lines (pow2_stream_kernel.cuh).
#include <stdlib.h>
#include <string.h>

#include "geom_fill.cuh"
#include "pow2_stream_kernel.cuh"

namespace rfb {

typedef CUresult (*EncodeTiledFnS)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnS encode_tiled_stream() {
    static EncodeTiledFnS fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            f = nullptr;
        }
        return (EncodeTiledFnS)f;
    }();
    return fn;
}

template <int LOGN, int W>
static bool launch_stream_inst(const LineJob &job, const std::vector<Dim> &dims, cudaStream_t s) {
    using SB = StreamBody<LOGN, W>;
    EncodeTiledFnS enc = encode_tiled_stream();
    if (!enc) return false;
    TileGeom<float> g;
    const uint64_t ntiles = fill_geom<float>(g, job, dims, (uint32_t)W, true, true);
    // input as a 4-D tensor of 8-byte items: [neighbouring lines][transform axis][batch dim 1][batch dim 2]
    const uint64_t span = (uint64_t)job.n * (uint64_t)job.is;  // (a stride for the dims that do not exist)
    const cuuint64_t dim[4] = {g.bext[0], (cuuint64_t)job.n, g.bext[1], g.bext[2]};
    const cuuint64_t str[3] = {(cuuint64_t)job.is, g.bext[1] > 1 ? (cuuint64_t)g.in_bs[1] : span,
                               g.bext[2] > 1 ? (cuuint64_t)g.in_bs[2] : span};
    const cuuint32_t box[4] = {(cuuint32_t)W, (cuuint32_t)SB::BOXROWS, 1, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    alignas(64) CUtensorMap map;
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<char *>(job.in), dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    StreamParams p;
    memset(&p, 0, sizeof(p));
    p.out = job.out;
    p.out_sa = job.os;
    p.out_bs1 = g.out_bs[1];
    p.out_bs2 = g.out_bs[2];
    p.bext0 = g.bext[0];
    p.ntiles = (uint32_t)ntiles;
    p.d_t0 = g.d_t0;
    p.d_e1 = g.d_e1;
    p.backward = job.backward ? 1 : 0;
    p.fct = (float)job.fct;
    const float2 *stw = (const float2 *)get_table(TAB_STOCKHAM, 0, 1ull << LOGN, 0);
    auto kern = fft_pow2_stream_kernel<LOGN, W>;
    static thread_local int dev_set = -1;
    static thread_local int sms = 0, per_sm = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev_set != dev) {
        RFB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SB::SMEM));
        RFB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SB::NT, SB::SMEM));
        RFB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        dev_set = dev;
    }
    if (per_sm < 1) return false;
    const unsigned grid = (unsigned)std::min<uint64_t>(ntiles, (uint64_t)sms * (uint64_t)per_sm);
    kern<<<grid, SB::NT, SB::SMEM, s>>>(p, map, stw);
    {
        char nm[64];
        snprintf(nm, sizeof nm, "fft_pow2_stream_kernel<%d,%d>", LOGN, W);
        count_launch(nm);
    }
    RFB_CUDA_CHECK(cudaGetLastError());
    return true;
}

// job: plain c2c of 512 / 1024 complex64 points along a strided axis whose neighbouring lines are adjacent (8 bytes apart) on
// both sides, everything 16-byte aligned (tensor-map rule on the way in, 16-byte stores on the way out), positive strides,
// enough tiles to fill the device a few times.  RFB200_STREAM=0 switches the kernel off.  false: nothing launched.
bool launch_pow2_stream_f32(const LineJob &job, const std::vector<Dim> &dims, cudaStream_t s) {
    static const int on = [] { const char *e = getenv("RFB200_STREAM"); return e ? atoi(e) : 1; }();
    if (!on || job.prec != 0 || (job.n != 512 && job.n != 1024)) return false;
    if (job.load_mode != LD_C2C || job.store_mode != ST_C2C || job.flags || (job.n_in && job.n_in != job.n) || job.twN ||
        job.pre_tab || job.post_tab || !job.split_out.empty() || job.conv)
        return false;
    if (dims.empty() || dims.size() > (size_t)MAXB || dims[0].is != 8 || dims[0].os != 8 || dims[0].n < 16) return false;
    if (((uintptr_t)job.in & 15) || ((uintptr_t)job.out & 15)) return false;
    // rows further apart than 64 KiB (every row of a box on its own page): the copy engine is slower than plain loads (see the kernel's header)
    static const int64_t max_stride = [] { const char *e = getenv("RFB200_STREAM_MAX_STRIDE"); return e ? (int64_t)atoll(e) : (int64_t)65536; }();
    if (job.is > max_stride) return false;
    if (job.is <= 0 || job.os <= 0 || (job.is & 15) || (job.os & 15) || (uint64_t)job.is >= (1ull << 40)) return false;
    for (size_t d = 1; d < dims.size(); ++d)
        if (dims[d].is <= 0 || dims[d].os <= 0 || (dims[d].is & 15) || (dims[d].os & 15) || (uint64_t)dims[d].is >= (1ull << 40)) return false;
    uint64_t ntiles = (uint64_t)(dims[0].n + 15) / 16;
    for (size_t d = 1; d < dims.size(); ++d) ntiles *= (uint64_t)dims[d].n;
    static const int min_tiles = [] { const char *e = getenv("RFB200_STREAM_MIN_TILES"); return e ? atoi(e) : 592; }();
    if (ntiles < (uint64_t)min_tiles) return false;
    return job.n == 1024 ? launch_stream_inst<10, 16>(job, dims, s) : launch_stream_inst<9, 16>(job, dims, s);
}

}  // namespace rfb
