// Host side of the fused four-step kernel (fused4v2_kernel.cuh): eligibility, ring / counter scratch, launch.
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "fused4v2_kernel.cuh"
#include "geom_fill.cuh"

namespace rfb {

static int env_i(const char *name, int dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

// One pinned, device-visible word per process: a kernel that gave up waiting for a dependency sets it; the next library call
// (and the synchronous host-array path) reports it instead of returning a wrong result silently.
uint32_t *async_error_word() {
    static std::mutex mu;
    static uint32_t *w = nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (!w) {
        void *p = nullptr;
        if (cudaHostAlloc(&p, 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        memset(p, 0, 64);
        w = (uint32_t *)p;
    }
    return w;
}

void check_async_error() {
    uint32_t *w = async_error_word();
    if (w && *(volatile uint32_t *)w) {
        *(volatile uint32_t *)w = 0;
        set_error("fused four-step kernel: a tile waited for a dependency that never completed (result invalid)");
        throw Error();
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            f = nullptr;
        }
        return (EncodeTiledFn)f;
    }();
    return fn;
}

// 4-D map over 8-byte items: [x, y, z, w] with byte strides (8, sy, sz, sw) and a box of box_x x 1 x 128 x 1 items
static bool make_map(CUtensorMap *m, const void *base, uint64_t nx, uint64_t ny, uint64_t nz, uint64_t nw, uint64_t sy, uint64_t sz,
                     uint64_t sw, CUtensorMapL2promotion prom, uint32_t box_x = 32) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    const cuuint64_t dim[4] = {nx, ny, nz, nw};
    const cuuint64_t str[3] = {sy, sz, sw};
    const cuuint32_t box[4] = {box_x, 1, 128, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void *>(base), dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int STAGES, bool B_TMA_STORE, bool A_REG, bool DBG_COPY = false>
static bool launch_f4v2(const F4v2Params &p, const F4v2Maps &maps, cudaStream_t s) {
    auto kern = fft_fourstep_fused2_kernel<STAGES, B_TMA_STORE, A_REG, DBG_COPY>;
    const size_t smem = (size_t)STAGES * f4v2::STAGE_BYTES + (size_t)STAGES * 32;
    static thread_local int dev_set = -1;
    static thread_local int per_sm = 0, sms = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev_set != dev) {
        RFB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RFB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, f4v2::NTHREADS, smem));
        RFB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        dev_set = dev;
    }
    if (per_sm < 1) return false;
    const unsigned grid = (unsigned)std::min<uint64_t>(p.total_items, (uint64_t)sms * (uint64_t)per_sm);
    kern<<<grid, f4v2::NTHREADS, smem, s>>>(p, maps);
    count_launch(STAGES == 2 ? "fft_fourstep_fused2_kernel<2>" : "fft_fourstep_fused2_kernel<3>");
    RFB_CUDA_CHECK(cudaGetLastError());
    return true;
}

// job: n = 16384 complex64 points, plain c2c; dims: [neighbouring lines (8 bytes apart on both sides)] or [lines, outer].
// false (nothing launched) if the job is not one this kernel takes.
bool launch_fourstep_fused2_f32(const LineJob &job, const std::vector<Dim> &dims, cudaStream_t s) {
    const int64_t esz = 8;
    if (job.prec != 0 || job.n != 16384) return false;
    if (job.load_mode != LD_C2C || job.store_mode != ST_C2C || job.flags || (job.n_in && job.n_in != job.n) || job.twN ||
        job.pre_tab || job.post_tab || !job.split_out.empty() || job.conv)
        return false;
    if (dims.empty() || dims.size() > 2) return false;
    if (dims[0].is != esz || dims[0].os != esz || job.is < 16 * esz || job.os < 16 * esz || (job.is % esz) || (job.os % esz)) return false;
    if (((uintptr_t)job.in % esz) || ((uintptr_t)job.out % esz)) return false;
    int64_t outer = 1, in_outer = 16384 * job.is, out_outer = 16384 * job.os;
    if (dims.size() == 2) {
        outer = dims[1].n;
        in_outer = dims[1].is;
        out_outer = dims[1].os;
        // the outer stride is a tensor-map stride (multiple of 16 bytes) and the arrays' outer items do not interleave
        if (in_outer < 16384 * job.is || out_outer < 16384 * job.os || (in_outer % 16) || (out_outer % 16)) return false;
    }
    const int64_t cols = dims[0].n;
    // smaller arrays: the two-launch path keeps its intermediate in L2 by itself
    static const int min_mb = env_i("RFB200_FUSE4_MIN_MB", 64);
    if ((uint64_t)cols * (uint64_t)outer * job.n * (uint64_t)esz < ((uint64_t)min_mb << 20) || cols < 32) return false;
    const int64_t spo = (cols + 31) / 32;
    // Strips per group (see the kernel's header): 1.  Measured on B200 (profiles/r02h_fuse4v2_strip_groups.log): grouping 2 / 4 / 8
    // adjacent strips in the tile order -- more contiguous bytes per array row in flight -- is slower (0.67 / 0.73 / 0.79 ms
    // against 0.67 ms): the copy engine's 256-byte row requests spread best over the L2 slices when they are scattered.
    const int glog = 0;
    const int64_t G = 1ll << glog;
    const int64_t gpo = (spo + G - 1) / G;
    const uint64_t ngroups = (uint64_t)outer * (uint64_t)gpo, S = ngroups * (uint64_t)G;
    if (S >= (1u << 22) || (uint64_t)(16384 * job.is) >= (1ull << 40) || (uint64_t)in_outer >= (1ull << 40) ||
        (uint64_t)out_outer >= (1ull << 40))
        return false;
    // 6 strips between the two steps of a strip, a ring of 10 slots (40 MiB): measured against 3 / 6 and 8 / 14
    // (profiles/r02c_fuse4v2_tma_sweep.log); the tiles of about 7 strips are in flight at any time.
    static const int lag_d = env_i("RFB200_FUSE4_LAG", 6), ring_d = env_i("RFB200_FUSE4_RING", 10);  // (sweep: profiles/r02z_fuse4v2_ring_lag.log)
    F4v2Params p;
    memset(&p, 0, sizeof(p));
    p.nstrips = (uint32_t)S;
    p.spo = (uint32_t)spo;
    p.ngroups = (uint32_t)ngroups;
    p.gpo = (uint32_t)gpo;
    p.glog = (uint32_t)glog;
    p.lag = (uint32_t)std::min<uint64_t>((uint64_t)lag_d, ngroups);
    p.ring = (uint32_t)std::max<int64_t>(ring_d, (int64_t)p.lag + 1);
    p.total_items = (uint32_t)(2 * S * 128);
    p.d_gpo = make_fastdiv((uint32_t)gpo);
    p.d_ring = make_fastdiv(p.ring << glog);
    p.backward = job.backward ? 1 : 0;
    p.fct = (float)job.fct;
    p.stw = (const float2 *)get_table(TAB_STOCKHAM, 0, 128, 0);
    const uint32_t tS = split_size(job.n);
    p.d_twS = make_fastdiv(tS);
    p.twA = (const float2 *)get_table(TAB_SPLIT_A, 0, job.n, tS);
    p.twB = (const float2 *)get_table(TAB_SPLIT_B, 0, job.n, tS);
    p.max_idle = 1u << 22;
    static const int dbg_copy = env_i("RFB200_FUSE4_DEBUG_COPY", 0);  // measurement aid (wrong results): no transforms
    static const int dbg_only = env_i("RFB200_FUSE4_DEBUG_ONLY", 0);  // measurement aid (wrong results): 1 = A tiles only, 2 = B tiles only
    p.dbg_only = (uint32_t)dbg_only;
    if (dbg_only) p.total_items = (uint32_t)(S * 128);
    p.host_err = async_error_word();
    const size_t slot_bytes = (size_t)128 * f4v2::TILE_BYTES;
    const size_t ring_slots = (size_t)p.ring << glog;
    const size_t ring_bytes = ring_slots * slot_bytes, ctr_bytes = (2 * S + 2) * sizeof(uint32_t);
    Scratch sc(ring_bytes + ctr_bytes, s);
    p.ring_mem = (char *)sc.p;
    p.ctr = (uint32_t *)(p.ring_mem + ring_bytes);
    // Tensor maps.  Array side: rows of one parity (a tile's rows j1*128 + j0 all have the parity of j0) are 2 * pitch apart,
    // which is a multiple of 16 bytes even when the pitch is not.  A parity class whose rows start 8 bytes past a 16-byte
    // boundary gets its map base moved 8 bytes down (x = 0 is the element before the row) and, on the input side, a box of
    // 34 elements: the box then starts on a 16-byte boundary and the tile sits 8 bytes into its rows.  On the output side
    // such a class has no map (its rows are stored from registers).
    p.out = job.out;
    p.out_pitch = job.os;
    p.out_outer = out_outer;
    p.cols = (uint32_t)cols;
    alignas(64) F4v2Maps maps;
    memset(&maps, 0, sizeof(maps));
    for (int side = 0; side < 2; ++side) {
        const char *base = side ? job.out : job.in;
        const int64_t pitch = side ? job.os : job.is, ostr = side ? out_outer : in_outer;
        for (int c = 0; c < 2; ++c) {
            const char *b0 = base + c * pitch;
            const uint32_t mis = (uint32_t)((uintptr_t)b0 & 15u);
            (side ? p.mis_out : p.mis_in)[c] = mis ? 1u : 0u;
            if (side && mis) {
                maps.m[2 + c] = maps.m[0];  // never used
                continue;
            }
            // Extent along x.  Measured: the copy engine checks the x bound in 16-byte granules -- with an odd number of
            // 8-byte elements the last one reads as zero.  Loads: round the extent up to even (the extra element is the
            // first one of the next row, never used; refused when it would lie past the end of the array).  Stores: only
            // full strips go through the copy engine, so the extent is the full strips' (a multiple of 32).
            uint64_t nx = (uint64_t)cols + mis / 8;
            if (side) nx = (uint64_t)(cols / 32) * 32;
            else if (nx & 1) {
                if (c == 1) return false;  // this class holds the last row of the array: the extra element would be out of bounds
                ++nx;
            }
            if (nx == 0) { maps.m[2 * side + c] = maps.m[0]; continue; }
            if (!make_map(&maps.m[2 * side + c], b0 - mis, nx, 64, 128, (uint64_t)outer, (uint64_t)(2 * pitch),
                          (uint64_t)(128 * pitch), (uint64_t)ostr, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, mis ? 34u : 32u))
                return false;
        }
    }
    // ring: [32 lines][j0: 256 B][k1: 32 KiB][slot: 4 MiB]; A stores the box (all lines, one j0, all k1) of its slot
    if (!make_map(&maps.m[4], p.ring_mem, 32, 128, 128, ring_slots, 256, f4v2::TILE_BYTES, slot_bytes, CU_TENSOR_MAP_L2_PROMOTION_NONE))
        return false;
    RFB_CUDA_CHECK(cudaMemsetAsync(p.ctr, 0, ctr_bytes, s));
    // 2 stages x 3 CTAs per SM with every tile staged and stored by the copy engine.  Measured alternatives
    // (profiles/r02*_fuse4v2_*.log): 3 stages x 2 CTAs 0.70 ms, B tiles stored from registers 0.654 ms, A tiles stored from
    // registers + barrier/fence publication 0.654-0.86 ms, against 0.645 ms.
    if (dbg_copy) return launch_f4v2<2, true, false, true>(p, maps, s);
    return launch_f4v2<2, true, false, false>(p, maps, s);
}

}  // namespace rfb
