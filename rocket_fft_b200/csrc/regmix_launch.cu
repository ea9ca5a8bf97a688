// Host side of the register-resident mixed-radix kernel (regmix_kernel.cuh).
#include <algorithm>

#include "geom_fill.cuh"
#include "regmix_kernel.cuh"

namespace rfb {

static const size_t RM_MAX_SMEM = 227 * 1024;

template <typename T, bool ALIGNED, int E, int THREADS, int MINB>
static void launch_regmix_inst(const LineJob &job, const std::vector<Dim> &dims, const RmPlan &pl, uint32_t W,
                               bool load_lf, bool store_lf, cudaStream_t s) {
    TileGeom<T> g;
    const uint64_t ntiles = fill_geom<T>(g, job, dims, W, load_lf, store_lf);
    set_prefetch_by_mode<T>(g, job, dims, W);
    g.ptw = (const cx<T> *)get_table(TAB_REGMIX, job.prec, job.n, pl.cap);
    const size_t smem = (size_t)W * pl.pitch * sizeof(cx<T>);
    auto kern = fft_regmix_kernel<T, ALIGNED, E, THREADS, MINB>;
    static thread_local int dev_set = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev_set != dev) {
        RFB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RM_MAX_SMEM));
        dev_set = dev;
    }
    kern<<<(unsigned)ntiles, W * pl.TPL, smem, s>>>(g, pl);
    count_launch(sizeof(T) == 8 ? "fft_regmix_kernel<double>" : "fft_regmix_kernel<float>");
    RFB_CUDA_CHECK(cudaGetLastError());
}

template <typename T, bool ALIGNED>
static void launch_regmix_typed(const LineJob &job, const std::vector<Dim> &dims, const RmPlan &pl, uint32_t W,
                                uint32_t E, bool load_lf, bool store_lf, cudaStream_t s) {
    constexpr int E1 = sizeof(T) == 4 ? 16 : 8, E2 = 2 * E1;
    // small register tiles and CTAs (<= 256 threads, 3 per SM) for short lines, double tiles for long ones
    const bool small_cta = W * pl.TPL <= 256;
    if (E == (uint32_t)E1) {
        if (small_cta) launch_regmix_inst<T, ALIGNED, E1, 256, 3>(job, dims, pl, W, load_lf, store_lf, s);
        else launch_regmix_inst<T, ALIGNED, E1, 512, 1>(job, dims, pl, W, load_lf, store_lf, s);
    } else {
        if (small_cta) launch_regmix_inst<T, ALIGNED, E2, 256, 2>(job, dims, pl, W, load_lf, store_lf, s);
        else launch_regmix_inst<T, ALIGNED, E2, 512, 1>(job, dims, pl, W, load_lf, store_lf, s);
    }
}

#ifdef RFB_RM_DOUBLE
bool launch_regmix_f64(const LineJob &job, const std::vector<Dim> &dims, bool load_lf, bool store_lf, bool aligned,
                       cudaStream_t s) {
#else
bool launch_regmix_f64(const LineJob &job, const std::vector<Dim> &dims, bool load_lf, bool store_lf, bool aligned,
                       cudaStream_t s);
bool launch_regmix(const LineJob &job, const std::vector<Dim> &dims, bool load_lf, bool store_lf, bool aligned,
                   cudaStream_t s) {
    if (job.prec) return launch_regmix_f64(job, dims, load_lf, store_lf, aligned, s);
#endif
    if (dims.size() > (size_t)MAXB || !aligned) return false;
    const uint64_t n = job.n;
    if (n < 6 || n > 16384) return false;
    if (job.store_mode == ST_HC || job.load_mode >= LD_DCT2 || job.store_mode >= ST_DCT2) return false;
    if (!job.split_out.empty() || job.pre_tab || job.post_tab) return false;
    // the very schedule the pass-major twiddle table (TAB_REGMIX) is built for; points per thread E:
    // the small tile if the line then needs at most 256 threads, else the double tile
    const uint32_t E1 = job.prec ? 8 : 16;
    uint32_t cap = E1;  // radix cap = table key
    std::vector<uint32_t> sched = regmix_schedule(n, cap);
    if (sched.empty() && job.prec) {  // doubles with a factor 11 or 13: only the 16-point tile takes them
        cap = 16;
        sched = regmix_schedule(n, cap);
    }
    if (sched.empty() || sched.size() > (size_t)RM_MAXP) return false;
    RmPlan pl;
    memset(&pl, 0, sizeof(pl));
    pl.npass = (uint32_t)sched.size();
    pl.cap = cap;
    auto tpl_for = [&](uint32_t e, bool &ok) {
        ok = true;
        uint32_t tpl = 1;
        for (auto R : sched) {
            if (R > e) { ok = false; return 0u; }
            const uint32_t nb = (uint32_t)(n / R), jmax = e / R;
            tpl = std::max(tpl, (nb + jmax - 1) / jmax);
        }
        return tpl;
    };
    bool ok1 = false, ok2 = false;
    uint32_t E = E1, TPL = tpl_for(E1, ok1);
    if (!ok1 || TPL > 512) {  // (measured: the small tile with one 512-thread CTA beats the double tile)
        E = 2 * E1;
        TPL = tpl_for(E, ok2);
        if (!ok2) return false;
    }
    const size_t esz = job.prec ? 16 : 8;
    const bool lf = load_lf || store_lf;
    uint32_t W;
    if (lf) {
        W = (uint32_t)(128 / esz);  // 128-byte rows of neighbouring lines
        if (W * TPL > 512) W = (uint32_t)(64 / esz);
        if (W * TPL > 512) W = (uint32_t)(32 / esz);  // one 32-byte sector per row: still beats the generic kernel
        if (W * TPL > 512) return false;
    } else {
        W = std::max<uint32_t>(1, 256 / TPL);
        if (TPL > 512) return false;
    }
    const uint64_t e0 = dims.empty() ? 1 : (uint64_t)dims[0].n;
    if (!lf) W = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(W, e0));
    pl.TPL = TPL;
    pl.d_TPL = make_fastdiv(TPL);
    uint32_t l1 = 1, twoff = 0;
    for (uint32_t i = 0; i < pl.npass; ++i) {
        const uint32_t R = sched[i];
        pl.R[i] = R;
        pl.ido[i] = (uint32_t)(n / ((uint64_t)l1 * R));
        pl.d_ido[i] = make_fastdiv(pl.ido[i]);
        const uint32_t nb = (uint32_t)(n / R);
        pl.J[i] = (nb + TPL - 1) / TPL;
        pl.twoff[i] = twoff;
        if (pl.ido[i] > 1) twoff += (R - 1) * pl.ido[i];
        l1 *= R;
    }
    pl.pitch = (uint32_t)n | 1u;
    if ((size_t)W * pl.pitch * esz > RM_MAX_SMEM) return false;
    // (unaligned complex arrays stay on the generic tile kernel: halves the build time of this file)
#ifdef RFB_RM_DOUBLE
    launch_regmix_typed<double, true>(job, dims, pl, W, E, load_lf, store_lf, s);
#else
    launch_regmix_typed<float, true>(job, dims, pl, W, E, load_lf, store_lf, s);
#endif
    return true;
}

}  // namespace rfb
