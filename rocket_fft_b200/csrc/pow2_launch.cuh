// Dispatch of a line job onto the instantiations of fft_pow2_kernel (one translation unit per
// precision, see pow2_launch_f32.cu / pow2_launch_f64.cu).
#pragma once
#include <stdio.h>
#include <stdlib.h>

#include "geom_fill.cuh"
#include "pow2_kernel.cuh"
#include "pow2_dual_kernel.cuh"
#include "pow2_pair_kernel.cuh"

namespace rfb {

// Long packed real lines (single precision, 16384 reals) as two interleaved half-length transforms per thread
// (pow2_dual_kernel.cuh).  RFB200_DUAL=0 switches back to the one-transform-per-line register kernel, 1 = every
// twiddle loaded from the tables, 2 (default) = twiddles composed from a few table entries.  Measured on B200,
// r2c of 16384 x 16384 float32 rows: 0.573 ms (57 % of the HBM copy peak) / 0.421 ms (78 %) / 0.403 ms (82 %).
inline int dual_variant() {
    static const int v = [] { const char *e = getenv("RFB200_DUAL"); return e ? atoi(e) : 2; }();
    return v;
}

template <int LOGNH, int MODE, bool TWC>
bool launch_dual_inst(const LineJob &job, const std::vector<Dim> &dims, cudaStream_t s) {
    using Body = DualBody<LOGNH, MODE, TWC>;
    // contiguous, fully present lines whose 16-byte {z[2e], z[2e+1]} / {y[2m], y[2m+1]} groups are aligned
    if (job.n_in != 0 && job.n_in != job.n) return false;
    if (MODE == 0) {
        // contiguous complex lines, plain load/store, 16-byte aligned on both sides
        const int64_t e = (int64_t)sizeof(float2);
        if (job.is != e || job.os != e || job.load_mode != LD_C2C || job.store_mode != ST_C2C || job.flags || job.twN ||
            job.pre_tab || job.post_tab || !job.split_out.empty() || job.conv)
            return false;
        if (((uint64_t)(uintptr_t)job.in % 16) != 0 || ((uint64_t)(uintptr_t)job.out % 16) != 0) return false;
        for (auto &d : dims) if (d.is % 16 || d.os % 16) return false;
    } else if (MODE == 1) {
        if (job.is != (int64_t)sizeof(float) || ((uint64_t)(uintptr_t)job.in % 16) != 0) return false;
        for (auto &d : dims) if (d.is % 16) return false;
    } else {
        if (job.is != (int64_t)sizeof(float2) || job.os != (int64_t)sizeof(float) || ((uint64_t)(uintptr_t)job.out % 16) != 0) return false;
        for (auto &d : dims) if (d.os % 16) return false;
    }
    TileGeom<float> g;
    LineJob j2 = job;
    if (MODE != 0) j2.n = job.n / 2;  // geometry in complex points
    const uint64_t ntiles = fill_geom<float>(g, j2, dims, 1u, false, false);
    if (MODE != 0) g.n_out = (uint32_t)(job.n / 2 + 1);
    g.twA = (const float2 *)get_table(TAB_LINE, job.prec, job.n, 0);
    if (MODE == 0) set_prefetch<float>(g, job, dims, 1u, sizeof(float2), job.n);
    else if (MODE == 1) {
        g.n_in = (uint32_t)job.n;
        set_prefetch<float>(g, job, dims, 1u, sizeof(float), job.n);
    } else set_prefetch<float>(g, job, dims, 1u, sizeof(float2), job.n / 2 + 1);
    const float2 *stw = (const float2 *)get_table(TAB_STOCKHAM, job.prec, 1ull << LOGNH, 0);
    const size_t smem = (size_t)Body::PITCH * sizeof(float4);
    auto kern = fft_pow2_dual_kernel<LOGNH, MODE, TWC>;
    static thread_local int dev_set = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev_set != dev) {
        RFB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dev_set = dev;
    }
    kern<<<(unsigned)ntiles, Body::NT, smem, s>>>(g, stw);
    {
        char nm[64];
        snprintf(nm, sizeof nm, "fft_pow2_dual_kernel<%d,%d,%s>", LOGNH, MODE, TWC ? "true" : "false");
        count_launch(nm);
    }
    RFB_CUDA_CHECK(cudaGetLastError());
    return true;
}

// lines per CTA: element-fast tiles aim at 256 threads, line-fast tiles at >= 128-byte rows
constexpr int p2_we(int logn) { return logn <= 10 ? (256 >> (logn - 4)) : (logn == 11 ? 2 : 1); }
// (measured on B200: 1024-point float lines 8 MiB apart: W=8 -> 2.9 TB/s, W=16 -> 3.7 TB/s;
//  128-point float lines of the four-step column passes: W=32 -> 0.82 ms, W=64 -> 0.88 ms per 16384^2 image;
//  1024-point double lines of the fused DCT, 64 contiguous doubles per row: W=8 -> 2.64 ms, W=4 -> 3.86 ms)
constexpr int p2_wl(int logn, bool dbl) {
    return logn <= 8 ? p2_we(logn) : (logn == 9 ? 16 : (logn == 10 ? (dbl ? 8 : 16) : (logn == 11 ? (dbl ? 4 : 8) : 0)));
}

template <typename T, int LOGN, int W, int MODE>
void launch_pow2_inst(const LineJob &job, const std::vector<Dim> &dims, bool load_lf, bool store_lf, cudaStream_t s) {
    using Body = Pow2Body<T, LOGN, W, MODE>;
    TileGeom<T> g;
    LineJob j2 = job;
    if (MODE != 0 && MODE != 5) j2.n = job.n / 2;  // geometry in complex points
    const uint64_t ntiles = fill_geom<T>(g, j2, dims, (uint32_t)W, load_lf, store_lf);
    if (MODE != 0 && MODE != 5) {
        g.n_out = (uint32_t)(job.n / 2 + 1);
        g.twA = (const cx<T> *)get_table(TAB_LINE, job.prec, job.n, 0);
    }
    if (MODE == 1) g.n_in = (uint32_t)(job.n_in ? job.n_in : job.n);  // real samples present
    if (MODE == 3 || MODE == 4) g.twB = (const cx<T> *)get_table(TAB_QUARTER, job.prec, job.n, 0);
    if (MODE == 2) {
        // the two reals of an output point are stored as one complex value when aligned
        const uint64_t csz = 2 * sizeof(T);
        bool al = ((uint64_t)(uintptr_t)job.out % csz) == 0;
        for (auto &d : dims) al = al && (d.os % (int64_t)csz) == 0;
        g.flags = al ? 0 : 1;
    }
    if (MODE == 0 && (LOGN == 4 || (LOGN == 5 && sizeof(T) == 4)) && W > 1 && !load_lf && !store_lf && !dims.empty()) {
        // short lines: dense tiles go through shared memory for coalescing (pow2_kernel.cuh).  Measured on B200, % of the
        // HBM copy peak without -> with staging: c64 n=16 28 -> 88, n=32 46 -> 66, n=64 67 -> 64; c128 n=16 53 -> 75,
        // n=32 63 -> 56, n=64 80 -> 57: only the first three are switched on.
        const bool stage = true;
        const int64_t e = (int64_t)sizeof(cx<T>), line = (int64_t)job.n * e;
        const bool plain_in = job.load_mode == LD_C2C && (job.n_in == 0 || job.n_in == job.n) && !job.pre_tab;
        const bool plain_out = job.store_mode == ST_C2C && job.twN == 0 && !job.post_tab && job.split_out.empty();
        if (stage && plain_in && job.is == e && dims[0].is == line) g.stage_io |= 1;
        if (stage && plain_out && job.os == e && dims[0].os == line) g.stage_io |= 2;
    }
    if (load_lf) {
        if (MODE == 0) set_prefetch_by_mode<T>(g, job, dims, (uint32_t)W);
        else if (MODE == 3 || MODE == 4) set_prefetch_rows<T>(g, job, dims, (uint32_t)W, sizeof(T), job.n);
    } else {
        // input items per line as they lie in memory: complex points (c2c), reals (r2c, DCT), Hermitian bins (c2r)
        if (MODE == 5) {
            LineJob jp = job;
            jp.pre_tab = nullptr;  // here the table multiplies the spectrum, every input element is read
            set_prefetch<T>(g, jp, dims, (uint32_t)W, sizeof(cx<T>), job.n);
        } else if (MODE == 0 && job.load_mode == LD_C2C) set_prefetch<T>(g, job, dims, (uint32_t)W, sizeof(cx<T>), job.n_in ? job.n_in : job.n);
        else if (MODE == 0 && job.load_mode == LD_REAL) set_prefetch<T>(g, job, dims, (uint32_t)W, sizeof(T), job.n_in ? job.n_in : job.n);
        else if (MODE == 1) set_prefetch<T>(g, job, dims, (uint32_t)W, sizeof(T), job.n_in ? job.n_in : job.n);
        else if (MODE == 3 || MODE == 4) set_prefetch<T>(g, job, dims, (uint32_t)W, sizeof(T), job.n);
        else if (MODE == 2) set_prefetch<T>(g, job, dims, (uint32_t)W, sizeof(cx<T>), job.n / 2 + 1);
    }
    const cx<T> *stw = (const cx<T> *)get_table(TAB_STOCKHAM, job.prec, 1ull << LOGN, 0);
    const size_t smem = (size_t)W * Body::PITCH * sizeof(cx<T>);
    auto kern = fft_pow2_kernel<T, LOGN, W, MODE>;
    static thread_local int dev_set = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev_set != dev) {
        RFB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dev_set = dev;
    }
    kern<<<(unsigned)ntiles, Body::NT, smem, s>>>(g, stw);
    {
        char nm[64];
        snprintf(nm, sizeof nm, "fft_pow2_kernel<%s,%d,%d,%d>", sizeof(T) == 8 ? "double" : "float", LOGN, W, MODE);
        count_launch(nm);
    }
    RFB_CUDA_CHECK(cudaGetLastError());
}

template <int LOGN, int W>
bool launch_pair_inst(const LineJob &job, const std::vector<Dim> &dims, cudaStream_t s);

// Strided float32 lines with adjacent neighbours (line-fast tiles) keep two lines per thread (pow2_pair_kernel.cuh).
// RFB200_PAIR=0 falls back to one line per thread (pow2_kernel.cuh).  (Round 1 also tried persistent CTAs staging the
// next tile with 8-byte cp.async -- slower, LDGSTS saturates the MIO path, profiles/r01e_* -- and 5 / 6 resident CTAs per
// SM -- slower, profiles/r02a_sweep_lf_ctas.log; both were removed.)
inline int lf_mode() {
    static const int v = [] {
        const char *p = getenv("RFB200_PAIR");
        return (p && atoi(p) == 0) ? 0 : 1;
    }();
    return v;
}
inline bool lf_plain_job(const LineJob &job, const std::vector<Dim> &dims) {
    const int64_t e = (int64_t)sizeof(float2);
    if (dims.empty() || dims[0].is != e || dims[0].os != e || (dims[0].tw && job.twN)) return false;
    // (scatter output -- the fused transform + NVLink push of the slab fftn -- included: the store picks its buffer per bin)
    return job.load_mode == LD_C2C && job.store_mode == ST_C2C && !job.flags && (job.n_in == 0 || job.n_in == job.n) &&
           !job.pre_tab && !job.post_tab && !job.conv;
}

template <typename T, int LOGN>
bool launch_pow2_logn(const LineJob &job, const std::vector<Dim> &dims, bool load_lf, bool store_lf, int mode,
                      cudaStream_t s) {
    constexpr bool dbl = sizeof(T) == 8;
    constexpr int WE = p2_we(LOGN), WL = p2_wl(LOGN, dbl);
    const bool lf = load_lf || store_lf;
    if (mode == 5) {
        if constexpr (LOGN >= 8 && LOGN <= 13) {
            if (lf) return false;
            launch_pow2_inst<T, LOGN, WE, 5>(job, dims, false, false, s);
            return true;
        } else return false;
    }
    if (mode == 1 || mode == 2) {
        if (lf) return false;
        if constexpr (sizeof(T) == 4 && LOGN == 13) {
            const int dv = dual_variant();
            if (dv == 1 && (mode == 1 ? launch_dual_inst<12, 1, false>(job, dims, s) : launch_dual_inst<12, 2, false>(job, dims, s)))
                return true;
            if (dv == 2 && (mode == 1 ? launch_dual_inst<12, 1, true>(job, dims, s) : launch_dual_inst<12, 2, true>(job, dims, s)))
                return true;
        }
        if (mode == 1) launch_pow2_inst<T, LOGN, WE, 1>(job, dims, load_lf, store_lf, s);
        else launch_pow2_inst<T, LOGN, WE, 2>(job, dims, load_lf, store_lf, s);
        return true;
    }
    if (mode == 3 || mode == 4) {
        // in place along a strided axis both sides are line-fast or both element-fast
        if (!lf) {
            if (mode == 3) launch_pow2_inst<T, LOGN, WE, 3>(job, dims, false, false, s);
            else launch_pow2_inst<T, LOGN, WE, 4>(job, dims, false, false, s);
            return true;
        }
        if constexpr (WL == 0) return false;
        else {
            if (mode == 3) launch_pow2_inst<T, LOGN, WL, 3>(job, dims, true, true, s);
            else launch_pow2_inst<T, LOGN, WL, 4>(job, dims, true, true, s);
            return true;
        }
    }
    if (!lf) {
        if constexpr (sizeof(T) == 4 && (LOGN == 13 || LOGN == 14)) {
            // long contiguous complex lines: two interleaved half-length transforms per thread.
            // Measured on B200, % of the HBM copy peak without -> with: n = 8192 62.1 -> 82.0 (cuFFT 75.2),
            // n = 16384 51.9 -> 56.3 (cuFFT 55.1)   (profiles/r01h_ab_dual_c2c.log)
            if (dual_variant() != 0 && launch_dual_inst<LOGN - 1, 0, true>(job, dims, s)) return true;
        }
        launch_pow2_inst<T, LOGN, WE, 0>(job, dims, load_lf, store_lf, s);
        return true;
    }
    if constexpr (WL == 0) return false;
    else {
        if constexpr (sizeof(T) == 4 && LOGN >= 7 && LOGN <= 10 && WL >= 2) {
            if (lf_mode() >= 1 && load_lf && store_lf && launch_pair_inst<LOGN, WL>(job, dims, s)) return true;
        }
        launch_pow2_inst<T, LOGN, WL, 0>(job, dims, load_lf, store_lf, s);
        return true;
    }
}

// Strided float32 lines whose neighbours are adjacent on both sides: two lines per thread (pow2_pair_kernel.cuh).
template <int LOGN, int W>
bool launch_pair_inst(const LineJob &job, const std::vector<Dim> &dims, cudaStream_t s) {
    using Body = PairBody<LOGN, W>;
    if (!lf_plain_job(job, dims)) return false;
    TileGeom<float> g;
    const uint64_t ntiles = fill_geom<float>(g, job, dims, (uint32_t)W, true, true);
    set_prefetch_by_mode<float>(g, job, dims, (uint32_t)W);
    const float2 *stw = (const float2 *)get_table(TAB_STOCKHAM, job.prec, 1ull << LOGN, 0);
    const size_t smem = (size_t)Body::WP * Body::PITCH * sizeof(float4);
    // 16-byte accesses when everything is a multiple of 16 bytes (the half spectrum of rfft2, rows of 8193 points, is not)
    bool al16 = (((uintptr_t)job.in | (uintptr_t)job.out | (uintptr_t)job.is | (uintptr_t)job.os) & 15) == 0;
    for (size_t d = 1; d < dims.size(); ++d) al16 = al16 && ((dims[d].is | dims[d].os) & 15) == 0;
    for (auto sp : job.split_out) al16 = al16 && ((uintptr_t)sp & 15) == 0;
    void (*kern)(const TileGeom<float>, const float2 *) = al16 ? fft_pow2_pair_kernel<LOGN, W, true> : fft_pow2_pair_kernel<LOGN, W, false>;
    static thread_local int dev_set[2] = {-1, -1};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev_set[al16] != dev) {
        RFB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dev_set[al16] = dev;
    }
    kern<<<(unsigned)ntiles, Body::NT, smem, s>>>(g, stw);
    {
        char nm[64];
        snprintf(nm, sizeof nm, "fft_pow2_pair_kernel<%d,%d>", LOGN, W);
        count_launch(nm);
    }
    RFB_CUDA_CHECK(cudaGetLastError());
    return true;
}

template <typename T, int MAXLOG>
bool launch_pow2_any(const LineJob &job, const std::vector<Dim> &dims, bool load_lf, bool store_lf, cudaStream_t s) {
    if (dims.size() > (size_t)MAXB) return false;
    // which flavour?
    int mode = 0;
    uint64_t n = job.n;
    const bool plain_in = job.n_in == 0 || job.n_in == job.n;
    if (job.conv) {
        if (job.load_mode != LD_C2C || job.store_mode != ST_C2C || !job.pre_tab || job.post_tab || !plain_in || load_lf ||
            store_lf || !job.split_out.empty() || job.flags)
            return false;
        mode = 5;
    } else
    if (job.load_mode == LD_REAL && job.store_mode == ST_HALF && job.flags == 0 && job.twN == 0 &&
        (n % 2 == 0) && !load_lf && !store_lf && n >= 32) {
        // the packed load reads two reals as one complex value: needs complex alignment
        const uint64_t csz = 2 * sizeof(T);
        bool al = ((uint64_t)(uintptr_t)job.in % csz) == 0;
        for (auto &d : dims) al = al && (d.is % (int64_t)csz) == 0;
        if (!al && job.is == (int64_t)sizeof(T)) return false;
        mode = 1;
        n /= 2;
    } else if (job.load_mode == LD_HERM && job.store_mode == ST_REAL && job.flags == 0 && plain_in && job.twN == 0 &&
               (n % 2 == 0) && !load_lf && !store_lf && n >= 32) {
        mode = 2;
        n /= 2;
    } else if ((job.load_mode == LD_DCT2 && job.store_mode == ST_DCT2) || (job.load_mode == LD_DCT3 && job.store_mode == ST_DCT3)) {
        if (n % 2 || n < 32 || job.twN) return false;
        mode = job.load_mode == LD_DCT2 ? 3 : 4;
        n /= 2;
    } else if (job.store_mode == ST_HC || job.load_mode >= LD_DCT2 || job.store_mode >= ST_DCT2) return false;
    if (!job.split_out.empty() && (mode != 0 || job.store_mode != ST_C2C)) return false;
    if ((job.pre_tab || job.post_tab) && ((mode != 0 && mode != 5) || job.store_mode != ST_C2C || job.load_mode != LD_C2C)) return false;
    if (n < 16 || (n & (n - 1))) return false;
    int logn = 0;
    while ((1ull << logn) < n) ++logn;
    if (logn > MAXLOG) return false;
    switch (logn) {
#define RFB_P2_CASE(L) case L: if constexpr (L <= MAXLOG) return launch_pow2_logn<T, L>(job, dims, load_lf, store_lf, mode, s); else return false;
        RFB_P2_CASE(4) RFB_P2_CASE(5) RFB_P2_CASE(6) RFB_P2_CASE(7) RFB_P2_CASE(8) RFB_P2_CASE(9) RFB_P2_CASE(10)
        RFB_P2_CASE(11) RFB_P2_CASE(12) RFB_P2_CASE(13) RFB_P2_CASE(14)
#undef RFB_P2_CASE
    }
    return false;
}

}  // namespace rfb
