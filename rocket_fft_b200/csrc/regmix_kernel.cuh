// Register-resident mixed-radix Stockham kernel for smooth lengths (prime factors <= 13).
//
// Same idea as pow2_kernel.cuh -- the first pass reads global memory straight into registers, the
// last pass writes registers straight to global memory (both coalesced: thread t touches t + m*n/R),
// and between passes the line makes one trip through shared memory -- but the radix of every pass is
// a run-time choice from {2,...,13,15,16} (6, 9, 10, 12, 15 as nested butterflies).  A thread owns up to 16 points: J = ceil((n/R)/TPL)
// butterflies of radix R per pass (J*R <= 16), butterfly b = t + j*TPL.  TPL (threads per line) is
// chosen by the host so that every pass fits.  Covers the lengths the reference's cfftp handles
// with pass2/3/4/5/7/8/11 (rocket_fft/_pocketfft_hdronly.h:1079-1573) plus 13 and 16.
#pragma once
#include "common.cuh"
#include "line_io.cuh"
#include "radix.cuh"
#include "tile_kernel.cuh"

namespace rfb {

// points per thread E: 16 (short lines, small CTAs) or 32 (long lines); doubles use half of that
constexpr int RM_MAXP = 12;    // passes

struct RmPlan {
    uint32_t npass, TPL, cap;
    FastDiv d_TPL;
    uint32_t R[RM_MAXP], ido[RM_MAXP], J[RM_MAXP], twoff[RM_MAXP];
    FastDiv d_ido[RM_MAXP];
    uint32_t pitch;
};

template <typename T>
struct RmCtx {
    uint32_t w_first, wvalid, i1, i2;
    int64_t in_base, out_base;
    cx<T> *buf;
    uint32_t tid;
};

template <typename T, int R, bool ALIGNED, int E>
__device__ __forceinline__ void rm_pass(const TileGeom<T> &g, const RmPlan &pl, const RmCtx<T> &c, uint32_t s,
                                     bool &need_sync) {
    using C = cx<T>;
    constexpr int JMAX = E / R;
    C v[JMAX * R];
    const uint32_t n = g.n, nb = n / R, ido = pl.ido[s], J = pl.J[s], TPL = pl.TPL;
    const bool first = (s == 0), last = (s + 1 == pl.npass);
    const bool lf = first ? (g.load_line_fast != 0) : (g.store_line_fast != 0);
    uint32_t w, t;
    if (lf) fdivmod(c.tid, g.d_W, t, w);
    else fdivmod(c.tid, pl.d_TPL, w, t);
    const bool wok = w < c.wvalid;
    C *sl = c.buf + (size_t)w * pl.pitch;
    const C *tws = g.ptw + pl.twoff[s];
    if (first) {
        // ---- global -> registers: all loads of all of this thread's butterflies first (l1 == 1: i == b) ----
        const char *line = g.in + c.in_base + (int64_t)w * g.in_bs[0];
        const bool plain = g.load_mode == LD_C2C && g.n_in == n;
#pragma unroll
        for (int j = 0; j < JMAX; ++j) {
            const uint32_t b = t + j * TPL;
            const bool act = ((uint32_t)j < J) && (b < nb) && wok;
#pragma unroll
            for (int m = 0; m < R; ++m) {
                C val = mk<T>(T(0), T(0));
                if (act) {
                    const uint32_t e = b + ido * m;
                    if (plain) val = ld_cx<T, ALIGNED>(line + (int64_t)e * g.in_sa);
                    else val = load_value<T, ALIGNED>(g.load_mode, g.flags, line, g.in_sa, e, n, g.n_in);
                }
                v[j * R + m] = val;
            }
        }
        if (g.backward) {
#pragma unroll
            for (int q = 0; q < JMAX * R; ++q) v[q] = cswap(v[q]);
        }
#pragma unroll
        for (int j = 0; j < JMAX; ++j) {
            const uint32_t b = t + j * TPL;
            if (((uint32_t)j < J) && (b < nb)) {
                Dft<T, R>::run(v + j * R);
                if (ido > 1 && b > 0) {
#pragma unroll
                    for (int q = 1; q < R; ++q) v[j * R + q] = cmul(v[j * R + q], __ldg(tws + b + (q - 1) * ido));
                }
            }
        }
    } else {
        // ---- shared -> registers, one butterfly after the other ----
#pragma unroll
        for (int j = 0; j < JMAX; ++j) {
            const uint32_t b = t + j * TPL;
            if (((uint32_t)j < J) && (b < nb)) {
                uint32_t k, i;
                fdivmod(b, pl.d_ido[s], k, i);
                const C *src = sl + i + ido * R * k;
#pragma unroll
                for (int m = 0; m < R; ++m) v[j * R + m] = src[ido * m];
                Dft<T, R>::run(v + j * R);
                if (ido > 1 && i > 0) {
#pragma unroll
                    for (int q = 1; q < R; ++q) v[j * R + q] = cmul(v[j * R + q], __ldg(tws + i + (q - 1) * ido));
                }
            }
        }
    }
    // ---- outputs: bin / slot b + q * n/R ------------------------------------------------------------
    if (last) {
        if (!wok) return;
        char *line = g.out + c.out_base + (int64_t)w * g.out_bs[0];
        const uint32_t cc = (g.tw_dim == 0) ? (c.w_first + w) : (g.tw_dim == 1 ? c.i1 : c.i2);
        const bool plain = g.store_mode == ST_C2C && g.tw_dim < 0;
#pragma unroll
        for (int j = 0; j < JMAX; ++j) {
            const uint32_t b = t + j * TPL;
            if (((uint32_t)j < J) && (b < nb)) {
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const uint32_t f = b + q * nb;
                    C val = v[j * R + q];
                    if (g.tw_dim >= 0) {
                        uint32_t hi, lo;
                        fdivmod(cc * f, g.d_twS, hi, lo);
                        val = cmul(val, cmul(__ldg(g.twA + hi), __ldg(g.twB + lo)));
                    }
                    val = cscale(val, g.fct);
                    if (g.backward) val = cswap(val);
                    if (plain) st_cx<T, ALIGNED>(line + (int64_t)f * g.out_sa, val);
                    else store_bin_value<T, ALIGNED>(g.store_mode, g.flags, line, g.out_sa, f, n, val);
                }
            }
        }
    } else {
        if (need_sync) __syncthreads();  // everybody has read the previous pass's data
#pragma unroll
        for (int j = 0; j < JMAX; ++j) {
            const uint32_t b = t + j * TPL;
            if (((uint32_t)j < J) && (b < nb)) {
#pragma unroll
                for (int q = 0; q < R; ++q) sl[b + q * nb] = v[j * R + q];
            }
        }
        __syncthreads();
        need_sync = true;
    }
}

// Passes are grouped by radix in the fixed order 16,8,4,2,12,10,6,15,13,11,9,7,5,3 (regmix_schedule in
// plan.cpp): one run-time loop per radix keeps every unrolled pass body -- and its register tile -- in a
// straight-line region of the kernel (a switch inside a loop over passes made ptxas demote the tiles to
// local memory).
template <typename T, bool ALIGNED, int E, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) fft_regmix_kernel(const TileGeom<T> g, const RmPlan pl) {
    using C = cx<T>;
    extern __shared__ __align__(16) unsigned char smem_raw_rm[];
    RmCtx<T> c;
    c.buf = reinterpret_cast<C *>(smem_raw_rm);
    c.tid = threadIdx.x;
    uint32_t t0, rest;
    fdivmod(blockIdx.x, g.d_t0, rest, t0);
    fdivmod(rest, g.d_e1, c.i2, c.i1);
    c.w_first = t0 * g.W;
    prefetch_later_tile<T>(g, g.W);
    c.wvalid = min(g.W, g.bext[0] - c.w_first);
    c.in_base = (int64_t)c.w_first * g.in_bs[0] + (int64_t)c.i1 * g.in_bs[1] + (int64_t)c.i2 * g.in_bs[2];
    c.out_base = (int64_t)c.w_first * g.out_bs[0] + (int64_t)c.i1 * g.out_bs[1] + (int64_t)c.i2 * g.out_bs[2];
    bool need_sync = false;
    uint32_t s = 0;
#define RFB_RM_RUN(RR)                                           \
    while (s < pl.npass && pl.R[s] == RR) {                      \
        rm_pass<T, RR, ALIGNED, E>(g, pl, c, s, need_sync);          \
        ++s;                                                     \
    }
    if constexpr (E >= 16) { RFB_RM_RUN(16) }
    RFB_RM_RUN(8) RFB_RM_RUN(4) RFB_RM_RUN(2)
    if constexpr (E >= 16) { RFB_RM_RUN(12) RFB_RM_RUN(10) }
    RFB_RM_RUN(6)
    if constexpr (E >= 16) { RFB_RM_RUN(15) RFB_RM_RUN(13) RFB_RM_RUN(11) RFB_RM_RUN(9) }
    RFB_RM_RUN(7) RFB_RM_RUN(5) RFB_RM_RUN(3)
#undef RFB_RM_RUN
}

}  // namespace rfb
