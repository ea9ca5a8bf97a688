// Strided power-of-two lines (single precision), TWO neighbouring lines per thread.
//
// Line-fast tiles of the register kernel (pow2_kernel.cuh) -- the column passes of a 2-D transform, the outer axes of
// a volume, both steps of the four-step split -- are limited by instruction issue and load latency, not by DRAM (ncu on
// the 128-point column passes of rfft2 16384^2: issue slots 55-66 % busy, DRAM 62-65 %, ~58-73 instructions per point
// of which less than half are butterflies).  The W lines of such a tile are adjacent in memory, so a thread can own
// the same 16 points of two neighbouring lines: one address computation, one bounds predicate, one twiddle load and
// one four-step factor serve both lines, and the exchanges move 16-byte {line 2p, line 2p+1} pairs (LDS.128/STS.128).
// The lines need only complex (8-byte) alignment: global accesses stay 8 bytes wide, the second at offset +8.
// (Counterpart of general_nd + copy_input/copy_output with vlen = 2 lines in the reference,
// _pocketfft_hdronly.h:3496-3607; different algorithm.)
#pragma once
#include "pow2_kernel.cuh"

namespace rfb {

template <int LOGN, int W>
struct PairBody {
    using T = float;
    using C = float2;
    using PL = P2<LOGN>;
    static constexpr int N = PL::N, TPL = PL::TPL, WP = W / 2, NT = WP * TPL;
    static constexpr int PITCH = (N + PL::PAD) | 1;  // in 16-byte pairs; odd: neighbouring pairs start in different banks

    template <int P>
    static __device__ __forceinline__ void compute2(C *a, C *b, int t, const C *__restrict__ stw) {
        constexpr int R = PL::radix(P), NB = 16 / R, ido = PL::ido(P);
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            Dft<T, R>::run(a + j * R);
            Dft<T, R>::run(b + j * R);
        }
        if constexpr (ido > 1) {
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const int i = (t + j * TPL) % ido;
                const C *tw = stw + PL::twoff(P) + i;
#pragma unroll
                for (int q = 1; q < R; ++q) {
                    const C w = __ldg(tw + (q - 1) * ido);  // one load serves both lines
                    a[j * R + q] = cmul(a[j * R + q], w);
                    b[j * R + q] = cmul(b[j * R + q], w);
                }
            }
        }
    }

    template <int P>
    static __device__ __forceinline__ void exchange2(C *a, C *b, float4 *line, int t, bool first) {
        constexpr int Rp = PL::radix(P - 1), NBp = 16 / Rp;
        constexpr int ido = PL::ido(P);
        if (!first) __syncthreads();
#pragma unroll
        for (int j = 0; j < NBp; ++j)
#pragma unroll
            for (int q = 0; q < Rp; ++q)
                line[p2_phys<LOGN, P>(t + j * TPL + q * (N / Rp))] =
                    make_float4(a[j * Rp + q].x, a[j * Rp + q].y, b[j * Rp + q].x, b[j * Rp + q].y);
        __syncthreads();
        const int i = t % ido, k = t / ido;
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const float4 u = line[p2_phys<LOGN, P>(i + ido * (m + 16 * k))];
            a[m] = mk<T>(u.x, u.y);
            b[m] = mk<T>(u.z, u.w);
        }
    }

    // AL16: both arrays and all their strides are multiples of 16 bytes -- the pair {line 2p, line 2p+1} moves as ONE 16-byte
    // access (ncu on 1024-point lines 4 MiB apart: 21 % of the stall samples were LG throttle, the load/store queue full of
    // 8-byte requests)
    template <bool AL16>
    static __device__ __forceinline__ void run(const TileGeom<T> &g, const C *__restrict__ stw, float4 *buf) {
        uint32_t t0, i1, i2, rest;
        fdivmod(blockIdx.x, g.d_t0, rest, t0);
        fdivmod(rest, g.d_e1, i2, i1);
        const uint32_t w_first = t0 * W;
        const int wvalid = (int)min((uint32_t)W, g.bext[0] - w_first);
        const int tid = threadIdx.x;
        const int wp = tid % WP, t = tid / WP;  // pair of lines, butterfly
        const bool ok0 = 2 * wp < wvalid, ok1 = 2 * wp + 1 < wvalid;
        // neighbouring lines are sizeof(C) apart on both sides (checked by the launcher)
        const int64_t in_base = (int64_t)(w_first + 2 * wp) * (int64_t)sizeof(C) + (int64_t)i1 * g.in_bs[1] + (int64_t)i2 * g.in_bs[2];
        const int64_t out_base = (int64_t)(w_first + 2 * wp) * (int64_t)sizeof(C) + (int64_t)i1 * g.out_bs[1] + (int64_t)i2 * g.out_bs[2];
        C a[16], b[16];
        prefetch_later_tile<T>(g, (uint32_t)W);

        // ---- pass 0: global -> registers, all loads issued back to back --------------------------------------------
        {
            constexpr int R = PL::radix(0), NB = 16 / R, ido = PL::ido(0);
            const int64_t sa = g.in_sa;
            const int64_t step_m = (int64_t)ido * sa, step_j = (int64_t)TPL * sa;
            const char *pj = g.in + in_base + (int64_t)t * sa;
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const char *pm = pj;
#pragma unroll
                for (int m = 0; m < R; ++m) {
                    if (AL16 && ok1) {
                        const float4 u = *reinterpret_cast<const float4 *>(pm);
                        a[j * R + m] = mk<T>(u.x, u.y);
                        b[j * R + m] = mk<T>(u.z, u.w);
                    } else {
                        a[j * R + m] = ok0 ? *reinterpret_cast<const C *>(pm) : mk<T>(T(0), T(0));
                        b[j * R + m] = ok1 ? *reinterpret_cast<const C *>(pm + sizeof(C)) : mk<T>(T(0), T(0));
                    }
                    pm += step_m;
                }
                pj += step_j;
            }
            if (g.backward) {
#pragma unroll
                for (int i = 0; i < 16; ++i) { a[i] = cswap(a[i]); b[i] = cswap(b[i]); }
            }
        }
        float4 *line = buf + wp * PITCH;
        compute2<0>(a, b, t, stw);
        if constexpr (PL::NPASS > 1) { exchange2<1>(a, b, line, t, true); compute2<1>(a, b, t, stw); }
        if constexpr (PL::NPASS > 2) { exchange2<2>(a, b, line, t, false); compute2<2>(a, b, t, stw); }
        if constexpr (PL::NPASS > 3) { exchange2<3>(a, b, line, t, false); compute2<3>(a, b, t, stw); }

        // ---- thread t holds bins t + j*TPL + q*N/RL of both lines: strided stores, optionally with the four-step
        //      factor exp(-2 pi i c k / bigN) (exact two-level look-up for every 4th bin, recurrence in between) ----------
        if (!ok0) return;
        constexpr int RL = PL::radix(PL::NPASS - 1), NBL = 16 / RL;
        const T f = g.fct;
        const bool bw = g.backward != 0;
        const bool tw = g.tw_dim >= 0;
        const uint32_t c = tw ? ((g.tw_dim == 0) ? 0u : (g.tw_dim == 1 ? i1 : i2)) : 0u;  // (dim 0 is the tile dim: never the factor's)
        auto lookup = [&](uint32_t x) {
            uint32_t hi, lo;
            fdivmod(x, g.d_twS, hi, lo);
            return cmul(__ldg(g.twA + hi), __ldg(g.twB + lo));
        };
        C step = mk<T>(T(1), T(0));
        if (tw) step = lookup(c * (uint32_t)(N / RL));
        const int64_t step_q = (int64_t)(N / RL) * g.out_sa, step_j = (int64_t)TPL * g.out_sa;
        char *pj = g.out + out_base + (int64_t)t * g.out_sa;
#pragma unroll
        for (int j = 0; j < NBL; ++j) {
            char *pq = pj;
            C wq = mk<T>(T(1), T(0));
#pragma unroll
            for (int q = 0; q < RL; ++q) {
                C va = a[j * RL + q], vb = b[j * RL + q];
                if (tw) {
                    if ((q & 3) == 0) wq = lookup(c * (uint32_t)(t + j * TPL + q * (N / RL)));
                    else wq = cmul(wq, step);
                    va = cmul(va, wq);
                    vb = cmul(vb, wq);
                }
                va = cscale(va, f);
                vb = cscale(vb, f);
                if (bw) { va = cswap(va); vb = cswap(vb); }
                char *dst = pq;
                // fused exchange (rfb200_c2c_scatter): the bin's block decides which (peer) buffer receives it
                if (g.split_blk) dst = g.split_base[fdiv((uint32_t)(t + j * TPL + q * (N / RL)), g.d_split)] + (pq - g.out);
                if (AL16 && ok1) *reinterpret_cast<float4 *>(dst) = make_float4(va.x, va.y, vb.x, vb.y);
                else {
                    *reinterpret_cast<C *>(dst) = va;
                    if (ok1) *reinterpret_cast<C *>(dst + sizeof(C)) = vb;
                }
                pq += step_q;
            }
            pj += step_j;
        }
    }
};

template <int LOGN, int W, bool AL16>
__global__ void __launch_bounds__((W / 2) * (1 << LOGN) / 16, 512 / ((W / 2) * (1 << LOGN) / 16))
    fft_pow2_pair_kernel(const TileGeom<float> g, const float2 *__restrict__ stw) {
    extern __shared__ __align__(16) unsigned char smem_raw_p2p[];
    PairBody<LOGN, W>::template run<AL16>(g, stw, reinterpret_cast<float4 *>(smem_raw_p2p));
}

}  // namespace rfb
