// Strided power-of-two lines (single precision, 128 / 256 points): persistent CTAs whose NEXT tile is already on its way
// into shared memory (cp.async) while the current tile is transformed.
//
// ncu on the line-fast tiles of pow2_kernel.cuh / pow2_pair_kernel.cuh (the column passes of rfft2 16384^2) shows them
// latency-bound: 52 % of the stall samples wait on the tile's own global loads, DRAM is 62 % busy, issue slots 44-66 %.
// A CTA that loads into registers can only have its own tile in flight, and only during its load phase (about 128 KB per
// SM for a third of the time).  Here every thread copies ITS OWN 16 points of the next tile asynchronously into a
// landing buffer (8-byte cp.async: the rows of a half spectrum are only 8-byte aligned, which rules out TMA boxes and
// 16-byte copies), so no barrier is needed to consume them -- cp.async.wait_group is per thread -- and three CTAs per SM
// keep three tiles (96 KB) in flight all the time.  The landing buffer of tile i doubles as its exchange buffer; two
// buffers per CTA.
#pragma once
#include "pow2_kernel.cuh"

namespace rfb {

__device__ __forceinline__ void cp_async8(uint32_t smem_addr, const void *gptr) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

template <int LOGN, int W>
struct AsyncLfBody {
    using T = float;
    using C = float2;
    using PL = P2<LOGN>;
    using PB = Pow2Body<float, LOGN, W, 0>;
    static constexpr int N = PL::N, TPL = PL::TPL, NT = W * TPL, PITCH = PB::PITCH;
    static constexpr int BUF = W * PITCH;  // complex slots per buffer (>= W*N: the landing layout [e][w] fits)
    static constexpr int R0 = PL::radix(0), NB0 = 16 / R0, IDO0 = PL::ido(0);

    struct Ctx {
        int64_t in_base, out_base;
        uint32_t i1, i2;
        int wvalid;
    };
    static __device__ __forceinline__ Ctx locate(const TileGeom<T> &g, uint32_t tile) {
        uint32_t t0, rest;
        Ctx c;
        fdivmod(tile, g.d_t0, rest, t0);
        fdivmod(rest, g.d_e1, c.i2, c.i1);
        const uint32_t w_first = t0 * W;
        c.wvalid = (int)min((uint32_t)W, g.bext[0] - w_first);
        c.in_base = (int64_t)w_first * g.in_bs[0] + (int64_t)c.i1 * g.in_bs[1] + (int64_t)c.i2 * g.in_bs[2];
        c.out_base = (int64_t)w_first * g.out_bs[0] + (int64_t)c.i1 * g.out_bs[1] + (int64_t)c.i2 * g.out_bs[2];
        return c;
    }

    // this thread's own 16 points of `tile` -> stage[e*W + w], asynchronously; one commit group per call
    static __device__ __forceinline__ void issue(const TileGeom<T> &g, uint32_t tile, C *stage, int w, int t) {
        const Ctx c = locate(g, tile);
        if (w < c.wvalid) {
            const int64_t sa = g.in_sa;
            const int64_t step_m = (int64_t)IDO0 * sa, step_j = (int64_t)TPL * sa;
            const char *pj = g.in + c.in_base + (int64_t)w * g.in_bs[0] + (int64_t)t * sa;
            const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(stage + t * W + w);
#pragma unroll
            for (int j = 0; j < NB0; ++j) {
                const char *pm = pj;
#pragma unroll
                for (int m = 0; m < R0; ++m) {
                    cp_async8(s0 + (uint32_t)((j * TPL + m * IDO0) * W * (int)sizeof(C)), pm);
                    pm += step_m;
                }
                pj += step_j;
            }
        }
        cp_async_commit();
    }

    static __device__ __forceinline__ void run(const TileGeom<T> &g, const C *__restrict__ stw, C *smem, uint32_t ntiles) {
        const int tid = threadIdx.x;
        const int w = tid % W, t = tid / W;  // line-fast mapping, as in Pow2Body
        uint32_t tile = blockIdx.x;
        if (tile < ntiles) issue(g, tile, smem, w, t);
        for (uint32_t it = 0; tile < ntiles; ++it, tile += gridDim.x) {
            C *cur = smem + (it & 1u) * BUF, *nxt = smem + ((it & 1u) ^ 1u) * BUF;
            if (tile + gridDim.x < ntiles) issue(g, tile + gridDim.x, nxt, w, t);
            else cp_async_commit();  // an empty group keeps the wait count uniform
            cp_async_wait<1>();      // everything but the newest group has landed: this thread's points of `tile`
            const Ctx c = locate(g, tile);
            const bool wok = w < c.wvalid;
            C v[16];
#pragma unroll
            for (int j = 0; j < NB0; ++j)
#pragma unroll
                for (int m = 0; m < R0; ++m)
                    v[j * R0 + m] = wok ? cur[(t + j * TPL + m * IDO0) * W + w] : mk<T>(T(0), T(0));
            if (g.backward) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = cswap(v[i]);
            }
            PB::template compute<0>(v, t, stw);
            // the landing buffer becomes the exchange buffer (first = false: a barrier separates the two uses)
            if constexpr (PL::NPASS > 1) { PB::template exchange<1>(v, cur, tid, true, true, false); PB::template compute<1>(v, t, stw); }
            if constexpr (PL::NPASS > 2) { PB::template exchange<2>(v, cur, tid, true, true, false); PB::template compute<2>(v, t, stw); }
            if constexpr (PL::NPASS > 3) { PB::template exchange<3>(v, cur, tid, true, true, false); PB::template compute<3>(v, t, stw); }
            __syncthreads();  // every thread has read its exchange data: `cur` may receive the tile after the next

            // ---- strided stores, optionally with the four-step factor exp(-2 pi i c k / bigN) --------------------------
            if (wok) {
                constexpr int RL = PL::radix(PL::NPASS - 1), NBL = 16 / RL;
                const T f = g.fct;
                const bool bw = g.backward != 0;
                const bool tw = g.tw_dim >= 0;
                const uint32_t cc = tw ? ((g.tw_dim == 1) ? c.i1 : c.i2) : 0u;  // (the tile dim never carries the factor here)
                auto lookup = [&](uint32_t x) {
                    uint32_t hi, lo;
                    fdivmod(x, g.d_twS, hi, lo);
                    return cmul(__ldg(g.twA + hi), __ldg(g.twB + lo));
                };
                C step = mk<T>(T(1), T(0));
                if (tw) step = lookup(cc * (uint32_t)(N / RL));
                const int64_t step_q = (int64_t)(N / RL) * g.out_sa, step_j = (int64_t)TPL * g.out_sa;
                char *pj = g.out + c.out_base + (int64_t)w * g.out_bs[0] + (int64_t)t * g.out_sa;
#pragma unroll
                for (int j = 0; j < NBL; ++j) {
                    char *pq = pj;
                    C wq = mk<T>(T(1), T(0));
#pragma unroll
                    for (int q = 0; q < RL; ++q) {
                        C val = v[j * RL + q];
                        if (tw) {
                            if ((q & 3) == 0) wq = lookup(cc * (uint32_t)(t + j * TPL + q * (N / RL)));
                            else wq = cmul(wq, step);
                            val = cmul(val, wq);
                        }
                        val = cscale(val, f);
                        if (bw) val = cswap(val);
                        *reinterpret_cast<C *>(pq) = val;
                        pq += step_q;
                    }
                    pj += step_j;
                }
            }
        }
        cp_async_wait<0>();
    }
};

template <int LOGN, int W>
__global__ void __launch_bounds__(W *(1 << LOGN) / 16, 3)
    fft_pow2_async_kernel(const TileGeom<float> g, const float2 *__restrict__ stw, uint32_t ntiles) {
    extern __shared__ __align__(16) unsigned char smem_raw_p2a[];
    AsyncLfBody<LOGN, W>::run(g, stw, reinterpret_cast<float2 *>(smem_raw_p2a), ntiles);
}

}  // namespace rfb
