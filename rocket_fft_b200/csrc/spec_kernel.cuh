// Fully specialised register-resident Stockham kernel: length, radix schedule, threads per line and
// lines per CTA are compile-time constants of a plan type P, so every index computation folds to
// shifts / constant multiplies and every pass is straight-line code (the structure of pow2_kernel.cuh for
// arbitrary smooth lengths).  Instantiated at run time through NVRTC (jit.cu) for the length at hand --
// the same source a build-time instantiation would use -- with regmix_kernel.cuh as the fallback.
//
// P must provide:  N, TPL, W, NPASS, MINB (static constexpr int) and static constexpr int radix(int s).
// MODE 0: complex line job (all load / store modes of line_io.cuh);  MODE 1: packed real -> half spectrum of a
// contiguous even-length real line (N = half the real length), the structure of MODE 1 in pow2_kernel.cuh;
// MODE 2: its inverse, N+1 Hermitian bins -> even-length real line (MODE 2 of pow2_kernel.cuh).
#pragma once
#include "common.cuh"
#include "line_io.cuh"
#include "radix.cuh"
#include "tile_kernel.cuh"

namespace rfb {

template <typename P>
struct SpecInfo {
    __host__ __device__ static constexpr int l1(int s) {
        int l = 1;
        for (int i = 0; i < s; ++i) l *= P::radix(i);
        return l;
    }
    __host__ __device__ static constexpr int ido(int s) { return P::N / (l1(s) * P::radix(s)); }
    __host__ __device__ static constexpr int twoff(int s) {
        int o = 0;
        for (int i = 0; i < s; ++i)
            if (ido(i) > 1) o += (P::radix(i) - 1) * ido(i);
        return o;
    }
    __host__ __device__ static constexpr int J(int s) { return (P::N / P::radix(s) + P::TPL - 1) / P::TPL; }
};

template <typename T, typename P, int S, bool ALIGNED, int MODE>
struct SpecPass {
    using C = cx<T>;
    using I = SpecInfo<P>;
    static constexpr int N = P::N, TPL = P::TPL, W = P::W, R = P::radix(S);
    static constexpr int NB = N / R, IDO = I::ido(S), J = I::J(S);
    static constexpr bool FIRST = S == 0, LAST = S == P::NPASS - 1, EXACT = (J * TPL == NB);
    static constexpr int PITCH = (W == 1) ? N : (N | 1);

    static __device__ __forceinline__ void run(const TileGeom<T> &g, C *buf, int tid, uint32_t w_first, int wvalid,
                                               uint32_t i1, uint32_t i2, int64_t in_base, int64_t out_base) {
        C v[J * R];
        const bool lf = FIRST ? (g.load_line_fast != 0) : (g.store_line_fast != 0);
        int w, t;
        if (lf) { w = tid % W; t = tid / W; }
        else { w = tid / TPL; t = tid % TPL; }
        const bool wok = (W == 1) ? true : (w < wvalid);
        C *sl = buf + w * PITCH;
        const C *tws = g.ptw + I::twoff(S);
        if constexpr (FIRST) {
            const char *line = g.in + in_base + (int64_t)w * g.in_bs[0];
            const bool plain = MODE == 0 && g.load_mode == LD_C2C && g.n_in == (uint32_t)N;
            if (MODE == 1) {
                // packed real transform: the contiguous real line (n_in samples present, the rest zero padding) is
                // read as N complex points x[2e] + i x[2e+1]; the host guarantees complex alignment
                const C *p = reinterpret_cast<const C *>(line) + t;
                const bool full = g.n_in == 2u * (uint32_t)N;
#pragma unroll
                for (int j = 0; j < J; ++j)
#pragma unroll
                    for (int m = 0; m < R; ++m) {
                        const int e = t + j * TPL + m * IDO;
                        C val = mk<T>(T(0), T(0));
                        if (wok && (EXACT || t + j * TPL < NB)) {
                            if (full || 2u * (uint32_t)e + 1u < g.n_in) val = __ldcs(p + j * TPL + m * IDO);
                            else if (2u * (uint32_t)e < g.n_in) val.x = *reinterpret_cast<const T *>(line + (int64_t)(2 * e) * sizeof(T));
                        }
                        v[j * R + m] = val;
                    }
            } else if (MODE == 2) {
                // packed inverse real transform: the N+1 Hermitian bins X are folded into the N-point spectrum
                // Z[e] = (X[e] + conj X[N-e]) + i w^e (X[e] - conj X[N-e]), w = exp(+2 pi i/(2N)), whose backward DFT is
                // x[2m] + i x[2m+1].  Im X[0], Im X[N] are ignored (H:3830, 3845-3846); forward=true: X -> conj X.
                const bool cj = g.backward == 0;
                const int64_t sa = g.in_sa;
#pragma unroll
                for (int j = 0; j < J; ++j)
#pragma unroll
                    for (int m = 0; m < R; ++m) {
                        const int e = t + j * TPL + m * IDO;
                        C z = mk<T>(T(0), T(0));
                        if (wok && (EXACT || t + j * TPL < NB)) {
                            C A = ld_cx<T, ALIGNED>(line + (int64_t)e * sa);
                            C B = ld_cx<T, ALIGNED>(line + (int64_t)(N - e) * sa);
                            if (e == 0) { A.y = T(0); B.y = T(0); }
                            if (cj) { A.y = -A.y; B.y = -B.y; }
                            const C sm = mk<T>(A.x + B.x, A.y - B.y);
                            const C df = mk<T>(A.x - B.x, A.y + B.y);
                            const C wd = cmulc(df, __ldg(g.twA + e));  // d * conj(exp(-2 pi i e/(2N)))
                            z = cswap(mk<T>(sm.x - wd.y, sm.y + wd.x));  // s + i*wd, swapped: backward = swap.forward.swap
                        }
                        v[j * R + m] = z;
                    }
            } else if (plain && g.in_sa == (int64_t)sizeof(C)) {
                const C *p = reinterpret_cast<const C *>(line) + t;
#pragma unroll
                for (int j = 0; j < J; ++j)
#pragma unroll
                    for (int m = 0; m < R; ++m) {
                        const bool act = wok && (EXACT || t + j * TPL < NB);
                        v[j * R + m] = act ? __ldcs(p + j * TPL + m * IDO) : mk<T>(T(0), T(0));
                    }
            } else {
#pragma unroll
                for (int j = 0; j < J; ++j)
#pragma unroll
                    for (int m = 0; m < R; ++m) {
                        const int b = t + j * TPL;
                        C val = mk<T>(T(0), T(0));
                        if (wok && (EXACT || b < NB)) {
                            const uint32_t e = (uint32_t)(b + IDO * m);
                            if (plain) val = ld_cx<T, ALIGNED>(line + (int64_t)e * g.in_sa);
                            else val = load_value<T, ALIGNED>(g.load_mode, g.flags, line, g.in_sa, e, (uint32_t)N, g.n_in);
                        }
                        v[j * R + m] = val;
                    }
            }
            if (MODE == 0 && g.backward) {
#pragma unroll
                for (int q = 0; q < J * R; ++q) v[q] = cswap(v[q]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int b = t + j * TPL;
                if (EXACT || b < NB) {
                    const int i = b % IDO, k = b / IDO;
                    const C *src = sl + i + IDO * R * k;
#pragma unroll
                    for (int m = 0; m < R; ++m) v[j * R + m] = src[IDO * m];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int b = t + j * TPL;
            if (EXACT || b < NB) {
                Dft<T, R>::run(v + j * R);
                if (IDO > 1) {
                    const int i = b % IDO;
#pragma unroll
                    for (int q = 1; q < R; ++q) v[j * R + q] = cmul(v[j * R + q], __ldg(tws + i + (q - 1) * IDO));
                }
            }
        }
        if constexpr (LAST && MODE == 1) {
            // Hermitian unpack: Z = DFT_N(x[2j] + i x[2j+1]) -> X[k] = E + O, X[N-k] = conj(E - O) with
            // E = (Z[k] + conj Z[N-k])/2, O = -i w^k (Z[k] - conj Z[N-k])/2, w = exp(-2 pi i/(2N)), through one more trip
            // over shared memory (natural order)
            if (P::NPASS > 1) __syncthreads();
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int b = t + j * TPL;
                if (EXACT || b < NB) {
#pragma unroll
                    for (int q = 0; q < R; ++q) sl[b + q * NB] = v[j * R + q];
                }
            }
            __syncthreads();
            if (!wok) return;
            char *line = g.out + out_base + (int64_t)w * g.out_bs[0];
            const T half = T(0.5) * g.fct;
            const bool conj_out = g.backward != 0;
            const bool contig_out = g.out_sa == (int64_t)sizeof(C);
            for (int k = t; k <= N / 2; k += TPL) {
                const C a = sl[k];
                const C bz = sl[k ? N - k : 0];
                const C bb = mk<T>(bz.x, -bz.y);
                const C e = mk<T>((a.x + bb.x) * half, (a.y + bb.y) * half);
                const C d = mk<T>((a.x - bb.x) * half, (a.y - bb.y) * half);
                const C wd = cmul(__ldg(g.twA + k), d);
                const C o = mk<T>(wd.y, -wd.x);  // -i w^k d
                C x0 = mk<T>(e.x + o.x, e.y + o.y);
                C x1 = mk<T>(e.x - o.x, -(e.y - o.y));
                if (conj_out) { x0.y = -x0.y; x1.y = -x1.y; }
                if (contig_out) {
                    __stcs(reinterpret_cast<C *>(line) + k, x0);
                    __stcs(reinterpret_cast<C *>(line) + (N - k), x1);
                } else {
                    st_cx<T, ALIGNED>(line + (int64_t)k * g.out_sa, x0);
                    st_cx<T, ALIGNED>(line + (int64_t)(N - k) * g.out_sa, x1);
                }
            }
        } else if constexpr (LAST && MODE == 2) {
            // bins hold swap(x[2k] + i x[2k+1]): deliver the two reals (as one complex store when aligned; flags bit 0
            // set by the host when the output is not complex-aligned)
            if (!wok) return;
            char *line = g.out + out_base + (int64_t)w * g.out_bs[0];
            const T f = g.fct;
            const bool vec = g.out_sa == (int64_t)sizeof(T) && g.flags == 0;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int b = t + j * TPL;
                if (EXACT || b < NB) {
#pragma unroll
                    for (int q = 0; q < R; ++q) {
                        const int k = b + q * NB;
                        const C val = mk<T>(v[j * R + q].y * f, v[j * R + q].x * f);
                        if (vec) __stcs(reinterpret_cast<C *>(line) + k, val);
                        else {
                            *reinterpret_cast<T *>(line + (int64_t)(2 * k) * g.out_sa) = val.x;
                            *reinterpret_cast<T *>(line + (int64_t)(2 * k + 1) * g.out_sa) = val.y;
                        }
                    }
                }
            }
        } else if constexpr (LAST) {
            if (!wok) return;
            char *line = g.out + out_base + (int64_t)w * g.out_bs[0];
            const bool plain = g.store_mode == ST_C2C && g.tw_dim < 0;
            if (plain && g.out_sa == (int64_t)sizeof(C)) {
                C *p = reinterpret_cast<C *>(line) + t;
                const T f = g.fct;
                const bool bw = g.backward != 0;
#pragma unroll
                for (int j = 0; j < J; ++j)
#pragma unroll
                    for (int q = 0; q < R; ++q) {
                        if (EXACT || t + j * TPL < NB) {
                            C val = cscale(v[j * R + q], f);
                            if (bw) val = cswap(val);
                            __stcs(p + j * TPL + q * NB, val);
                        }
                    }
                return;
            }
            const uint32_t cc = (g.tw_dim == 0) ? (w_first + w) : (g.tw_dim == 1 ? i1 : i2);
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int b = t + j * TPL;
                if (EXACT || b < NB) {
#pragma unroll
                    for (int q = 0; q < R; ++q) {
                        const uint32_t f = (uint32_t)(b + q * NB);
                        C val = v[j * R + q];
                        if (g.tw_dim >= 0) {
                            uint32_t hi, lo;
                            fdivmod(cc * f, g.d_twS, hi, lo);
                            val = cmul(val, cmul(__ldg(g.twA + hi), __ldg(g.twB + lo)));
                        }
                        val = cscale(val, g.fct);
                        if (g.backward) val = cswap(val);
                        if (plain) st_cx<T, ALIGNED>(line + (int64_t)f * g.out_sa, val);
                        else store_bin_value<T, ALIGNED>(g.store_mode, g.flags, line, g.out_sa, f, (uint32_t)N, val);
                    }
                }
            }
        } else {
            if (S > 0) __syncthreads();
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int b = t + j * TPL;
                if (EXACT || b < NB) {
#pragma unroll
                    for (int q = 0; q < R; ++q) sl[b + q * NB] = v[j * R + q];
                }
            }
            __syncthreads();
            SpecPass<T, P, S + 1, ALIGNED, MODE>::run(g, buf, tid, w_first, wvalid, i1, i2, in_base, out_base);
        }
    }
};

template <typename T, typename P, bool ALIGNED, int MODE>
__global__ void __launch_bounds__(P::W *P::TPL, P::MINB) fft_spec_kernel(const TileGeom<T> g) {
    extern __shared__ __align__(16) unsigned char smem_raw_sp[];
    uint32_t t0, i1, i2, rest;
    fdivmod(blockIdx.x, g.d_t0, rest, t0);
    fdivmod(rest, g.d_e1, i2, i1);
    const uint32_t w_first = t0 * P::W;
    prefetch_later_tile<T>(g, (uint32_t)P::W);
    const int wvalid = (int)min((uint32_t)P::W, g.bext[0] - w_first);
    const int64_t in_base = (int64_t)w_first * g.in_bs[0] + (int64_t)i1 * g.in_bs[1] + (int64_t)i2 * g.in_bs[2];
    const int64_t out_base = (int64_t)w_first * g.out_bs[0] + (int64_t)i1 * g.out_bs[1] + (int64_t)i2 * g.out_bs[2];
    SpecPass<T, P, 0, ALIGNED, MODE>::run(g, reinterpret_cast<cx<T> *>(smem_raw_sp), (int)threadIdx.x, w_first, wvalid, i1, i2,
                                    in_base, out_base);
}

}  // namespace rfb
