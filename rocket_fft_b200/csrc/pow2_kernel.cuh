// Register-resident Stockham kernel for power-of-two lines, 16 <= N <= 16384.
//
// A CTA owns W lines; every thread keeps 16 points of one line in registers through all
// passes (radix-16 butterflies, preceded by one radix-2/4/8 pass when log2 N is not a
// multiple of 4).  The first pass reads global memory directly into registers, the last
// pass writes registers directly to global memory -- both fully coalesced, because in the
// autosort formulation thread t touches t + m*N/R -- and between two passes the data makes
// exactly one trip through shared memory, laid out so that the writer (contiguous) and the
// reader (stride ido*R between neighbouring butterflies) are both bank-conflict free.
// Twiddles come from a per-length table stored pass-major/q-major so that the 32 lanes of a
// warp read consecutive entries.
//
// MODE 1 fuses the even-length real transform: the n = 2N real line is read as N complex
// points, transformed, and the Hermitian unpack (one extra trip through shared memory) writes
// bins 0..N.  (Counterpart of rfftp + general_r2c in the reference,
// _pocketfft_hdronly.h:1836-2717, 3723-3779; different algorithm.)
#pragma once
#include "common.cuh"
#include "line_io.cuh"
#include "radix.cuh"
#include "tile_kernel.cuh"

namespace rfb {

// L2 cache policies for the fused four-step kernel (pow2_fused4_kernel.cuh): the scratch ring is kept in L2
// (evict_last), the array itself streams through (evict_first); neither side allocates in L1 (another CTA
// rewrites the scratch, a stale L1 line must never be hit).
__device__ __forceinline__ uint64_t l2_policy_keep() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_stream() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float2 ld_policy(const float2 *p, uint64_t pol) {
    float2 r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f32 {%0,%1}, [%2], %3;" : "=f"(r.x), "=f"(r.y) : "l"(p), "l"(pol) : "memory");
    return r;
}
__device__ __forceinline__ double2 ld_policy(const double2 *p, uint64_t pol) {
    double2 r;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(pol) : "memory");
    return r;
}
__device__ __forceinline__ void st_policy(float2 *p, float2 v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_policy(double2 *p, double2 v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}

template <int LOGN>
struct P2 {
    static constexpr int N = 1 << LOGN;
    static constexpr int REM = LOGN % 4;
    static constexpr int NP16 = LOGN / 4;
    static constexpr int RT = 1 << REM;  // small radix, executed first (1: none)
    static constexpr int NPASS = NP16 + (REM ? 1 : 0);
    static constexpr int TPL = N / 16;   // threads per line
    __host__ __device__ static constexpr int radix(int p) { return (REM && p == 0) ? RT : 16; }
    __host__ __device__ static constexpr int l1(int p) {
        int l = 1;
        for (int i = 0; i < p; ++i) l *= radix(i);
        return l;
    }
    __host__ __device__ static constexpr int ido(int p) { return N / (l1(p) * radix(p)); }
    __host__ __device__ static constexpr int twoff(int p) {
        int o = 0;
        for (int i = 0; i < p; ++i) o += (radix(i) - 1) * ido(i);
        return o;
    }
    static constexpr int PAD = N / 16;
};

// shared-memory slot of logical element a in the exchange that FEEDS pass P
template <int LOGN, int P>
__device__ __forceinline__ int p2_phys(int a) {
    constexpr int ido = P2<LOGN>::ido(P), R = P2<LOGN>::radix(P);
    if constexpr (ido < 16) return a + ido * (a / (ido * R));
    else return a;
}

template <typename T, int LOGN, int W, int MODE>
struct Pow2Body {
    using C = cx<T>;
    using PL = P2<LOGN>;
    static constexpr int N = PL::N, TPL = PL::TPL, NT = W * TPL;
    static constexpr int PITCH = (W == 1) ? (N + PL::PAD) : ((N + PL::PAD) | 1);

    static __device__ __forceinline__ void map(bool lf, int tid, int &w, int &t) {
        if (lf) { w = tid % W; t = tid / W; }
        else { w = tid / TPL; t = tid % TPL; }
    }

    // butterflies + twiddles of pass P on the 16 registers
    template <int P>
    static __device__ __forceinline__ void compute(C *v, int t, const C *__restrict__ stw) {
        constexpr int R = PL::radix(P), NB = 16 / R, ido = PL::ido(P);
#pragma unroll
        for (int j = 0; j < NB; ++j) Dft<T, R>::run(v + j * R);
        if constexpr (ido > 1) {
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                const int i = (t + j * TPL) % ido;
                const C *tw = stw + PL::twoff(P) + i;
                if constexpr (R == 16 && sizeof(T) == 8 && ido >= 64) {
                    // fp64: the table of a long pass (15*ido entries) does not stay in L1 next to the data;
                    // load w^1..w^4 and build the other powers (error a few ulp << the 1e-13 budget)
                    C w[16];
#pragma unroll
                    for (int q = 1; q <= 4; ++q) w[q] = __ldg(tw + (q - 1) * ido);
                    w[5] = cmul(w[4], w[1]); w[6] = cmul(w[4], w[2]); w[7] = cmul(w[4], w[3]); w[8] = cmul(w[4], w[4]);
#pragma unroll
                    for (int q = 1; q <= 8; ++q) v[j * R + q] = cmul(v[j * R + q], w[q]);
#pragma unroll
                    for (int q = 9; q < 16; ++q) v[j * R + q] = cmul(v[j * R + q], cmul(w[8], w[q - 8]));
                } else {
#pragma unroll
                    for (int q = 1; q < R; ++q) v[j * R + q] = cmul(v[j * R + q], __ldg(tw + (q - 1) * ido));
                }
            }
        }
    }

    // registers (outputs of pass P-1) -> shared -> registers (inputs of pass P)
    template <int P>
    static __device__ __forceinline__ void exchange(C *v, C *buf, int tid, bool lf_prev, bool lf_next, bool first) {
        constexpr int Rp = PL::radix(P - 1), NBp = 16 / Rp;
        constexpr int ido = PL::ido(P);
        int w, t;
        map(lf_prev, tid, w, t);
        if (!first) __syncthreads();
        C *line = buf + w * PITCH;
#pragma unroll
        for (int j = 0; j < NBp; ++j)
#pragma unroll
            for (int q = 0; q < Rp; ++q) line[p2_phys<LOGN, P>(t + j * TPL + q * (N / Rp))] = v[j * Rp + q];
        __syncthreads();
        map(lf_next, tid, w, t);
        line = buf + w * PITCH;
        if (P == PL::NPASS - 1 && SHFL_POST) t = pair_perm(t);
        const int i = t % ido, k = t / ido;
#pragma unroll
        for (int m = 0; m < 16; ++m) v[m] = line[p2_phys<LOGN, P>(i + ido * (m + 16 * k))];
    }

    // MODE 1 with >= 32 threads per line: in the last pass lane l of a warp takes butterfly b and lane
    // 31-l takes butterfly TPL-b, so that the Hermitian partner Z[N-k] of every bin a lane holds sits in
    // the mirrored lane's registers (k = b + TPL q  <->  N-k = (TPL-b) + TPL (15-q)) and the unpack needs
    // warp shuffles instead of another trip through shared memory.  b = 0 and b = TPL/2 pair with themselves.
    // Measured on B200 (round 1): 0.575 ms vs 0.550 ms for the shared-memory unpack on 16384 x 16384 rows --
    // 16 SHFL + divergent selects + two 128-byte store runs per warp cost more than 33 LDS/STS + 2 barriers --
    // so the variant is compiled only with -DRFB_SHFL_UNPACK=1.
#ifndef RFB_SHFL_UNPACK
#define RFB_SHFL_UNPACK 0
#endif
    static constexpr bool SHFL_POST = (RFB_SHFL_UNPACK != 0) && (MODE == 1) && (TPL >= 32);
    static __device__ __forceinline__ int pair_perm(int t) {
        const int wl = t >> 5, l = t & 31;
        if (l < 16) return 16 * wl + l;
        const int b = TPL - 16 * wl - (31 - l);
        return b == TPL ? TPL / 2 : b;
    }

    static __device__ __forceinline__ void run(const TileGeom<T> &g, const C *__restrict__ stw, C *buf) {
        run_tile<0>(g, stw, buf, blockIdx.x, 0, 0, g.bext[0]);
    }

    // POL 0: one tile per CTA (tile = blockIdx.x).  POL 1 / 2: steps A / B of the fused four-step kernel -- the
    // tile index and the byte offsets of the current strip come from the caller; strided plain loads and strided
    // complex stores carry L2 policies (1: array -> scratch ring, 2: scratch ring -> array); ext0 = lines of the
    // tile dim that exist in this strip (the caller skips tiles that lie entirely beyond it).
    template <int POL>
    static __device__ __forceinline__ void run_tile(const TileGeom<T> &g, const C *__restrict__ stw, C *buf, uint32_t tile,
                                                    int64_t in_off, int64_t out_off, uint32_t ext0) {
        uint32_t t0, i1, i2, rest;
        fdivmod(tile, g.d_t0, rest, t0);
        fdivmod(rest, g.d_e1, i2, i1);
        const uint32_t w_first = t0 * W;
        const int wvalid = (int)min((uint32_t)W, ext0 - w_first);
        const int64_t in_base = in_off + (int64_t)w_first * g.in_bs[0] + (int64_t)i1 * g.in_bs[1] + (int64_t)i2 * g.in_bs[2];
        const int64_t out_base = out_off + (int64_t)w_first * g.out_bs[0] + (int64_t)i1 * g.out_bs[1] + (int64_t)i2 * g.out_bs[2];
        const int tid = threadIdx.x;
        const bool lf_in = g.load_line_fast != 0, lf_out = g.store_line_fast != 0;
        C v[16];
        bool staged_in = false;
        uint64_t pol_in = 0, pol_out = 0;
        if constexpr (POL == 1) { pol_in = l2_policy_stream(); pol_out = l2_policy_keep(); }
        if constexpr (POL == 2) { pol_in = l2_policy_keep(); pol_out = l2_policy_stream(); }
        if constexpr (POL == 0) prefetch_later_tile<T>(g, (uint32_t)W);

        // ---- pass 0: global -> registers -------------------------------------------------
        {
            constexpr int R = PL::radix(0), NB = 16 / R, ido = PL::ido(0);
            int w, t;
            map(lf_in, tid, w, t);
            const bool wok = (W == 1) ? true : (w < wvalid);  // W == 1: the grid has exactly one CTA per line
            const char *line = g.in + in_base + (int64_t)w * g.in_bs[0];
            // All 16 loads are issued back to back with nothing depending on them in between
            // (memory-level parallelism: 16 independent requests per thread in flight).
            // MODE 1: n_in counts REAL samples present (2N when the line is not zero-padded)
            const bool packed_vec = (MODE == 1) && g.in_sa == (int64_t)sizeof(T) && g.n_in == 2u * (uint32_t)N;
            const bool plain = (MODE == 0 || MODE == 5) && g.load_mode == LD_C2C && g.n_in == (uint32_t)N;
            if (MODE == 2) {
                // Packed inverse real transform: the N+1 Hermitian bins X are folded into the N-point
                // complex spectrum  Z[e] = (X[e] + conj X[N-e]) + i w^e (X[e] - conj X[N-e]),
                // w = exp(+2 pi i/(2N)), whose backward DFT is x[2m] + i x[2m+1].  Im X[0], Im X[N]
                // are ignored (reference: general_c2r, H:3830, 3845-3846).  forward=true: X -> conj X.
                const bool cj = g.backward == 0;
                const int64_t sa = g.in_sa;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    C a[8], b[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int idx = h * 8 + i;
                        const int e = t + (idx / R) * TPL + (idx % R) * ido;
                        a[i] = wok ? *reinterpret_cast<const C *>(line + (int64_t)e * sa) : mk<T>(T(0), T(0));
                        b[i] = wok ? *reinterpret_cast<const C *>(line + (int64_t)(N - e) * sa) : mk<T>(T(0), T(0));
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int idx = h * 8 + i;
                        const int e = t + (idx / R) * TPL + (idx % R) * ido;
                        C A = a[i], B = b[i];
                        if (e == 0) { A.y = T(0); B.y = T(0); }
                        if (cj) { A.y = -A.y; B.y = -B.y; }
                        const C s = mk<T>(A.x + B.x, A.y - B.y);
                        const C d = mk<T>(A.x - B.x, A.y + B.y);
                        const C wc = __ldg(g.twA + e);           // exp(-2 pi i e/(2N))
                        const C wd = cmulc(d, wc);                // d * conj(wc)
                        const C z = mk<T>(s.x - wd.y, s.y + wd.x);  // s + i*wd
                        v[idx] = cswap(z);
                    }
                }
            } else if (MODE == 3) {
                // DCT-II / DST-II of a real line of length 2N (N = complex points here): Makhoul's
                // reordering v = [x0, x2, x4, ..., x5, x3, x1] packed two-by-two into complex points;
                // the sine transform alternates the sign of the odd samples.
                const bool sine = (g.flags & FLAG_SINE) != 0;
                const int64_t sa = g.in_sa;
#pragma unroll
                for (int j = 0; j < NB; ++j)
#pragma unroll
                    for (int m = 0; m < R; ++m) {
                        const int e = t + j * TPL + m * ido;
                        const bool lo = e < N / 2;
                        const int i0 = lo ? 4 * e : 4 * N - 1 - 4 * e;
                        const int i1 = lo ? 4 * e + 2 : 4 * N - 3 - 4 * e;
                        C val = mk<T>(T(0), T(0));
                        if (wok) {
                            val.x = *reinterpret_cast<const T *>(line + (int64_t)i0 * sa);
                            val.y = *reinterpret_cast<const T *>(line + (int64_t)i1 * sa);
                        }
                        if (sine && !lo) { val.x = -val.x; val.y = -val.y; }
                        v[j * R + m] = val;
                    }
            } else if (MODE == 4) {
                // DCT-III / DST-III: the Hermitian spectrum H_k = conj(c_k) (x_k - i x_{2N-k}),
                // c_k = exp(-i pi k / (4N)), of the reordered output is built on the fly and folded into
                // the packed inverse real transform exactly as in MODE 2.
                const bool sine = (g.flags & FLAG_SINE) != 0, ortho = (g.flags & FLAG_ORTHO) != 0;
                const int NR = 2 * N;  // real length
                const int scaled = (sine && (g.flags & FLAG_QUIRK) == 0) ? NR - 1 : 0;  // caller's index scaled by sqrt2
                const int64_t sa = g.in_sa;
                auto ld = [&](int k) -> T {  // transform's element k (k == NR reads as 0)
                    if (k >= NR || !wok) return T(0);
                    const int io = sine ? NR - 1 - k : k;
                    T r = *reinterpret_cast<const T *>(line + (int64_t)io * sa);
                    if (ortho && io == scaled) r *= T(1.4142135623730951);
                    return r;
                };
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    T r[4][4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int idx = h * 4 + i;
                        const int e = t + (idx / R) * TPL + (idx % R) * ido;
                        r[i][0] = ld(e);
                        r[i][1] = ld(NR - e);
                        r[i][2] = ld(N - e);
                        r[i][3] = ld(N + e);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int idx = h * 4 + i;
                        const int e = t + (idx / R) * TPL + (idx % R) * ido;
                        const C ce = __ldg(g.twB + e), cm = __ldg(g.twB + (N - e));
                        C A = cmulc(mk<T>(r[i][0], -r[i][1]), ce);
                        C B = cmulc(mk<T>(r[i][2], -r[i][3]), cm);
                        if (e == 0) { A.y = T(0); B.y = T(0); }
                        const C s = mk<T>(A.x + B.x, A.y - B.y);
                        const C d = mk<T>(A.x - B.x, A.y + B.y);
                        const C wd = cmulc(d, __ldg(g.twA + e));
                        v[idx] = cswap(mk<T>(s.x - wd.y, s.y + wd.x));
                    }
                }
            } else if (POL == 0 && MODE == 0 && g.pre_tab != nullptr) {
                // zero-padded load with a fused element-wise factor (Bluestein's chirp): global index
                // gi = e * g_mul + c * c_mul decides both the padding and the table entry
                const uint32_t c = (g.c_dim < 0 ? 0u : ((g.c_dim == 0) ? (w_first + w) : (g.c_dim == 1 ? i1 : i2))) * g.c_mul;
                const int64_t sa = g.in_sa;
#pragma unroll
                for (int j = 0; j < NB; ++j)
#pragma unroll
                    for (int m = 0; m < R; ++m) {
                        const uint32_t e = (uint32_t)(t + j * TPL + m * ido);
                        const uint32_t gi = e * g.g_mul + c;
                        const bool ok = wok && gi < g.pre_bound;
                        v[j * R + m] = ok ? *reinterpret_cast<const C *>(line + (int64_t)e * sa) : mk<T>(T(0), T(0));
                    }
#pragma unroll
                for (int j = 0; j < NB; ++j)
#pragma unroll
                    for (int m = 0; m < R; ++m) {
                        const uint32_t e = (uint32_t)(t + j * TPL + m * ido);
                        const uint32_t gi = e * g.g_mul + c;
                        if (gi < g.pre_bound) {
                            C val = v[j * R + m];
                            if (g.pre_swap) val = cswap(val);
                            v[j * R + m] = cmul(val, __ldg(g.pre_tab + gi));
                        }
                    }
            } else if (POL == 0 && MODE == 0 && LOGN <= 6 && (g.stage_io & 1)) {
                // Short lines, many per CTA (one to four threads per line): read straight into registers, every load
                // instruction of a warp would touch 32 different 128-byte lines.  The tile is one dense run of
                // W*N points, so it is copied to shared memory with fully coalesced loads and picked up from there.
                const C *tile = reinterpret_cast<const C *>(g.in + in_base);
                const int cnt = wvalid * N;
                for (int idx = tid; idx < cnt; idx += NT) buf[(idx >> LOGN) * PITCH + (idx & (N - 1))] = __ldcs(tile + idx);
                __syncthreads();
                staged_in = true;
                const C *sl = buf + w * PITCH + t;
#pragma unroll
                for (int j = 0; j < NB; ++j)
#pragma unroll
                    for (int m = 0; m < R; ++m) v[j * R + m] = wok ? sl[j * TPL + m * ido] : mk<T>(T(0), T(0));
            } else if (POL == 0 && (packed_vec || (plain && g.in_sa == (int64_t)sizeof(C)))) {
                // contiguous line: one base pointer, compile-time offsets
                const C *p = reinterpret_cast<const C *>(line) + t;
#pragma unroll
                for (int j = 0; j < NB; ++j)
#pragma unroll
                    for (int m = 0; m < R; ++m) v[j * R + m] = wok ? __ldcs(p + j * TPL + m * ido) : mk<T>(T(0), T(0));
            } else if (POL != 0 || plain) {  // (the fused four-step kernel only ever takes this path)
                // strided line: running pointers instead of a 64-bit multiply per element
                const int64_t sa = g.in_sa;
                const int64_t step_m = (int64_t)ido * sa, step_j = (int64_t)TPL * sa;
                const char *pj = line + (int64_t)t * sa;
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    const char *pm = pj;
#pragma unroll
                    for (int m = 0; m < R; ++m) {
                        if constexpr (POL != 0) v[j * R + m] = wok ? ld_policy(reinterpret_cast<const C *>(pm), pol_in) : mk<T>(T(0), T(0));
                        else v[j * R + m] = wok ? *reinterpret_cast<const C *>(pm) : mk<T>(T(0), T(0));
                        pm += step_m;
                    }
                    pj += step_j;
                }
            } else {
#pragma unroll
                for (int j = 0; j < NB; ++j)
#pragma unroll
                    for (int m = 0; m < R; ++m) {
                        const uint32_t e = (uint32_t)(t + j * TPL + m * ido);
                        C val = mk<T>(T(0), T(0));
                        if (wok) {
                            if (MODE == 1) {
                                if (2 * e < g.n_in) val.x = *reinterpret_cast<const T *>(line + (int64_t)(2 * e) * g.in_sa);
                                if (2 * e + 1 < g.n_in) val.y = *reinterpret_cast<const T *>(line + (int64_t)(2 * e + 1) * g.in_sa);
                            } else val = load_value<T, true>(g.load_mode, g.flags, line, g.in_sa, e, (uint32_t)N, g.n_in);
                        }
                        v[j * R + m] = val;
                    }
            }
            if (MODE == 0 && g.backward) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = cswap(v[i]);
            }
            compute<0>(v, t, stw);
        }
        bool first = !staged_in;  // false: shared memory is in use, the next writer synchronises first
        if constexpr (PL::NPASS > 1) {
            exchange<1>(v, buf, tid, lf_in, lf_out, first);
            first = false;
            int w, t;
            map(lf_out, tid, w, t);
            compute<1>(v, t, stw);
        }
        if constexpr (PL::NPASS > 2) {
            exchange<2>(v, buf, tid, lf_out, lf_out, first);
            int w, t;
            map(lf_out, tid, w, t);
            compute<2>(v, t, stw);
        }
        if constexpr (PL::NPASS > 3) {
            exchange<3>(v, buf, tid, lf_out, lf_out, first);
            int w, t;
            map(lf_out, tid, w, t);
            compute<3>(v, t, stw);
        }

        // ---- after the last pass thread t holds bins t + q*N/R --------------------------------
        constexpr int RL = PL::radix(PL::NPASS - 1), NBL = 16 / RL;
        int w, t;
        map(PL::NPASS > 1 ? lf_out : lf_in, tid, w, t);
        if constexpr (MODE == 5) {
            // Circular-convolution row (Bluestein's middle): the spectrum just computed is multiplied by
            // pre_tab[bin * g_mul + c * c_mul], and transformed back without leaving the SM.  One trip through shared
            // memory turns the output distribution of the last pass (bins t + q N/RL) into the input distribution
            // of pass 0 (elements t + m ido).  The backward transform is swap . forward . swap; the final swap
            // is the store path's.
            const uint32_t cc = (g.c_dim < 0 ? 0u : ((g.c_dim == 0) ? (w_first + w) : (g.c_dim == 1 ? i1 : i2))) * g.c_mul;
            const bool wok5 = (W == 1) ? true : (w < wvalid);
#pragma unroll
            for (int j = 0; j < NBL; ++j)
#pragma unroll
                for (int q = 0; q < RL; ++q) {
                    const uint32_t k = (uint32_t)(t + j * TPL + q * (N / RL));
                    C val = v[j * RL + q];
                    if (wok5) val = cmul(val, __ldg(g.pre_tab + (k * g.g_mul + cc)));
                    v[j * RL + q] = cswap(val);
                }
            if (!first) __syncthreads();
            C *sl = buf + w * PITCH;
#pragma unroll
            for (int j = 0; j < NBL; ++j)
#pragma unroll
                for (int q = 0; q < RL; ++q) sl[t + j * TPL + q * (N / RL)] = v[j * RL + q];
            __syncthreads();
            {
                constexpr int R0 = PL::radix(0), NB0 = 16 / R0, ido0 = PL::ido(0);
#pragma unroll
                for (int j = 0; j < NB0; ++j)
#pragma unroll
                    for (int m = 0; m < R0; ++m) v[j * R0 + m] = sl[t + j * TPL + m * ido0];
            }
            compute<0>(v, t, stw);
            if constexpr (PL::NPASS > 1) { exchange<1>(v, buf, tid, false, false, false); compute<1>(v, t, stw); }
            if constexpr (PL::NPASS > 2) { exchange<2>(v, buf, tid, false, false, false); compute<2>(v, t, stw); }
            if constexpr (PL::NPASS > 3) { exchange<3>(v, buf, tid, false, false, false); compute<3>(v, t, stw); }
        }
        if (MODE == 2 || MODE == 4) {
            // bins hold swap(x[2k] + i x[2k+1]); deliver the two reals
            if (w >= wvalid) return;
            char *line = g.out + out_base + (int64_t)w * g.out_bs[0];
            const T f = g.fct;
            const bool vec = MODE == 2 && g.out_sa == (int64_t)sizeof(T) && g.flags == 0;  // flags bit0: output not complex-aligned
            const bool sine = MODE == 4 && (g.flags & FLAG_SINE) != 0;
#pragma unroll
            for (int j = 0; j < NBL; ++j)
#pragma unroll
                for (int q = 0; q < RL; ++q) {
                    const int k = t + j * TPL + q * (N / RL);
                    C val = mk<T>(v[j * RL + q].y * f, v[j * RL + q].x * f);
                    if (MODE == 2 && vec) { __stcs(reinterpret_cast<C *>(line + (int64_t)k * 2 * sizeof(T)), val); continue; }
                    if (MODE == 4) {
                        // undo Makhoul's reordering; the sine transform carries (-1)^j
                        const bool lo = k < N / 2;
                        const int j0 = lo ? 4 * k : 4 * N - 1 - 4 * k;
                        const int j1 = lo ? 4 * k + 2 : 4 * N - 3 - 4 * k;
                        if (sine && !lo) { val.x = -val.x; val.y = -val.y; }
                        *reinterpret_cast<T *>(line + (int64_t)j0 * g.out_sa) = val.x;
                        *reinterpret_cast<T *>(line + (int64_t)j1 * g.out_sa) = val.y;
                    } else if (vec) *reinterpret_cast<C *>(line + (int64_t)k * 2 * sizeof(T)) = val;
                    else {
                        *reinterpret_cast<T *>(line + (int64_t)(2 * k) * g.out_sa) = val.x;
                        *reinterpret_cast<T *>(line + (int64_t)(2 * k + 1) * g.out_sa) = val.y;
                    }
                }
            return;
        }
        if (MODE == 0 || MODE == 5) {
            if (POL == 0 && MODE == 0 && LOGN <= 6 && (g.stage_io & 2)) {
                // short lines: through shared memory, then one dense, fully coalesced run of stores (see the load)
                const T f = g.fct;
                const bool bw = g.backward != 0;
                if (!first) __syncthreads();
                if (w < wvalid) {
                    C *sl = buf + w * PITCH + t;
#pragma unroll
                    for (int j = 0; j < NBL; ++j)
#pragma unroll
                        for (int q = 0; q < RL; ++q) {
                            C val = cscale(v[j * RL + q], f);
                            if (bw) val = cswap(val);
                            sl[j * TPL + q * (N / RL)] = val;
                        }
                }
                __syncthreads();
                C *tile = reinterpret_cast<C *>(g.out + out_base);
                const int cnt = wvalid * N;
                for (int idx = tid; idx < cnt; idx += NT) __stcs(tile + idx, buf[(idx >> LOGN) * PITCH + (idx & (N - 1))]);
                return;
            }
            if (w >= wvalid) return;
            char *line = g.out + out_base + (int64_t)w * g.out_bs[0];
            if (POL == 0 && g.store_mode == ST_C2C && g.tw_dim < 0 && g.out_sa == (int64_t)sizeof(C) && g.split_blk == 0 &&
                g.post_tab == nullptr) {
                const T f = g.fct;
                const bool bw = (MODE == 5) || g.backward != 0;
                C *p = reinterpret_cast<C *>(line) + t;
#pragma unroll
                for (int j = 0; j < NBL; ++j)
#pragma unroll
                    for (int q = 0; q < RL; ++q) {
                        C val = cscale(v[j * RL + q], f);
                        if (bw) val = cswap(val);
                        __stcs(p + j * TPL + q * (N / RL), val);
                    }
                return;
            }
            if (POL != 0 || g.store_mode == ST_C2C) {
                // strided output, optionally with the four-step factor exp(-2 pi i c k / bigN):
                // exact two-level table look-ups for every 4th bin, three recurrence steps between
                const T f = g.fct;
                const bool bw = (MODE == 5) || g.backward != 0;
                const bool tw = (POL == 2) ? false : g.tw_dim >= 0;  // fused four-step: step A has the factor, step B not
                const uint32_t c = tw ? ((g.tw_dim == 0) ? (w_first + w) : (g.tw_dim == 1 ? i1 : i2)) : 0u;
                auto lookup = [&](uint32_t x) {
                    uint32_t hi, lo;
                    fdivmod(x, g.d_twS, hi, lo);
                    return cmul(__ldg(g.twA + hi), __ldg(g.twB + lo));
                };
                C step = mk<T>(T(1), T(0));
                if (tw) step = lookup(c * (uint32_t)(N / RL));
                const int64_t step_q = (int64_t)(N / RL) * g.out_sa, step_j = (int64_t)TPL * g.out_sa;
                char *pj = line + (int64_t)t * g.out_sa;
#pragma unroll
                for (int j = 0; j < NBL; ++j) {
                    char *pq = pj;
                    C wq = mk<T>(T(1), T(0));
#pragma unroll
                    for (int q = 0; q < RL; ++q) {
                        C val = v[j * RL + q];
                        if (tw) {
                            if ((q & 3) == 0) wq = lookup(c * (uint32_t)(t + j * TPL + q * (N / RL)));
                            else wq = cmul(wq, step);
                            val = cmul(val, wq);
                        }
                        val = cscale(val, f);
                        if (bw) val = cswap(val);
                        if (POL == 0 && g.post_tab != nullptr) {
                            // fused element-wise factor / truncation on the way out (Bluestein)
                            const uint32_t cc = (g.c_dim < 0 ? 0u : ((g.c_dim == 0) ? (w_first + w) : (g.c_dim == 1 ? i1 : i2))) * g.c_mul;
                            const uint32_t bin = (uint32_t)(t + j * TPL + q * (N / RL)) * g.g_mul + cc;
                            if (bin >= g.post_bound) { pq += step_q; continue; }
                            val = cmul(val, __ldg(g.post_tab + bin));
                            if (g.post_swap) val = cswap(val);
                        }
                        if (POL == 0 && g.split_blk) {
                            // fused exchange: the bin's block decides which (peer) buffer receives it
                            const uint32_t k = (uint32_t)(t + j * TPL + q * (N / RL));
                            char *dst = g.split_base[fdiv(k, g.d_split)] + (pq - g.out);
                            *reinterpret_cast<C *>(dst) = val;
                        } else if constexpr (POL != 0) st_policy(reinterpret_cast<C *>(pq), val, pol_out);
                        else *reinterpret_cast<C *>(pq) = val;
                        pq += step_q;
                    }
                    pj += step_j;
                }
                return;
            }
#pragma unroll
            for (int j = 0; j < NBL; ++j)
#pragma unroll
                for (int q = 0; q < RL; ++q) {
                    const uint32_t k = (uint32_t)(t + j * TPL + q * (N / RL));
                    C val = v[j * RL + q];
                    if (g.store_mode == ST_HALF && 2 * k > (uint32_t)N) continue;
                    if (g.tw_dim >= 0) {
                        const uint32_t c = (g.tw_dim == 0) ? (w_first + w) : (g.tw_dim == 1 ? i1 : i2);
                        uint32_t hi, lo;
                        fdivmod(c * k, g.d_twS, hi, lo);
                        val = cmul(val, cmul(__ldg(g.twA + hi), __ldg(g.twB + lo)));
                    }
                    val = cscale(val, g.fct);
                    if (g.backward) val = cswap(val);
                    store_value<T, true>(g.store_mode, g.flags, line, g.out_sa, k, val);
                }
        } else {
            // ---- Hermitian unpack of the packed real transform --------------------------------
            // Z = DFT_N(x[2j] + i x[2j+1]);  X[k] = E + O, X[N-k] = conj(E - O) with
            // E = (Z[k] + conj Z[N-k])/2, O = -i w^k (Z[k] - conj Z[N-k])/2, w = exp(-2 pi i/(2N))
            if constexpr (SHFL_POST) {
                if (w >= wvalid) return;
                const int b = pair_perm(t);  // this lane's butterfly: bins b + TPL*q
                char *line = g.out + out_base + (int64_t)w * g.out_bs[0];
                const T half = T(0.5) * g.fct;
                const bool conj_out = g.backward != 0;
                const bool contig_out = g.out_sa == (int64_t)sizeof(C);
                auto emit2 = [&](int k, C a, C bz) {
                    const C bb = mk<T>(bz.x, -bz.y);
                    const C e = mk<T>((a.x + bb.x) * half, (a.y + bb.y) * half);
                    const C d = mk<T>((a.x - bb.x) * half, (a.y - bb.y) * half);
                    const C wd = cmul(__ldg(g.twA + k), d);
                    const C o = mk<T>(wd.y, -wd.x);
                    C x0 = mk<T>(e.x + o.x, e.y + o.y);
                    C x1 = mk<T>(e.x - o.x, -(e.y - o.y));
                    if (conj_out) { x0.y = -x0.y; x1.y = -x1.y; }
                    if (contig_out) {
                        __stcs(reinterpret_cast<C *>(line) + k, x0);
                        __stcs(reinterpret_cast<C *>(line) + (N - k), x1);
                    } else {
                        st_cx<T, true>(line + (int64_t)k * g.out_sa, x0);
                        st_cx<T, true>(line + (int64_t)(N - k) * g.out_sa, x1);
                    }
                };
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    C pz;  // Z[N - (b + TPL q)]: the mirrored lane's register 15-q
                    pz.x = __shfl_xor_sync(0xffffffffu, v[15 - q].x, 31);
                    pz.y = __shfl_xor_sync(0xffffffffu, v[15 - q].y, 31);
                    if (b == 0) pz = v[(16 - q) & 15];
                    else if (b == TPL / 2) pz = v[15 - q];
                    emit2(b + TPL * q, v[q], pz);
                }
                if (b == 0) emit2(N / 2, v[8], v[8]);
                return;
            }
            if (!first) __syncthreads();
            C *sl = buf + w * PITCH;
#pragma unroll
            for (int j = 0; j < NBL; ++j)
#pragma unroll
                for (int q = 0; q < RL; ++q) sl[t + j * TPL + q * (N / RL)] = v[j * RL + q];
            __syncthreads();
            if (w >= wvalid) return;
            char *line = g.out + out_base + (int64_t)w * g.out_bs[0];
            const T half = T(0.5) * g.fct;
            const bool conj_out = g.backward != 0;
            const bool contig_out = g.out_sa == (int64_t)sizeof(C);
            auto emit = [&](int k) {
                const C a = sl[k];
                const C bz = sl[(N - k) & (N - 1)];
                const C b = mk<T>(bz.x, -bz.y);
                const T hh = (MODE == 3) ? g.fct : half;  // DCT: the factor 2 of y = 2 Re(c V) cancels the 1/2
                const C e = mk<T>((a.x + b.x) * hh, (a.y + b.y) * hh);
                const C d = mk<T>((a.x - b.x) * hh, (a.y - b.y) * hh);
                const C wk = __ldg(g.twA + k);
                // o = -i * wk * d
                const C wd = cmul(wk, d);
                const C o = mk<T>(wd.y, -wd.x);
                C x0 = mk<T>(e.x + o.x, e.y + o.y);
                C x1 = mk<T>(e.x - o.x, -(e.y - o.y));
                if (MODE == 3) {
                    // y[k] = Re(c_k V_k), y[2N-k] = -Im(c_k V_k) for V_k = x0 and V_{N-k} = x1
                    const bool sine = (g.flags & FLAG_SINE) != 0, ortho = (g.flags & FLAG_ORTHO) != 0;
                    const int NR = 2 * N;
                    const int scaled = (sine && (g.flags & FLAG_QUIRK) != 0) ? NR - 1 : 0;  // cosine-order index scaled by 1/sqrt2
                    auto put = [&](int idx, T val) {
                        if (ortho && idx == scaled) val *= T(0.70710678118654752);
                        const int pos = sine ? NR - 1 - idx : idx;
                        *reinterpret_cast<T *>(line + (int64_t)pos * g.out_sa) = val;
                    };
                    const C p0 = cmul(__ldg(g.twB + k), x0);
                    put(k, p0.x);
                    if (k > 0) put(NR - k, -p0.y);
                    const C p1 = cmul(__ldg(g.twB + (N - k)), x1);
                    put(N - k, p1.x);
                    put(N + k, -p1.y);
                    return;
                }
                if (conj_out) { x0.y = -x0.y; x1.y = -x1.y; }
                if (contig_out) {
                    __stcs(reinterpret_cast<C *>(line) + k, x0);
                    __stcs(reinterpret_cast<C *>(line) + (N - k), x1);
                } else {
                    st_cx<T, true>(line + (int64_t)k * g.out_sa, x0);
                    st_cx<T, true>(line + (int64_t)(N - k) * g.out_sa, x1);
                }
            };
#pragma unroll
            for (int j = 0; j < 8; ++j) emit(t + j * TPL);
            if (t == 0) emit(N / 2);
        }
    }
};

// occupancy target: 1024 threads/SM for float (<= 64 registers), 512 for double (<= 128)
template <typename T, int NT>
constexpr int p2_min_blocks() {
    return (sizeof(T) == 8 ? 512 : 1024) / NT > 0 ? (sizeof(T) == 8 ? 512 : 1024) / NT : 1;
}

template <typename T, int LOGN, int W, int MODE>
__global__ void __launch_bounds__(W *(1 << LOGN) / 16, p2_min_blocks<T, W *(1 << LOGN) / 16>()) fft_pow2_kernel(const TileGeom<T> g, const cx<T> *__restrict__ stw) {
    extern __shared__ __align__(16) unsigned char smem_raw_p2[];
    Pow2Body<T, LOGN, W, MODE>::run(g, stw, reinterpret_cast<cx<T> *>(smem_raw_p2));
}

}  // namespace rfb
