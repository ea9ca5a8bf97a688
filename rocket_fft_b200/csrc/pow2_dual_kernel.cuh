// Packed real transform of a LONG power-of-two line (single precision) as two interleaved half-length complex
// transforms per thread.
//
// The register kernel of pow2_kernel.cuh handles the real line of 2N samples as ONE N-point complex transform
// (z[e] = x[2e] + i x[2e+1]) followed by the Hermitian unpack.  For N = 8192 that is radix 2 . 16 . 16 . 16: four
// passes, three trips through shared memory plus a fourth for the unpack, and ncu shows the kernel limited by the
// L1/shared-memory data path (l1tex 78 % busy, MIO-throttle stalls), not by HBM (profiles/r01_ncu_rfft2_kernels_v2.txt).
//
// Here the radix-2 stage is taken out of the pass sequence (decimation in time for r2c, in frequency for c2r):
//   r2c:  A = DFT_{N/2}(z[0::2]),  B = DFT_{N/2}(z[1::2]),  Z[k] = A[k] + w^k B[k],  Z[k + N/2] = A[k] - w^k B[k]
//         (w = exp(-2 pi i/N)); the combination is folded into the Hermitian unpack, which needs A, B at k and N/2-k only.
//   c2r:  the Hermitian fold produces Z[e] and Z[e + N/2] in the same thread; P = Z[e] + Z[e+N/2] and
//         Q = (Z[e] - Z[e+N/2]) conj(w)^e are transformed separately and give y[2m], y[2m+1]: one 16-byte store.
// One thread owns element e of BOTH half-length transforms (16 + 16 complex points in registers): one 16-byte load
// brings z[2e], z[2e+1]; the two transforms share every twiddle load; the exchanges move 16-byte {A, B} pairs
// (LDS.128 / STS.128: half the shared-memory instructions); N/2 = 4096 = 16^3 needs two exchanges.  Per line:
// 3 (r2c) / 2 (c2r) trips through shared memory instead of 4 / 3, ~56 instead of ~83 instructions per point.
// (Counterpart of rfftp + general_r2c / general_c2r in the reference, _pocketfft_hdronly.h:1836-2717, 3723-3852;
// different algorithm.)
#pragma once
#include "pow2_kernel.cuh"

namespace rfb {

template <bool B> struct BoolC { static constexpr bool value = B; };

// v * exp(-2 pi i m/64) for 0 <= m < 32; m is a compile-time constant once the caller's loop is unrolled
__device__ __forceinline__ float2 mul_root64(float2 v, int m) {
    if (m >= 16) { v = mk<float>(v.y, -v.x); m -= 16; }  // a quarter turn first
    if (m == 0) return v;
    float c = 1.f, s = 0.f;
    switch (m) {
        case 1: c = 0.995184727f; s = 0.0980171403f; break;
        case 2: c = 0.98078528f; s = 0.195090322f; break;
        case 3: c = 0.956940336f; s = 0.290284677f; break;
        case 4: c = 0.923879533f; s = 0.382683432f; break;
        case 5: c = 0.881921264f; s = 0.471396737f; break;
        case 6: c = 0.831469612f; s = 0.555570233f; break;
        case 7: c = 0.773010453f; s = 0.634393284f; break;
        case 8: c = 0.707106781f; s = 0.707106781f; break;
        case 9: c = 0.634393284f; s = 0.773010453f; break;
        case 10: c = 0.555570233f; s = 0.831469612f; break;
        case 11: c = 0.471396737f; s = 0.881921264f; break;
        case 12: c = 0.382683432f; s = 0.923879533f; break;
        case 13: c = 0.290284677f; s = 0.956940336f; break;
        case 14: c = 0.195090322f; s = 0.98078528f; break;
        case 15: c = 0.0980171403f; s = 0.995184727f; break;
        default: break;
    }
    return mk<float>(v.x * c + v.y * s, v.y * c - v.x * s);
}

// TWC: twiddles of long passes and of the pre/post stages are composed from a few table entries and compile-time
// roots of unity instead of being loaded one by one (fewer L1 wavefronts, a few more multiplies)
template <int LOGNH, int MODE, bool TWC = false>  // MODE 0: c2c (2*2^LOGNH points), 1: r2c, 2: c2r (as in pow2_kernel.cuh)
struct DualBody {
    using T = float;
    using C = float2;
    using PL = P2<LOGNH>;
    static constexpr int NH = PL::N, N = 2 * NH, TPL = PL::TPL, NT = TPL;
    static constexpr int PITCH = NH + PL::PAD;  // in {A, B} pairs of 16 bytes

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
                if constexpr (TWC && R == 16 && ido >= 64) {
                    // w^1..w^4 from the table, the other powers by multiplication (error a few ulp << the 1e-5 budget)
                    C w[9];
#pragma unroll
                    for (int q = 1; q <= 4; ++q) w[q] = __ldg(tw + (q - 1) * ido);
                    w[5] = cmul(w[4], w[1]); w[6] = cmul(w[4], w[2]); w[7] = cmul(w[4], w[3]); w[8] = cmul(w[4], w[4]);
#pragma unroll
                    for (int q = 1; q <= 8; ++q) {
                        a[j * R + q] = cmul(a[j * R + q], w[q]);
                        b[j * R + q] = cmul(b[j * R + q], w[q]);
                    }
#pragma unroll
                    for (int q = 9; q < 16; ++q) {
                        const C wq = cmul(w[8], w[q - 8]);
                        a[j * R + q] = cmul(a[j * R + q], wq);
                        b[j * R + q] = cmul(b[j * R + q], wq);
                    }
                } else {
#pragma unroll
                    for (int q = 1; q < R; ++q) {
                        const C w = __ldg(tw + (q - 1) * ido);  // one load serves both transforms
                        a[j * R + q] = cmul(a[j * R + q], w);
                        b[j * R + q] = cmul(b[j * R + q], w);
                    }
                }
            }
        }
    }

    // registers (outputs of pass P-1) -> shared -> registers (inputs of pass P), 16-byte {A, B} pairs
    template <int P>
    static __device__ __forceinline__ void exchange2(C *a, C *b, float4 *buf, int t, bool first) {
        constexpr int Rp = PL::radix(P - 1), NBp = 16 / Rp;
        constexpr int ido = PL::ido(P);
        if (!first) __syncthreads();
#pragma unroll
        for (int j = 0; j < NBp; ++j)
#pragma unroll
            for (int q = 0; q < Rp; ++q)
                buf[p2_phys<LOGNH, P>(t + j * TPL + q * (NH / Rp))] =
                    make_float4(a[j * Rp + q].x, a[j * Rp + q].y, b[j * Rp + q].x, b[j * Rp + q].y);
        __syncthreads();
        const int i = t % ido, k = t / ido;
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const float4 u = buf[p2_phys<LOGNH, P>(i + ido * (m + 16 * k))];
            a[m] = mk<T>(u.x, u.y);
            b[m] = mk<T>(u.z, u.w);
        }
    }

    static __device__ __forceinline__ void run(const TileGeom<T> &g, const C *__restrict__ stw, float4 *buf) {
        uint32_t t0, i1, i2, rest;
        fdivmod(blockIdx.x, g.d_t0, rest, t0);
        fdivmod(rest, g.d_e1, i2, i1);
        const int64_t in_base = (int64_t)t0 * g.in_bs[0] + (int64_t)i1 * g.in_bs[1] + (int64_t)i2 * g.in_bs[2];
        const int64_t out_base = (int64_t)t0 * g.out_bs[0] + (int64_t)i1 * g.out_bs[1] + (int64_t)i2 * g.out_bs[2];
        const int t = threadIdx.x;
        C a[16], b[16];
        prefetch_later_tile<T>(g, 1u);

        constexpr int R0 = PL::radix(0), NB0 = 16 / R0, ido0 = PL::ido(0);
        if (MODE == 1 || MODE == 0) {
            // x[4e .. 4e+3] = z[2e], z[2e+1] = element e of A and of B  (MODE 0: z is the complex line itself)
            const float4 *p = reinterpret_cast<const float4 *>(g.in + in_base) + t;
#pragma unroll
            for (int j = 0; j < NB0; ++j)
#pragma unroll
                for (int m = 0; m < R0; ++m) {
                    const float4 u = __ldcs(p + j * TPL + m * ido0);
                    a[j * R0 + m] = mk<T>(u.x, u.y);
                    b[j * R0 + m] = mk<T>(u.z, u.w);
                }
            if (MODE == 0 && g.backward) {
#pragma unroll
                for (int i = 0; i < 16; ++i) { a[i] = cswap(a[i]); b[i] = cswap(b[i]); }
            }
        } else {
            // Hermitian fold (see MODE 2 of pow2_kernel.cuh): Z[e] = s + i d conj(c_e) with s = X[e] + conj X[N-e],
            // d = X[e] - conj X[N-e], c_e = exp(-2 pi i e/(2N)); c_{e+N/2} = -i c_e.  Im X[0], Im X[N] are ignored
            // (reference: general_c2r, H:3830, 3845-3846); forward=true conjugates the input.
            const bool cj = g.backward == 0;
            const C *X = reinterpret_cast<const C *>(g.in + in_base);
            C ct = mk<T>(T(1), T(0));
            if (TWC) ct = __ldg(g.twA + t);  // c_e = c_t exp(-2 pi i (e - t)/(2N)), (e - t) a multiple of 2N/64
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                C xa[4], xb[4], ya[4], yb[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int idx = h * 4 + i;
                    const int e = t + (idx / R0) * TPL + (idx % R0) * ido0;
                    xa[i] = X[e];
                    xb[i] = X[N - e];
                    ya[i] = X[e + NH];
                    yb[i] = X[NH - e];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int idx = h * 4 + i;
                    const int e = t + (idx / R0) * TPL + (idx % R0) * ido0;
                    C A = xa[i], B = xb[i], A2 = ya[i], B2 = yb[i];
                    if (e == 0) { A.y = T(0); B.y = T(0); }
                    if (cj) { A.y = -A.y; B.y = -B.y; A2.y = -A2.y; B2.y = -B2.y; }
                    const C wc = TWC ? mul_root64(ct, (idx / R0) + (idx % R0) * NB0) : __ldg(g.twA + e);  // exp(-2 pi i e/(2N))
                    const C s = mk<T>(A.x + B.x, A.y - B.y), d = mk<T>(A.x - B.x, A.y + B.y);
                    const C wd = cmulc(d, wc);
                    const C z = mk<T>(s.x - wd.y, s.y + wd.x);  // s + i wd
                    const C s2 = mk<T>(A2.x + B2.x, A2.y - B2.y), d2 = mk<T>(A2.x - B2.x, A2.y + B2.y);
                    const C wd2 = cmulc(d2, wc);
                    const C z2 = mk<T>(s2.x - wd2.x, s2.y - wd2.y);  // s2 + i (i wd2)
                    const C pp = mk<T>(z.x + z2.x, z.y + z2.y);
                    const C qq = cmulc(mk<T>(z.x - z2.x, z.y - z2.y), cmul(wc, wc));  // (z - z2) exp(+2 pi i e/N)
                    a[idx] = cswap(pp);  // backward transform = swap . forward . swap
                    b[idx] = cswap(qq);
                }
            }
        }
        compute2<0>(a, b, t, stw);
        if constexpr (PL::NPASS > 1) { exchange2<1>(a, b, buf, t, true); compute2<1>(a, b, t, stw); }
        if constexpr (PL::NPASS > 2) { exchange2<2>(a, b, buf, t, false); compute2<2>(a, b, t, stw); }
        if constexpr (PL::NPASS > 3) { exchange2<3>(a, b, buf, t, false); compute2<3>(a, b, t, stw); }

        // ---- thread t now holds bins t + j*TPL + q*NH/RL of both transforms ------------------------------------
        constexpr int RL = PL::radix(PL::NPASS - 1), NBL = 16 / RL;
        if (MODE == 0) {
            // c2c: the radix-2 combination in registers, Z[k] = A[k] + w^k B[k], Z[k + N/2] = A[k] - w^k B[k];
            // w^k = w^t exp(-2 pi i q/32) for k = t + q*TPL (TPL = N/32): one table entry and compile-time roots
            static_assert(NBL == 1, "the last pass is radix 16");
            C *o = reinterpret_cast<C *>(g.out + out_base) + t;
            const C wt = __ldg(g.twA + t);  // exp(-2 pi i t/N)
            const T f = g.fct;
            const bool bw = g.backward != 0;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const C wb = cmul(mul_root64(wt, 2 * q), b[q]);
                C z0 = mk<T>((a[q].x + wb.x) * f, (a[q].y + wb.y) * f);
                C z1 = mk<T>((a[q].x - wb.x) * f, (a[q].y - wb.y) * f);
                if (bw) { z0 = cswap(z0); z1 = cswap(z1); }
                __stcs(o + q * TPL, z0);
                __stcs(o + q * TPL + NH, z1);
            }
            return;
        }
        if (MODE == 2) {
            float4 *o = reinterpret_cast<float4 *>(g.out + out_base) + t;
            const T f = g.fct;
#pragma unroll
            for (int j = 0; j < NBL; ++j)
#pragma unroll
                for (int q = 0; q < RL; ++q) {
                    const C va = a[j * RL + q], vb = b[j * RL + q];
                    __stcs(o + j * TPL + q * (NH / RL), make_float4(va.y * f, va.x * f, vb.y * f, vb.x * f));
                }
            return;
        }
        // r2c: radix-2 combination + Hermitian unpack from one trip through shared memory (index = bin)
        if constexpr (PL::NPASS > 1) __syncthreads();
#pragma unroll
        for (int j = 0; j < NBL; ++j)
#pragma unroll
            for (int q = 0; q < RL; ++q)
                buf[t + j * TPL + q * (NH / RL)] =
                    make_float4(a[j * RL + q].x, a[j * RL + q].y, b[j * RL + q].x, b[j * RL + q].y);
        __syncthreads();
        char *line = g.out + out_base;
        const T half = T(0.5) * g.fct;
        C ct = mk<T>(T(1), T(0));
        if (TWC) ct = __ldg(g.twA + t);  // c_k = c_t exp(-2 pi i j/64) for k = t + j*TPL (TPL = 2N/64)
        // the two output layouts and the two directions are separate straight-line copies of the loop
        // (one uniform branch per CTA instead of one per store)
        auto finish = [&](auto CONTIG, auto CONJ) {
            auto put = [&](int k, C x) {
                if (decltype(CONJ)::value) x.y = -x.y;
                if (decltype(CONTIG)::value) __stcs(reinterpret_cast<C *>(line) + k, x);
                else st_cx<T, true>(line + (int64_t)k * g.out_sa, x);
            };
            // X[k] = E + O, X[N-k] = conj(E - O), E = (Z[k] + conj Z[N-k])/2, O = -i c_k (Z[k] - conj Z[N-k])/2
            auto unpack = [&](int k, C za, C zp, C ck) {
                const C e = mk<T>((za.x + zp.x) * half, (za.y - zp.y) * half);
                const C d = mk<T>((za.x - zp.x) * half, (za.y + zp.y) * half);
                const C wd = cmul(ck, d);
                const C o = mk<T>(wd.y, -wd.x);
                put(k, mk<T>(e.x + o.x, e.y + o.y));
                put(N - k, mk<T>(e.x - o.x, -(e.y - o.y)));
            };
            auto item = [&](int j) {  // k = t + j*TPL
                const int k = t + j * TPL;
                const float4 u = buf[k], u2 = buf[(NH - k) & (NH - 1)];
                const C ak = mk<T>(u.x, u.y), bk = mk<T>(u.z, u.w), am = mk<T>(u2.x, u2.y), bm = mk<T>(u2.z, u2.w);
                const C c = TWC ? mul_root64(ct, j) : __ldg(g.twA + k);  // c_k = exp(-2 pi i k/(2N));  w^k = c_k^2
                const C w = cmul(c, c);
                const C wb = cmul(w, bk), cb = cmulc(bm, w);  // w^k B[k],  conj(w^k) B[N/2-k]
                // Z[k], Z[k+N/2]; Z[N-k] = A[N/2-k] + conj(w^k) B[N/2-k], Z[N/2-k] = A[N/2-k] - conj(w^k) B[N/2-k]
                const C zk = mk<T>(ak.x + wb.x, ak.y + wb.y), zkh = mk<T>(ak.x - wb.x, ak.y - wb.y);
                const C zp = mk<T>(am.x + cb.x, am.y + cb.y), zq = mk<T>(am.x - cb.x, am.y - cb.y);
                unpack(k, zk, zp, c);
                unpack(NH - k, zq, zkh, mk<T>(-c.y, -c.x));  // c_{N/2-k} = -i conj(c_k)
            };
#pragma unroll
            for (int j = 0; j < 8; ++j) item(j);
            if (t == 0) item(8);  // k = N/4 pairs with itself
        };
        const bool conj_out = g.backward != 0;
        if (g.out_sa == (int64_t)sizeof(C)) {
            if (conj_out) finish(BoolC<true>{}, BoolC<true>{});
            else finish(BoolC<true>{}, BoolC<false>{});
        } else {
            if (conj_out) finish(BoolC<false>{}, BoolC<true>{});
            else finish(BoolC<false>{}, BoolC<false>{});
        }
    }
};

template <int LOGNH, int MODE, bool TWC>
__global__ void __launch_bounds__((1 << LOGNH) / 16, (LOGNH >= 13 ? 1 : 2)) fft_pow2_dual_kernel(const TileGeom<float> g, const float2 *__restrict__ stw) {
    extern __shared__ __align__(16) unsigned char smem_raw_p2d[];
    DualBody<LOGNH, MODE, TWC>::run(g, stw, reinterpret_cast<float4 *>(smem_raw_p2d));
}

}  // namespace rfb
