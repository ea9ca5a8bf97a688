// How one element of a line is fetched from / delivered to global memory for every
// load/store mode of the line engine.  Shared by the tile kernel (fused) and the
// stand-alone gather/scatter kernels (long lines that go through scratch).
#pragma once
#include "common.cuh"
#include "modes.h"

namespace rfb {

// element e (0 <= e < n) of the complex line the transform sees
// (aux: exp(-2 pi i t / (4 N')) table for the DCT / DST modes: N' = N, or 2N for LD_G_DCT4)
template <typename T, bool ALIGNED>
__device__ __forceinline__ cx<T> load_value(int mode, int flags, const char *line, int64_t sa, uint32_t e,
                                            uint32_t n, uint32_t n_in, const cx<T> *__restrict__ aux = nullptr) {
    using C = cx<T>;
    C val = mk<T>(T(0), T(0));
    auto xr = [&](uint32_t i) { return *reinterpret_cast<const T *>(line + (int64_t)i * sa); };
    const bool sine = (flags & FLAG_SINE) != 0, ortho = (flags & FLAG_ORTHO) != 0;
    switch (mode) {
        case LD_G_DCT1: {
            const uint32_t N = n / 2 + 1, idx = e <= N - 1 ? e : n - e;
            val.x = xr(idx);
            if (ortho && (idx == 0 || idx == N - 1)) val.x *= T(1.4142135623730951);
            break;
        }
        case LD_G_DST1: {
            const uint32_t N = n / 2 - 1;
            if (e >= 1 && e <= N) val.x = xr(e - 1);
            else if (e > N + 1) val.x = -xr(n - e - 1);
            break;
        }
        case LD_G_DCT2: {
            const uint32_t N = n, idx = e < (N + 1) / 2 ? 2 * e : 2 * (N - 1 - e) + 1;
            val.x = xr(idx);
            if (sine && (idx & 1)) val.x = -val.x;  // DST-II(x)_k = DCT-II((-1)^j x_j)_{N-1-k}
            break;
        }
        case LD_G_DCT3: {
            // sine: DST-III(x)_k = (-1)^k DCT-III(x reversed)_k.  ortho scales one element of the caller's line by sqrt 2:
            // index 0 for the cosine transform; for the sine transform index 0 as the reference does (H:3038-3039) or,
            // without FLAG_QUIRK, index N-1 as SciPy does
            const uint32_t N = n, scaled = sine ? ((flags & FLAG_QUIRK) ? 0u : N - 1) : 0u;
            auto xs = [&](uint32_t i) {
                const uint32_t ci = sine ? N - 1 - i : i;
                T r = xr(ci);
                if (ortho && ci == scaled) r *= T(1.4142135623730951);
                return r;
            };
            const T a = xs(e), b = e ? xs(N - e) : T(0);
            const C w = __ldg(aux + e);  // (c, -s) = exp(-i pi e / 2N); the factor is its conjugate
            val = mk<T>(a * w.x - b * w.y, -a * w.y - b * w.x);
            break;
        }
        case LD_G_DCT4: {
            const uint32_t N = 2 * n;
            const T a = xr(sine ? N - 1 - 2 * e : 2 * e), b = xr(sine ? 2 * e : N - 1 - 2 * e);
            val = cmul(mk<T>(a, b), __ldg(aux + (4 * e + 1)));  // aux: exp(-2 pi i t / (8N))
            break;
        }
        case LD_G_DCT4Z: {
            const uint32_t N = n / 2;
            if (e < N) {
                const T a = xr(sine ? N - 1 - e : e);
                const C w = __ldg(aux + e);  // exp(-i pi e / 2N)
                val = mk<T>(a * w.x, a * w.y);
            }
            break;
        }
        case LD_C2C:
            if (e < n_in) val = ld_cx<T, ALIGNED>(line + (int64_t)e * sa);
            break;
        case LD_REAL:
            if (e < n_in) {
                val.x = *reinterpret_cast<const T *>(line + (int64_t)e * sa);
                if ((flags & FLAG_NEG_EVEN_IN) && e >= 2 && !(e & 1)) val.x = -val.x;
            }
            break;
        case LD_HERM: {
            // Hermitian extension of bins 0..n/2; Im of bin 0 (and of bin n/2, n even) ignored
            // (reference: general_c2r, _pocketfft_hdronly.h:3830, 3845-3846); bins >= n_in are zero padding
            const bool upper = 2 * e > n;
            const uint32_t k = upper ? n - e : e;
            if (k < n_in) {
                val = ld_cx<T, ALIGNED>(line + (int64_t)k * sa);
                if (k == 0 || 2 * k == n) val.y = T(0);
                if (upper) val.y = -val.y;
            }
            break;
        }
        case LD_HC: {
            // FFTPACK halfcomplex [r0, r1, i1, r2, i2, ..., (r_{n/2})]
            const bool upper = 2 * e > n;
            const uint32_t k = upper ? n - e : e;
            auto at = [&](uint32_t j) { return *reinterpret_cast<const T *>(line + (int64_t)j * sa); };
            val.x = (k == 0) ? at(0) : at(2 * k - 1);
            val.y = (k == 0 || 2 * k == n) ? T(0) : at(2 * k);
            if (upper) val.y = -val.y;
            break;
        }
    }
    return val;
}

// which spectrum bin feeds output slot j
__device__ __forceinline__ uint32_t store_bin(int mode, uint32_t j) {
    return mode == ST_HC ? (j + 1) >> 1 : j;
}

// deliver output slot j given the (already scaled, un-swapped) spectrum value
template <typename T, bool ALIGNED>
__device__ __forceinline__ void store_value(int mode, int flags, char *line, int64_t sa, uint32_t j, cx<T> val) {
    char *p = line + (int64_t)j * sa;
    switch (mode) {
        case ST_C2C:
        case ST_HALF: st_cx<T, ALIGNED>(p, val); break;
        case ST_REAL: {
            T r = val.x;
            if ((flags & FLAG_NEG_EVEN_OUT) && j >= 2 && !(j & 1)) r = -r;
            *reinterpret_cast<T *>(p) = r;
            break;
        }
        case ST_HC: *reinterpret_cast<T *>(p) = (j == 0 || (j & 1)) ? val.x : val.y; break;
        case ST_HARTLEY: *reinterpret_cast<T *>(p) = val.x + val.y; break;
    }
}

}  // namespace rfb
