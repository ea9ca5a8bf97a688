// Register butterflies: forward DFTs of a few points, X[q] = sum_m x[m] exp(-2 pi i q m / R),
// input and output in natural order in v[0..R).  Backward transforms are obtained by the
// callers through the re/im swap identity IDFT(x) = swap(DFT(swap(x))).
//
// Replaces (by different means) the pass2/3/4/5/7/8/11 and passg butterflies of the
// reference's cfftp (rocket_fft/_pocketfft_hdronly.h:1079-1688).
#pragma once
#include "common.cuh"

namespace rfb {

template <typename T> struct K {
    static constexpr T SQRT1_2 = T(0.7071067811865476);
    static constexpr T C_PI_8 = T(0.9238795325112867);
    static constexpr T S_PI_8 = T(0.3826834323650898);
};

// cos(2 pi j / R), sin(2 pi j / R), j = 1..(R-1)/2, correctly rounded to double.
template <int R> struct PrimeTab;
template <> struct PrimeTab<3> {
    __host__ __device__ static constexpr double c(int j) {
        constexpr double t[1] = {-0.5};
        return t[j];
    }
    __host__ __device__ static constexpr double s(int j) {
        constexpr double t[1] = {0.8660254037844386};
        return t[j];
    }
};
template <> struct PrimeTab<5> {
    __host__ __device__ static constexpr double c(int j) {
        constexpr double t[2] = {0.30901699437494745, -0.8090169943749475};
        return t[j];
    }
    __host__ __device__ static constexpr double s(int j) {
        constexpr double t[2] = {0.9510565162951535, 0.5877852522924731};
        return t[j];
    }
};
template <> struct PrimeTab<7> {
    __host__ __device__ static constexpr double c(int j) {
        constexpr double t[3] = {0.6234898018587335, -0.2225209339563144, -0.9009688679024191};
        return t[j];
    }
    __host__ __device__ static constexpr double s(int j) {
        constexpr double t[3] = {0.7818314824680298, 0.9749279121818236, 0.4338837391175581};
        return t[j];
    }
};
template <> struct PrimeTab<11> {
    __host__ __device__ static constexpr double c(int j) {
        constexpr double t[5] = {0.8412535328311812, 0.41541501300188644, -0.14231483827328514,
                                    -0.6548607339452851, -0.9594929736144974};
        return t[j];
    }
    __host__ __device__ static constexpr double s(int j) {
        constexpr double t[5] = {0.5406408174555976, 0.9096319953545183, 0.9898214418809327,
                                    0.7557495743542583, 0.28173255684142967};
        return t[j];
    }
};
template <> struct PrimeTab<13> {
    __host__ __device__ static constexpr double c(int j) {
        constexpr double t[6] = {0.8854560256532099, 0.5680647467311558, 0.12053668025532305,
                                    -0.3546048870425356, -0.7485107481711011, -0.970941817426052};
        return t[j];
    }
    __host__ __device__ static constexpr double s(int j) {
        constexpr double t[6] = {0.46472317204376856, 0.8229838658936564, 0.992708874098054,
                                    0.9350162426854148, 0.6631226582407952, 0.23931566428755777};
        return t[j];
    }
};

template <typename T, int R> struct Dft;

template <typename T> struct Dft<T, 2> {
    using C = cx<T>;
    static __device__ __forceinline__ void run(C *v) {
        C a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};

template <typename T> struct Dft<T, 4> {
    using C = cx<T>;
    static __device__ __forceinline__ void run(C *v) {
        C t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
        C t2 = cadd(v[1], v[3]), t3 = mul_mi(csub(v[1], v[3]));
        v[0] = cadd(t0, t2);
        v[1] = cadd(t1, t3);
        v[2] = csub(t0, t2);
        v[3] = csub(t1, t3);
    }
};

template <typename T> struct Dft<T, 8> {
    using C = cx<T>;
    static __device__ __forceinline__ void run(C *v) {
        C e[4] = {v[0], v[2], v[4], v[6]};
        C o[4] = {v[1], v[3], v[5], v[7]};
        Dft<T, 4>::run(e);
        Dft<T, 4>::run(o);
        const T h = K<T>::SQRT1_2;
        // o[k] *= exp(-2 pi i k / 8)
        C o1 = mk<T>((o[1].x + o[1].y) * h, (o[1].y - o[1].x) * h);
        C o2 = mul_mi(o[2]);
        C o3 = mk<T>((o[3].y - o[3].x) * h, -(o[3].x + o[3].y) * h);
        v[0] = cadd(e[0], o[0]); v[4] = csub(e[0], o[0]);
        v[1] = cadd(e[1], o1);   v[5] = csub(e[1], o1);
        v[2] = cadd(e[2], o2);   v[6] = csub(e[2], o2);
        v[3] = cadd(e[3], o3);   v[7] = csub(e[3], o3);
    }
};

template <typename T> struct Dft<T, 16> {
    using C = cx<T>;
    static __device__ __forceinline__ void run(C *v) {
        C e[8], o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
        Dft<T, 8>::run(e);
        Dft<T, 8>::run(o);
        const T h = K<T>::SQRT1_2, c1 = K<T>::C_PI_8, s1 = K<T>::S_PI_8;
        // w^k = exp(-2 pi i k/16) = (cos, -sin)
        C w[8];
        w[0] = mk<T>(T(1), T(0));
        w[1] = mk<T>(c1, -s1);
        w[2] = mk<T>(h, -h);
        w[3] = mk<T>(s1, -c1);
        w[4] = mk<T>(T(0), T(-1));
        w[5] = mk<T>(-s1, -c1);
        w[6] = mk<T>(-h, -h);
        w[7] = mk<T>(-c1, -s1);
        v[0] = cadd(e[0], o[0]);
        v[8] = csub(e[0], o[0]);
#pragma unroll
        for (int k = 1; k < 8; ++k) {
            C t = (k == 4) ? mul_mi(o[4]) : cmul(o[k], w[k]);
            v[k] = cadd(e[k], t);
            v[k + 8] = csub(e[k], t);
        }
    }
};

// Odd prime radix with compile-time tables: pairs (m, R-m).
template <typename T, int R> struct DftPrime {
    using C = cx<T>;
    static __device__ __forceinline__ void run(C *v) {
        constexpr int H = (R - 1) / 2;
        C a[H], b[H];
#pragma unroll
        for (int m = 0; m < H; ++m) {
            a[m] = cadd(v[m + 1], v[R - 1 - m]);
            b[m] = csub(v[m + 1], v[R - 1 - m]);
        }
        C x0 = v[0];
        C sum = x0;
#pragma unroll
        for (int m = 0; m < H; ++m) sum = cadd(sum, a[m]);
        v[0] = sum;
#pragma unroll
        for (int q = 1; q <= H; ++q) {
            T pr = x0.x, pi = x0.y, qr = T(0), qi = T(0);
#pragma unroll
            for (int m = 1; m <= H; ++m) {
                int j = (q * m) % R;
                // cos(2 pi j/R), sin(2 pi j/R) from the half table
                T c = (j <= H) ? T(PrimeTab<R>::c(j - 1)) : T(PrimeTab<R>::c(R - j - 1));
                T s = (j <= H) ? T(PrimeTab<R>::s(j - 1)) : T(-PrimeTab<R>::s(R - j - 1));
                pr += a[m - 1].x * c; pi += a[m - 1].y * c;
                qr += b[m - 1].x * s; qi += b[m - 1].y * s;
            }
            // X[q] = P - i*Q ; X[R-q] = P + i*Q      (-i*(qr + i qi) = qi - i qr)
            v[q] = mk<T>(pr + qi, pi - qr);
            v[R - q] = mk<T>(pr - qi, pi + qr);
        }
    }
};
template <typename T> struct Dft<T, 3> : DftPrime<T, 3> {};
template <typename T> struct Dft<T, 5> : DftPrime<T, 5> {};
template <typename T> struct Dft<T, 7> : DftPrime<T, 7> {};
template <typename T> struct Dft<T, 11> : DftPrime<T, 11> {};
template <typename T> struct Dft<T, 13> : DftPrime<T, 13> {};

// cos / sin of 2 pi j / R for the composite radices (correctly rounded doubles)
template <int R> struct RootTab;
template <> struct RootTab<6> {
    __host__ __device__ static constexpr double c(int j) {
        constexpr double t[6] = {1.0, 0.5, -0.5, -1.0, -0.5, 0.5};
        return t[j];
    }
    __host__ __device__ static constexpr double s(int j) {
        constexpr double t[6] = {0.0, 0.8660254037844386, 0.8660254037844386, 0.0, -0.8660254037844386, -0.8660254037844386};
        return t[j];
    }
};
template <> struct RootTab<9> {
    __host__ __device__ static constexpr double c(int j) {
        constexpr double t[9] = {1.0, 0.766044443118978, 0.17364817766693036, -0.5, -0.9396926207859084, -0.9396926207859084, -0.5, 0.17364817766693036, 0.766044443118978};
        return t[j];
    }
    __host__ __device__ static constexpr double s(int j) {
        constexpr double t[9] = {0.0, 0.6427876096865394, 0.984807753012208, 0.8660254037844386, 0.3420201433256687, -0.3420201433256687, -0.8660254037844386, -0.984807753012208, -0.6427876096865394};
        return t[j];
    }
};
template <> struct RootTab<10> {
    __host__ __device__ static constexpr double c(int j) {
        constexpr double t[10] = {1.0, 0.8090169943749475, 0.30901699437494745, -0.30901699437494745, -0.8090169943749475, -1.0, -0.8090169943749475, -0.30901699437494745, 0.30901699437494745, 0.8090169943749475};
        return t[j];
    }
    __host__ __device__ static constexpr double s(int j) {
        constexpr double t[10] = {0.0, 0.5877852522924731, 0.9510565162951535, 0.9510565162951535, 0.5877852522924731, 0.0, -0.5877852522924731, -0.9510565162951535, -0.9510565162951535, -0.5877852522924731};
        return t[j];
    }
};
template <> struct RootTab<12> {
    __host__ __device__ static constexpr double c(int j) {
        constexpr double t[12] = {1.0, 0.8660254037844386, 0.5, 0.0, -0.5, -0.8660254037844386, -1.0, -0.8660254037844386, -0.5, 0.0, 0.5, 0.8660254037844386};
        return t[j];
    }
    __host__ __device__ static constexpr double s(int j) {
        constexpr double t[12] = {0.0, 0.5, 0.8660254037844386, 1.0, 0.8660254037844386, 0.5, 0.0, -0.5, -0.8660254037844386, -1.0, -0.8660254037844386, -0.5};
        return t[j];
    }
};
template <> struct RootTab<15> {
    __host__ __device__ static constexpr double c(int j) {
        constexpr double t[15] = {1.0, 0.9135454576426009, 0.6691306063588582, 0.30901699437494745, -0.10452846326765347, -0.5, -0.8090169943749475, -0.9781476007338057, -0.9781476007338057, -0.8090169943749475, -0.5, -0.10452846326765347, 0.30901699437494745, 0.6691306063588582, 0.9135454576426009};
        return t[j];
    }
    __host__ __device__ static constexpr double s(int j) {
        constexpr double t[15] = {0.0, 0.4067366430758002, 0.7431448254773942, 0.9510565162951535, 0.9945218953682733, 0.8660254037844386, 0.5877852522924731, 0.20791169081775934, -0.20791169081775934, -0.5877852522924731, -0.8660254037844386, -0.9945218953682733, -0.9510565162951535, -0.7431448254773942, -0.4067366430758002};
        return t[j];
    }
};

// Composite radix R = P*Q in registers (Cooley-Tukey): Q transforms of size P over x[Q p + q], the
// internal factors exp(-2 pi i q k1 / R), then P transforms of size Q; X[k1 + P k2].
template <typename T, int P, int Q> struct DftComposite {
    using C = cx<T>;
    static __device__ __forceinline__ void run(C *v) {
        constexpr int R = P * Q;
        C a[Q][P];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
#pragma unroll
            for (int p = 0; p < P; ++p) a[q][p] = v[Q * p + q];
            Dft<T, P>::run(a[q]);
        }
#pragma unroll
        for (int q = 1; q < Q; ++q)
#pragma unroll
            for (int k1 = 1; k1 < P; ++k1) {
                constexpr int dummy = 0;
                (void)dummy;
                const int j = (q * k1) % R;
                a[q][k1] = cmul(a[q][k1], mk<T>(T(RootTab<R>::c(j)), T(-RootTab<R>::s(j))));
            }
#pragma unroll
        for (int k1 = 0; k1 < P; ++k1) {
            C b[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) b[q] = a[q][k1];
            Dft<T, Q>::run(b);
#pragma unroll
            for (int k2 = 0; k2 < Q; ++k2) v[k1 + P * k2] = b[k2];
        }
    }
};
template <typename T> struct Dft<T, 6> : DftComposite<T, 3, 2> {};
template <typename T> struct Dft<T, 9> : DftComposite<T, 3, 3> {};
template <typename T> struct Dft<T, 10> : DftComposite<T, 5, 2> {};
template <typename T> struct Dft<T, 12> : DftComposite<T, 4, 3> {};
template <typename T> struct Dft<T, 15> : DftComposite<T, 5, 3> {};

}  // namespace rfb
