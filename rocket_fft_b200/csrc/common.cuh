// Shared device/host helpers for the rocketfft_b200 kernels (sm_100a).
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#else
// NVRTC (run-time specialisation, jit.cu): no host headers
typedef unsigned char uint8_t;
typedef unsigned int uint32_t;
typedef int int32_t;
typedef unsigned long long uint64_t;
typedef long long int64_t;
typedef unsigned long size_t;
typedef unsigned long uintptr_t;
#endif

namespace rfb {

template <typename T> struct cx_of;
template <> struct cx_of<float> { using type = float2; };
template <> struct cx_of<double> { using type = double2; };
template <typename T> using cx = typename cx_of<T>::type;

template <typename T> __host__ __device__ __forceinline__ cx<T> mk(T a, T b) {
    cx<T> r; r.x = a; r.y = b; return r;
}
template <typename C> __device__ __forceinline__ C cadd(C a, C b) { a.x += b.x; a.y += b.y; return a; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { a.x -= b.x; a.y -= b.y; return a; }
#if defined(__CUDA_ARCH__) && defined(RFB_USE_F32X2)
// Blackwell packed single precision: one FADD2 per complex add/sub (PTX add/sub.f32x2, sm_100+).
// Measured on B200 (round 1): 208 FADD2 instead of 424 FADD per thread in the 8192-point kernel, but
// c2c c64 n=4096 dropped from 77.9 % to 70.2 % of HBM peak and the r2c row kernel gained < 1 % --
// the kernels are latency/occupancy limited, not FP32-issue limited -- so this stays opt-in.
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
    unsigned long long ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("add.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rd));
    return r;
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
    unsigned long long ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("sub.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(rd));
    return r;
}
#endif
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
    C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}
// a * conj(b)
template <typename C> __device__ __forceinline__ C cmulc(C a, C b) {
    C r; r.x = a.x * b.x + a.y * b.y; r.y = a.y * b.x - a.x * b.y; return r;
}
// multiply by -i  (forward quarter turn:  (x + iy)(-i) = y - ix)
template <typename C> __device__ __forceinline__ C mul_mi(C a) { C r; r.x = a.y; r.y = -a.x; return r; }
// multiply by +i
template <typename C> __device__ __forceinline__ C mul_pi(C a) { C r; r.x = -a.y; r.y = a.x; return r; }
template <typename C> __device__ __forceinline__ C cswap(C a) { C r; r.x = a.y; r.y = a.x; return r; }
template <typename C, typename T> __device__ __forceinline__ C cscale(C a, T s) { a.x *= s; a.y *= s; return a; }

// Exact unsigned division by a runtime constant for x < 2^31:  q = (x * mul) >> (31 + sh).
struct FastDiv {
    uint32_t d, mul, sh;
};
__host__ __device__ inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    uint32_t sh = 0;
    while ((1ull << sh) < d) ++sh;
    f.sh = sh;
    f.mul = (uint32_t)(((1ull << (31 + sh)) + d - 1) / d);
    return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t x, const FastDiv &f) {
    return (uint32_t)(((uint64_t)x * f.mul) >> (31 + f.sh));
}
__device__ __forceinline__ void fdivmod(uint32_t x, const FastDiv &f, uint32_t &q, uint32_t &r) {
    q = fdiv(x, f);
    r = x - q * f.d;
}

// Global loads/stores of one complex value at a byte address.  ALIGNED: the address is a
// multiple of sizeof(complex); otherwise only of sizeof(scalar).
template <typename T, bool ALIGNED>
__device__ __forceinline__ cx<T> ld_cx(const char *p) {
    if (ALIGNED) return *reinterpret_cast<const cx<T> *>(p);
    cx<T> r;
    r.x = reinterpret_cast<const T *>(p)[0];
    r.y = reinterpret_cast<const T *>(p)[1];
    return r;
}
template <typename T, bool ALIGNED>
__device__ __forceinline__ void st_cx(char *p, cx<T> v) {
    if (ALIGNED) { *reinterpret_cast<cx<T> *>(p) = v; return; }
    reinterpret_cast<T *>(p)[0] = v.x;
    reinterpret_cast<T *>(p)[1] = v.y;
}

}  // namespace rfb
