// Generic shared-memory line-FFT kernel: a CTA owns a tile of W lines of length n
// (any n that fits in shared memory, any factorisation into radices <= 64) and runs one
// in-place decimation-in-frequency pass per factor.  The FIRST pass reads its butterfly
// inputs straight from global memory (element i + m*n/R: coalesced) and the LAST pass writes
// its outputs straight to global memory: thread u computes the butterfly whose outputs are the
// bins u + q*n/R (coalesced), found through the mixed-radix digit reversal of u.  Global traffic is exactly one read
// and one write of the tile, coalesced either along the line ("element-fast", lines are
// contiguous) or across neighbouring lines ("line-fast", lines are strided but the
// tile's W lines are adjacent in memory -- no transposed copy is ever materialised).
//
// This is the GPU counterpart of general_nd + multi_iter + copy_input/copy_output +
// cfftp::pass_all of the reference (rocket_fft/_pocketfft_hdronly.h:3568-3607,
// 3272-3353, 3496-3553, 1690-1740); the algorithm (in-place DIF + permuted store) and
// the code are new.
#pragma once
#include "common.cuh"
#include "line_io.cuh"
#include "radix.cuh"

namespace rfb {

constexpr int MAXP = 16;        // passes per line (n <= 65536 with radix >= 2 ... >= 16 passes never needed)
constexpr int MAXB = 3;         // batch dims handled inside one launch
constexpr int RMAX_GENERIC = 64;  // largest prime radix done as a direct O(R^2) butterfly

struct PassInfo {
    uint32_t R, ido, l1;
    uint32_t twoff;  // offset of this pass's twiddles in the pass-major table ptw
    FastDiv d_ido;   // butterfly -> (k, i)
    FastDiv d_nbl;   // flat butterfly index -> (line, butterfly) ; nbl = n / R
    FastDiv d_R;     // digit extraction for the output permutation
};

template <typename T>
struct TileGeom {
    uint32_t n, npass;
    PassInfo pass[MAXP];
    uint32_t W;          // lines per tile
    uint32_t pitch;      // padded line pitch (elements)
    uint32_t padsh;      // pad one element every 2^padsh (31: none)
    FastDiv d_n, d_W;
    int load_line_fast, store_line_fast;
    int load_mode, store_mode, flags;   // line_io.cuh
    uint32_t n_in;       // number of input elements actually present per line (<= n; rest zero)
    uint32_t n_out;      // number of output slots stored per line
    uint32_t pf_dist, pf_bytes;  // L2 prefetch of the input of the CTA pf_dist tiles ahead (0: off), bytes per tile
    uint32_t pf_rows;            // line-fast tiles: rows of pf_bytes each (0: the tile is one run of pf_bytes)
    uint32_t stage_io;           // pow2 kernel, short lines: bit 0 / 1 = load / store the dense tile through shared memory
    FastDiv d_nout;
    int backward;        // 1: compute the +i transform through the swap identity
    int64_t in_sa, out_sa;           // axis strides, bytes
    uint32_t bext[MAXB];             // batch extents; dim 0 is the tile dim
    int64_t in_bs[MAXB], out_bs[MAXB];
    FastDiv d_t0, d_e1;              // tile id -> (t0, i1, i2)
    const char *in;
    char *out;
    const cx<T> *tw;     // exp(-2 pi i t / n), t in [0, n)   (roots of the generic-radix butterfly)
    const cx<T> *ptw;    // pass-major twiddles: pass s, entry [(q-1)*ido + i] = exp(-2 pi i i q / (R ido))
    T fct;
    // optional "four-step" factor on the output: out[k] *= exp(-2 pi i c k / bigN), c the
    // coordinate along batch dim tw_dim; exp(-2 pi i t/bigN) = twA[t / twS] * twB[t % twS]
    int tw_dim;          // -1: none
    FastDiv d_twS;
    const cx<T> *twA, *twB;
    // scatter of the output axis over several destination buffers (peer GPUs): bin k -> split_base[k / split_blk]
    uint32_t split_blk;  // 0: off
    FastDiv d_split;
    char *split_base[16];
    // fused element-wise factors (power-of-two kernel only), see LineJob
    int c_dim;           // batch dim whose coordinate enters g = e * g_mul + c * c_mul (-1: c = 0)
    uint32_t g_mul, c_mul, pre_bound, post_bound;
    const cx<T> *pre_tab, *post_tab;
    int pre_swap, post_swap;
    // tables of the DCT / DST load / store modes (line_io.cuh)
    const cx<T> *aux_ld, *aux_st;
};

// Ticket order of the fused four-step kernel: unit u of 2*S units -> (step, strip).  A(0) .. A(lag-1), then the pairs
// A(lag + i), B(i), then the last lag B's.  B(s) depends on A(s); A(s), s >= ring, on B(s - ring).  With ring > lag every
// unit comes after the units it depends on (checked on the host by tests/test_abi.py through rfb200_debug_fuse4_unit; used by fused4v2_kernel.cuh).
__host__ __device__ inline bool fuse4_decode_unit(uint32_t unit, uint32_t S, uint32_t lag, bool &stepB, uint32_t &strip) {
    if (unit >= 2 * S) return false;
    if (unit < lag) { stepB = false; strip = unit; }
    else if (unit < lag + 2 * (S - lag)) { const uint32_t v = unit - lag; stepB = (v & 1u) != 0; strip = stepB ? (v >> 1) : lag + (v >> 1); }
    else { stepB = true; strip = S - lag + (unit - lag - 2 * (S - lag)); }
    return true;
}

// The CTAs resident on an SM tend to load, compute and store in step, which leaves HBM idle while they compute.
// One thread asks the L2 to fetch the input of the tile that will run pf_dist tiles from now (one bulk-prefetch
// instruction for the whole tile): DRAM streams in the background and the later CTA's loads hit L2.
// Measured on B200: c2c complex128 4096 x 4096 lines 74.5 % -> 81.6 % of the HBM copy peak.
template <typename T>
__device__ __forceinline__ void prefetch_later_tile(const TileGeom<T> &g, uint32_t W) {
    if (g.pf_dist == 0) return;
    if (g.pf_rows == 0 && threadIdx.x != 0) return;
    const uint32_t nb = blockIdx.x + g.pf_dist;
    if (nb >= gridDim.x) return;
    uint32_t a0, a1, a2, r2;
    fdivmod(nb, g.d_t0, r2, a0);
    fdivmod(r2, g.d_e1, a2, a1);
    if ((a0 + 1) * W > g.bext[0]) return;  // partial tile at the edge: its span could leave the array
    const char *p = g.in + (int64_t)(a0 * W) * g.in_bs[0] + (int64_t)a1 * g.in_bs[1] + (int64_t)a2 * g.in_bs[2];
    if (g.pf_rows) {
        // line-fast tile: pf_rows rows of pf_bytes contiguous bytes (the W neighbouring lines), in_sa apart
        for (uint32_t i = threadIdx.x; i < g.pf_rows; i += blockDim.x) {
            const char *row = p + (int64_t)i * g.in_sa;
            for (uint32_t o = 0; o < g.pf_bytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + o));
        }
        return;
    }
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uintptr_t lo = (a + 15) & ~(uintptr_t)15, hi = (a + g.pf_bytes) & ~(uintptr_t)15;  // 16-byte granules inside the tile
    if (hi > lo) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(lo), "r"((uint32_t)(hi - lo)) : "memory");
}

__device__ __forceinline__ uint32_t padidx(uint32_t e, uint32_t sh) { return e + (e >> sh); }

// deliver spectrum bin f (all store modes); line_io.cuh's store_value works by output slot
template <typename T, bool ALIGNED>
__device__ __forceinline__ void store_bin_value(int mode, int flags, char *line, int64_t sa, uint32_t f, uint32_t n,
                                                cx<T> val, const cx<T> *__restrict__ aux = nullptr) {
    auto put = [&](uint32_t i, T r) { *reinterpret_cast<T *>(line + (int64_t)i * sa) = r; };
    const bool sine = (flags & FLAG_SINE) != 0, ortho = (flags & FLAG_ORTHO) != 0;
    switch (mode) {
        case ST_G_DCT1: {
            const uint32_t N = n / 2 + 1;
            if (f >= N) return;
            put(f, (ortho && (f == 0 || f == N - 1)) ? val.x * T(0.70710678118654752) : val.x);
            return;
        }
        case ST_G_DST1: {
            const uint32_t N = n / 2 - 1;
            if (f < 1 || f > N) return;
            put(f - 1, -val.y);
            return;
        }
        case ST_G_DCT2: {
            // ortho scales one output by 1/sqrt 2: index 0 (cosine); sine: the caller's index 0 as the reference does
            // (H:3033-3034) or, without FLAG_QUIRK, N-1 as SciPy does
            const uint32_t N = n, k = sine ? N - 1 - f : f, scaled = sine ? ((flags & FLAG_QUIRK) ? 0u : N - 1) : 0u;
            const cx<T> w = __ldg(aux + f);
            T r = T(2) * (w.x * val.x - w.y * val.y);
            if (ortho && k == scaled) r *= T(0.70710678118654752);
            put(k, r);
            return;
        }
        case ST_G_DCT3: {
            const uint32_t N = n, idx = f < (N + 1) / 2 ? 2 * f : 2 * (N - 1 - f) + 1;
            put(idx, (sine && (idx & 1)) ? -val.x : val.x);
            return;
        }
        case ST_G_DCT4: {
            const uint32_t N = 2 * n;
            const cx<T> w = cmul(val, __ldg(aux + 4 * f));  // aux: exp(-2 pi i t / (8N)); exp(-i pi f / N)
            put(2 * f, T(2) * w.x);
            put(N - 1 - 2 * f, sine ? T(2) * w.y : T(-2) * w.y);  // sine: (-1)^k with k = N-1-2f odd
            return;
        }
        case ST_G_DCT4Z: {
            const uint32_t N = n / 2;
            if (f >= N) return;
            const cx<T> w = __ldg(aux + (2 * f + 1));  // aux: exp(-2 pi i t / (8N))
            const T r = T(2) * (w.x * val.x - w.y * val.y);
            put(f, (sine && (f & 1)) ? -r : r);
            return;
        }
        default: break;
    }
    switch (mode) {
        case ST_HALF:
            if (2 * f > n) return;
        case ST_C2C: st_cx<T, ALIGNED>(line + (int64_t)f * sa, val); break;
        case ST_REAL: {
            T r = val.x;
            if ((flags & FLAG_NEG_EVEN_OUT) && f >= 2 && !(f & 1)) r = -r;
            *reinterpret_cast<T *>(line + (int64_t)f * sa) = r;
            break;
        }
        case ST_HARTLEY: *reinterpret_cast<T *>(line + (int64_t)f * sa) = val.x + val.y; break;
        case ST_HC:
            if (2 * f > n) return;
            if (f == 0) *reinterpret_cast<T *>(line) = val.x;
            else {
                *reinterpret_cast<T *>(line + (int64_t)(2 * f - 1) * sa) = val.x;
                if (2 * f < n) *reinterpret_cast<T *>(line + (int64_t)(2 * f) * sa) = val.y;
            }
            break;
    }
}

struct TileCtx {
    uint32_t w_first, wvalid, i1, i2;
    int64_t in_base, out_base;
};

// forward DFT of R points in v (natural order in and out); R == 0: runtime radix through roots table
template <typename T, int R>
__device__ __forceinline__ void butterfly(cx<T> *v, uint32_t Rr, uint32_t n, const cx<T> *__restrict__ roots) {
    if constexpr (R != 0) Dft<T, R>::run(v);
    else {
        using C = cx<T>;
        C o[RMAX_GENERIC];
        const uint32_t rs = n / Rr;
        for (uint32_t q = 0; q < Rr; ++q) {
            C acc = v[0];
            uint32_t j = 0;
            for (uint32_t m = 1; m < Rr; ++m) {
                j += q;
                if (j >= Rr) j -= Rr;
                const C r = __ldg(roots + j * rs);
                acc.x += v[m].x * r.x - v[m].y * r.y;
                acc.y += v[m].x * r.y + v[m].y * r.x;
            }
            o[q] = acc;
        }
        for (uint32_t q = 0; q < Rr; ++q) v[q] = o[q];
    }
}

enum { PASS_FIRST = 0, PASS_MID = 1, PASS_LAST = 2, PASS_ONLY = 3 };

template <typename T, int R, int WHERE, bool ALIGNED>
__device__ __forceinline__ void run_pass(const TileGeom<T> &g, const TileCtx &cx_, cx<T> *buf, const PassInfo &ps) {
    using C = cx<T>;
    constexpr int RA = R ? R : RMAX_GENERIC;
    const uint32_t Rr = R ? (uint32_t)R : ps.R;
    const uint32_t n = g.n, W = g.W, pitch = g.pitch, padsh = g.padsh;
    const uint32_t ido = ps.ido, l1 = ps.l1;
    const uint32_t nbl = ps.d_nbl.d;
    const uint32_t total = W * nbl;
    const bool from_global = (WHERE == PASS_FIRST || WHERE == PASS_ONLY);
    const bool to_global = (WHERE == PASS_LAST || WHERE == PASS_ONLY);
    const bool lf = to_global ? (g.store_line_fast != 0) : (from_global ? (g.load_line_fast != 0) : false);
    for (uint32_t bb = threadIdx.x; bb < total; bb += blockDim.x) {
        uint32_t w, b;
        if (lf) fdivmod(bb, g.d_W, b, w);
        else fdivmod(bb, ps.d_nbl, w, b);
        C v[RA];
        uint32_t i, e0, u = b;
        if (to_global) {
            // b = u is the low part of the output bin; the butterfly sits at the digit-reversed place
            uint32_t rem = u, p = 0;
            for (uint32_t s = 0; s + 1 < g.npass; ++s) {
                uint32_t q, r;
                fdivmod(rem, g.pass[s].d_R, q, r);
                p += r * g.pass[s].ido;
                rem = q;
            }
            i = 0;
            e0 = p;
        } else {
            uint32_t k;
            fdivmod(b, ps.d_ido, k, i);
            e0 = i + ido * Rr * k;
        }
        if (from_global) {
            const bool ok = w < cx_.wvalid;
            const char *line = g.in + cx_.in_base + (int64_t)w * g.in_bs[0];
#pragma unroll
            for (uint32_t m = 0; m < Rr; ++m) {
                C val = mk<T>(T(0), T(0));
                if (ok) {
                    val = load_value<T, ALIGNED>(g.load_mode, g.flags, line, g.in_sa, e0 + ido * m, n, g.n_in, g.aux_ld);
                    if (g.backward) val = cswap(val);
                }
                v[m] = val;
            }
        } else {
            const C *base = buf + (size_t)w * pitch;
#pragma unroll
            for (uint32_t m = 0; m < Rr; ++m) v[m] = base[padidx(e0 + ido * m, padsh)];
        }
        butterfly<T, R>(v, Rr, n, g.tw);
        if (ido > 1 && i > 0) {
            const C *tw = g.ptw + ps.twoff + i;
#pragma unroll
            for (uint32_t q = 1; q < Rr; ++q) v[q] = cmul(v[q], __ldg(tw + (q - 1) * ido));
        }
        if (to_global) {
            if (w >= cx_.wvalid) continue;
            char *line = g.out + cx_.out_base + (int64_t)w * g.out_bs[0];
            const uint32_t c = (g.tw_dim == 0) ? (cx_.w_first + w) : (g.tw_dim == 1 ? cx_.i1 : cx_.i2);
#pragma unroll
            for (uint32_t q = 0; q < Rr; ++q) {
                const uint32_t f = u + l1 * q;
                C val = v[q];
                if (g.tw_dim >= 0) {
                    uint32_t hi, lo;
                    fdivmod(c * f, g.d_twS, hi, lo);
                    val = cmul(val, cmul(__ldg(g.twA + hi), __ldg(g.twB + lo)));
                }
                val = cscale(val, g.fct);
                if (g.backward) val = cswap(val);
                store_bin_value<T, ALIGNED>(g.store_mode, g.flags, line, g.out_sa, f, n, val, g.aux_st);
            }
        } else {
            C *base = buf + (size_t)w * pitch;
#pragma unroll
            for (uint32_t q = 0; q < Rr; ++q) base[padidx(e0 + ido * q, padsh)] = v[q];
        }
    }
}

template <typename T, int WHERE, bool ALIGNED>
__device__ __forceinline__ void dispatch_pass(const TileGeom<T> &g, const TileCtx &cx_, cx<T> *buf, const PassInfo &ps) {
    switch (ps.R) {
        case 2: run_pass<T, 2, WHERE, ALIGNED>(g, cx_, buf, ps); break;
        case 3: run_pass<T, 3, WHERE, ALIGNED>(g, cx_, buf, ps); break;
        case 4: run_pass<T, 4, WHERE, ALIGNED>(g, cx_, buf, ps); break;
        case 5: run_pass<T, 5, WHERE, ALIGNED>(g, cx_, buf, ps); break;
        case 7: run_pass<T, 7, WHERE, ALIGNED>(g, cx_, buf, ps); break;
        case 8: run_pass<T, 8, WHERE, ALIGNED>(g, cx_, buf, ps); break;
        case 11: run_pass<T, 11, WHERE, ALIGNED>(g, cx_, buf, ps); break;
        case 13: run_pass<T, 13, WHERE, ALIGNED>(g, cx_, buf, ps); break;
        case 16: run_pass<T, 16, WHERE, ALIGNED>(g, cx_, buf, ps); break;
        default: run_pass<T, 0, WHERE, ALIGNED>(g, cx_, buf, ps); break;
    }
}

// float tiles run with up to 1024 threads (one big tile per SM needs the warps to hide latency),
// double tiles with up to 512 (register budget of the radix-16 butterfly)
template <typename T> constexpr int tile_max_threads() { return sizeof(T) == 4 ? 1024 : 512; }

template <typename T, bool ALIGNED>
__global__ void __launch_bounds__(tile_max_threads<T>()) fft_tile_kernel(const TileGeom<T> g) {
    using C = cx<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);

    TileCtx cx_;
    uint32_t t0, rest;
    fdivmod(blockIdx.x, g.d_t0, rest, t0);
    fdivmod(rest, g.d_e1, cx_.i2, cx_.i1);
    cx_.w_first = t0 * g.W;
    cx_.wvalid = min(g.W, g.bext[0] - cx_.w_first);
    cx_.in_base = (int64_t)cx_.w_first * g.in_bs[0] + (int64_t)cx_.i1 * g.in_bs[1] + (int64_t)cx_.i2 * g.in_bs[2];
    cx_.out_base = (int64_t)cx_.w_first * g.out_bs[0] + (int64_t)cx_.i1 * g.out_bs[1] + (int64_t)cx_.i2 * g.out_bs[2];

    if (g.npass == 1) {
        dispatch_pass<T, PASS_ONLY, ALIGNED>(g, cx_, buf, g.pass[0]);
        return;
    }
    dispatch_pass<T, PASS_FIRST, ALIGNED>(g, cx_, buf, g.pass[0]);
    __syncthreads();
    for (uint32_t s = 1; s + 1 < g.npass; ++s) {
        dispatch_pass<T, PASS_MID, ALIGNED>(g, cx_, buf, g.pass[s]);
        __syncthreads();
    }
    dispatch_pass<T, PASS_LAST, ALIGNED>(g, cx_, buf, g.pass[g.npass - 1]);
}

}  // namespace rfb
