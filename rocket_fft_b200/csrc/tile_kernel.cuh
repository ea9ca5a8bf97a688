// Generic shared-memory line-FFT kernel: a CTA owns a tile of W lines of length n
// (any n that fits in shared memory, any factorisation into radices <= 64), runs an
// in-place decimation-in-frequency pass per factor inside shared memory and undoes the
// digit reversal while streaming the result out.  Global traffic is exactly one read
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
    FastDiv d_nout;
    int backward;        // 1: compute the +i transform through the swap identity
    int64_t in_sa, out_sa;           // axis strides, bytes
    uint32_t bext[MAXB];             // batch extents; dim 0 is the tile dim
    int64_t in_bs[MAXB], out_bs[MAXB];
    FastDiv d_t0, d_e1;              // tile id -> (t0, i1, i2)
    const char *in;
    char *out;
    const cx<T> *tw;     // exp(-2 pi i t / n), t in [0, n)
    T fct;
    // optional "four-step" factor on the output: out[k] *= exp(-2 pi i c k / bigN), c the
    // coordinate along batch dim tw_dim; exp(-2 pi i t/bigN) = twA[t / twS] * twB[t % twS]
    int tw_dim;          // -1: none
    FastDiv d_twS;
    const cx<T> *twA, *twB;
};

__device__ __forceinline__ uint32_t padidx(uint32_t e, uint32_t sh) { return e + (e >> sh); }

template <typename T, int R>
__device__ __forceinline__ void pass_fixed(cx<T> *buf, const PassInfo &ps, uint32_t W, uint32_t pitch,
                                           uint32_t padsh, const cx<T> *__restrict__ tw) {
    using C = cx<T>;
    const uint32_t ido = ps.ido, l1 = ps.l1;
    const uint32_t nbl = ps.d_nbl.d;
    const uint32_t total = W * nbl;
    for (uint32_t bb = threadIdx.x; bb < total; bb += blockDim.x) {
        uint32_t w, b, k, i;
        fdivmod(bb, ps.d_nbl, w, b);
        fdivmod(b, ps.d_ido, k, i);
        C *base = buf + (size_t)w * pitch;
        const uint32_t e0 = i + ido * R * k;
        C v[R];
#pragma unroll
        for (int m = 0; m < R; ++m) v[m] = base[padidx(e0 + ido * m, padsh)];
        Dft<T, R>::run(v);
        if (ido > 1 && i > 0) {
            const uint32_t t = l1 * i;
#pragma unroll
            for (int q = 1; q < R; ++q) v[q] = cmul(v[q], __ldg(tw + t * q));
        }
#pragma unroll
        for (int q = 0; q < R; ++q) base[padidx(e0 + ido * q, padsh)] = v[q];
    }
}

// Any radix up to RMAX_GENERIC: O(R^2) butterfly with the roots read from the line table.
template <typename T>
__device__ __noinline__ void pass_generic(cx<T> *buf, const PassInfo &ps, uint32_t n, uint32_t W,
                                          uint32_t pitch, uint32_t padsh, const cx<T> *__restrict__ tw) {
    using C = cx<T>;
    const uint32_t R = ps.R, ido = ps.ido, l1 = ps.l1;
    const uint32_t nbl = ps.d_nbl.d;
    const uint32_t total = W * nbl;
    const uint32_t rs = n / R;  // exp(-2 pi i j / R) = tw[j * rs]
    for (uint32_t bb = threadIdx.x; bb < total; bb += blockDim.x) {
        uint32_t w, b, k, i;
        fdivmod(bb, ps.d_nbl, w, b);
        fdivmod(b, ps.d_ido, k, i);
        C *base = buf + (size_t)w * pitch;
        const uint32_t e0 = i + ido * R * k;
        C v[RMAX_GENERIC], o[RMAX_GENERIC];
        for (uint32_t m = 0; m < R; ++m) v[m] = base[padidx(e0 + ido * m, padsh)];
        for (uint32_t q = 0; q < R; ++q) {
            C acc = v[0];
            uint32_t j = 0;
            for (uint32_t m = 1; m < R; ++m) {
                j += q;
                if (j >= R) j -= R;
                C r = __ldg(tw + j * rs);
                acc.x += v[m].x * r.x - v[m].y * r.y;
                acc.y += v[m].x * r.y + v[m].y * r.x;
            }
            o[q] = acc;
        }
        const uint32_t t = l1 * i;
        base[padidx(e0, padsh)] = o[0];
        for (uint32_t q = 1; q < R; ++q) {
            C val = o[q];
            if (ido > 1 && i > 0) val = cmul(val, __ldg(tw + t * q));
            base[padidx(e0 + ido * q, padsh)] = val;
        }
    }
}

template <typename T, bool ALIGNED>
__global__ void __launch_bounds__(512) fft_tile_kernel(const TileGeom<T> g) {
    using C = cx<T>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C *buf = reinterpret_cast<C *>(smem_raw);

    // ---- which tile -----------------------------------------------------------------
    uint32_t t0, i1, i2, rest;
    fdivmod(blockIdx.x, g.d_t0, rest, t0);
    fdivmod(rest, g.d_e1, i2, i1);
    const uint32_t w_first = t0 * g.W;
    const uint32_t wvalid = min(g.W, g.bext[0] - w_first);
    const int64_t in_base = (int64_t)w_first * g.in_bs[0] + (int64_t)i1 * g.in_bs[1] + (int64_t)i2 * g.in_bs[2];
    const int64_t out_base = (int64_t)w_first * g.out_bs[0] + (int64_t)i1 * g.out_bs[1] + (int64_t)i2 * g.out_bs[2];
    const uint32_t n = g.n, W = g.W, pitch = g.pitch, padsh = g.padsh;
    const uint32_t tile_elems = W * n;

    // ---- load ---------------------------------------------------------------------------
    for (uint32_t idx = threadIdx.x; idx < tile_elems; idx += blockDim.x) {
        uint32_t w, e;
        if (g.load_line_fast) fdivmod(idx, g.d_W, e, w);
        else fdivmod(idx, g.d_n, w, e);
        C val = mk<T>(T(0), T(0));
        if (w < wvalid) {
            val = load_value<T, ALIGNED>(g.load_mode, g.flags, g.in + in_base + (int64_t)w * g.in_bs[0], g.in_sa, e, n,
                                         g.n_in);
            if (g.backward) val = cswap(val);
        }
        buf[(size_t)w * pitch + padidx(e, padsh)] = val;
    }
    __syncthreads();

    // ---- passes -------------------------------------------------------------------------
    for (uint32_t s = 0; s < g.npass; ++s) {
        const PassInfo &ps = g.pass[s];
        switch (ps.R) {
            case 2: pass_fixed<T, 2>(buf, ps, W, pitch, padsh, g.tw); break;
            case 3: pass_fixed<T, 3>(buf, ps, W, pitch, padsh, g.tw); break;
            case 4: pass_fixed<T, 4>(buf, ps, W, pitch, padsh, g.tw); break;
            case 5: pass_fixed<T, 5>(buf, ps, W, pitch, padsh, g.tw); break;
            case 7: pass_fixed<T, 7>(buf, ps, W, pitch, padsh, g.tw); break;
            case 8: pass_fixed<T, 8>(buf, ps, W, pitch, padsh, g.tw); break;
            case 11: pass_fixed<T, 11>(buf, ps, W, pitch, padsh, g.tw); break;
            case 13: pass_fixed<T, 13>(buf, ps, W, pitch, padsh, g.tw); break;
            case 16: pass_fixed<T, 16>(buf, ps, W, pitch, padsh, g.tw); break;
            default: pass_generic<T>(buf, ps, n, W, pitch, padsh, g.tw); break;
        }
        __syncthreads();
    }

    // ---- store (undo the digit reversal on the fly) -------------------------------------
    const uint32_t n_out = g.n_out;
    const uint32_t out_elems = W * n_out;
    for (uint32_t idx = threadIdx.x; idx < out_elems; idx += blockDim.x) {
        uint32_t w, j;
        if (g.store_line_fast) fdivmod(idx, g.d_W, j, w);
        else fdivmod(idx, g.d_nout, w, j);
        if (w >= wvalid) continue;
        const uint32_t k = store_bin(g.store_mode, j);
        uint32_t rem = k, p = 0;
        for (uint32_t s = 0; s < g.npass; ++s) {
            uint32_t q, r;
            fdivmod(rem, g.pass[s].d_R, q, r);
            p += r * g.pass[s].ido;
            rem = q;
        }
        C val = buf[(size_t)w * pitch + padidx(p, padsh)];
        if (g.tw_dim >= 0) {
            const uint32_t c = (g.tw_dim == 0) ? (w_first + w) : (g.tw_dim == 1 ? i1 : i2);
            uint32_t hi, lo;
            fdivmod(c * k, g.d_twS, hi, lo);
            val = cmul(val, cmul(__ldg(g.twA + hi), __ldg(g.twB + lo)));
        }
        val = cscale(val, g.fct);
        if (g.backward) val = cswap(val);
        store_value<T, ALIGNED>(g.store_mode, g.flags, g.out + out_base + (int64_t)w * g.out_bs[0], g.out_sa, j, val);
    }
}

}  // namespace rfb
