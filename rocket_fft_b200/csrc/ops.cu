// N-D drivers: what each entry point of the C ABI does in terms of line jobs.
// Semantics follow the reference's public templates (_pocketfft_hdronly.h:3875-4091) and
// its boundary file (_pocketfft_numba.cpp:25-223); see SURVEY.md section 8(a).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"
#include "engine.h"

namespace rfb {

static bool g_dst_quirk = true;
void set_dst_ortho_quirk(bool on) { g_dst_quirk = on; }

#define RFB_AFTER_LAUNCH2() RFB_CUDA_CHECK(cudaGetLastError())
void count_launch(const char *name = nullptr);

static bool any_zero(const std::vector<int64_t> &s) {
    for (auto v : s)
        if (v == 0) return true;
    return false;
}

static std::vector<Dim> batch_dims(const std::vector<int64_t> &shape, const std::vector<int64_t> &is,
                                   const std::vector<int64_t> &os, size_t axis) {
    std::vector<Dim> b;
    for (size_t d = 0; d < shape.size(); ++d)
        if (d != axis) b.push_back(Dim{shape[d], is[d], os[d], false});
    return b;
}

static std::vector<int64_t> c_strides(const std::vector<int64_t> &shape, int64_t esz) {
    std::vector<int64_t> st(shape.size());
    int64_t acc = esz;
    for (size_t i = shape.size(); i-- > 0;) {
        st[i] = acc;
        acc *= std::max<int64_t>(shape[i], 1);
    }
    return st;
}

// C-order strides of a scratch array whose rows (innermost dim) are padded to a multiple of 128 bytes: a half spectrum has
// n/2 + 1 points per row -- 8193 x 8 B for the 16384^2 images -- and dense rows would start at a different 8-byte phase each,
// which costs the strided kernels their 16-byte accesses and sector-aligned row segments (measured on the slab rfftn,
// where the same padding halved the time, and on the column kernel: B tiles 0.40 -> 0.24 ms with aligned rows).
// `bytes` receives the allocation size.
static std::vector<int64_t> c_strides_padded(const std::vector<int64_t> &shape, int64_t esz, uint64_t &bytes) {
    std::vector<int64_t> padded = shape;
    static const bool no_pad = [] { const char *e = getenv("RFB200_NO_PADDED_SCRATCH"); return e && atoi(e) != 0; }();  // (A/B aid)
    if (!no_pad && !shape.empty() && shape.back() * esz >= 1024) {
        const int64_t per = 128 / esz;
        padded.back() = (shape.back() + per - 1) / per * per;
    }
    std::vector<int64_t> st(shape.size());
    int64_t acc = esz;
    for (size_t i = shape.size(); i-- > 0;) {
        st[i] = acc;
        acc *= std::max<int64_t>(padded[i], 1);
    }
    bytes = (uint64_t)acc;
    return st;
}

static uint64_t prod(const std::vector<int64_t> &s) {
    uint64_t p = 1;
    for (auto v : s) p *= (uint64_t)v;
    return p;
}

// ---------------------------------------------------------------------------------------
// c2c  (reference: c2c H:3875-3889 -> general_nd H:3568-3607: axes in the given order,
// first axis reads `in`, later ones work on `out`; fct applied once)
// ---------------------------------------------------------------------------------------
static void c2c_one_axis(int prec, const std::vector<int64_t> &shape, const std::vector<int64_t> &si,
                         const std::vector<int64_t> &sout, size_t ax, const char *in, char *out, bool forward, double fct,
                         cudaStream_t s) {
    LineJob j;
    j.prec = prec;
    j.n = (uint64_t)shape[ax];
    j.in = in;
    j.out = out;
    j.is = si[ax];
    j.os = sout[ax];
    j.batch = batch_dims(shape, si, sout, ax);
    j.backward = !forward;
    j.fct = fct;
    run_lines(j, s);
}

// (Measured and dropped, profiles/r02l_l2_blocked_axis_groups.log: running the passes over axes (1, 2) of a 1024^3 volume slice
// group by slice group -- 16 / 32 / 64 MiB at a time, so that the second pass finds its input in L2 -- is SLOWER than two
// full walks, 12.6 / 10.5 / 9.5 ms against 7.2 ms: a strided-lines tile occupies an SM for ~9 us and a slice group is
// only 3.5 waves of them; the launches' ramps and tails cost more than the saved HBM round trip.)
static void c2c_axes(int prec, const std::vector<int64_t> &shape, const std::vector<int64_t> &sin,
                     const std::vector<int64_t> &sout, const uint64_t *axes, size_t naxes, const char *in, char *out,
                     bool forward, double fct, cudaStream_t s) {
    bool first = true;
    for (size_t i = 0; i < naxes; ++i) {
        c2c_one_axis(prec, shape, first ? sin : sout, sout, (size_t)axes[i], first ? in : out, out, forward, first ? fct : 1.0, s);
        first = false;
    }
}

void op_c2c(const NdArgs &a, bool forward, cudaStream_t s) {
    if (any_zero(a.shape)) return;
    c2c_axes(a.prec, a.shape, a.sin, a.sout, a.axes.data(), a.axes.size(), a.in, a.out, forward, a.fct, s);
}

// c2c along one axis with the output axis scattered over several buffers (the fused
// "local transform + all-to-all push" of the slab-decomposed fftn; no reference counterpart)
void op_c2c_scatter(const NdArgs &a, size_t axis, bool forward, const std::vector<char *> &parts, cudaStream_t s) {
    if (any_zero(a.shape)) return;
    const uint64_t n = (uint64_t)a.shape[axis];
    if (parts.empty() || n % parts.size()) { set_error("axis length must be divisible by the number of parts"); throw Error(); }
    LineJob j;
    j.prec = a.prec;
    j.n = n;
    j.in = a.in;
    j.out = nullptr;
    j.is = a.sin[axis];
    j.os = a.sout[axis];
    j.batch = batch_dims(a.shape, a.sin, a.sout, axis);
    j.backward = !forward;
    j.fct = a.fct;
    j.split_blk = n / parts.size();
    for (size_t h = 0; h < parts.size(); ++h)
        j.split_out.push_back(parts[h] - (int64_t)(h * j.split_blk) * a.sout[axis]);
    run_lines(j, s);
}

// ---------------------------------------------------------------------------------------
// r2c  (reference: r2c H:3955-3975: real transform along axes.back(), then c2c in place on
// the half-spectrum array over the remaining axes, fct applied by the real transform)
// ---------------------------------------------------------------------------------------
static void r2c_into(int prec, const std::vector<int64_t> &shape_in, const std::vector<int64_t> &sin,
                     const std::vector<int64_t> &sout, const std::vector<uint64_t> &axes, const char *in, char *out,
                     bool forward, double fct, cudaStream_t s) {
    const size_t L = (size_t)axes.back();
    LineJob j;
    j.prec = prec;
    j.n = (uint64_t)shape_in[L];
    j.in = in;
    j.out = out;
    j.is = sin[L];
    j.os = sout[L];
    j.batch = batch_dims(shape_in, sin, sout, L);
    j.backward = !forward;
    j.fct = fct;
    j.load_mode = LD_REAL;
    j.store_mode = ST_HALF;
    std::vector<int64_t> shape_out = shape_in;
    shape_out[L] = shape_in[L] / 2 + 1;
    // Three or more axes on a dense output whose rows are not multiples of 128 bytes (513 points for a 1024-point real axis):
    // the passes between the real transform and the last one would walk that array in place along strided axes with every
    // row at another 8-byte phase.  They run on a scratch copy with padded rows instead (c_strides_padded); the last pass
    // reads the scratch and writes the caller's array.  (Same arithmetic, same order of the axes.)
    const int64_t esz = prec ? 16 : 8;
    static const bool no_pad = [] { const char *e = getenv("RFB200_NO_PADDED_SCRATCH"); return e && atoi(e) != 0; }();  // (A/B aid)
    const bool padded_path = !no_pad && axes.size() >= 3 && L + 1 == shape_out.size() && sout[L] == esz && (shape_out[L] * esz) % 128 != 0 &&
                             shape_out[L] * esz >= 1024;
    if (padded_path) {
        uint64_t tbytes = 0;
        const std::vector<int64_t> st = c_strides_padded(shape_out, esz, tbytes);
        Scratch tmp(tbytes, s);
        j.out = (char *)tmp.p;
        j.os = st[L];
        j.batch = batch_dims(shape_in, sin, st, L);
        run_lines(j, s);
        const size_t nc = axes.size() - 1;  // complex passes
        c2c_axes(prec, shape_out, st, st, axes.data(), nc - 1, (const char *)tmp.p, (char *)tmp.p, forward, 1.0, s);
        c2c_one_axis(prec, shape_out, st, sout, (size_t)axes[nc - 1], (const char *)tmp.p, out, forward, 1.0, s);
        return;
    }
    run_lines(j, s);
    if (axes.size() > 1) c2c_axes(prec, shape_out, sout, sout, axes.data(), axes.size() - 1, out, out, forward, 1.0, s);
}

void op_r2c(const NdArgs &a, bool forward, cudaStream_t s) {
    if (any_zero(a.shape) || a.axes.empty()) return;
    r2c_into(a.prec, a.shape, a.sin, a.sout, a.axes, a.in, a.out, forward, a.fct, s);
}

// ---------------------------------------------------------------------------------------
// c2r  (reference: c2r H:3995-4023: c2c over axes[:-1] into a contiguous temporary, then the
// Hermitian -> real transform along axes.back(); shape is the OUTPUT's)
// ---------------------------------------------------------------------------------------
void op_c2r(const NdArgs &a, bool forward, cudaStream_t s) {
    if (any_zero(a.shape) || a.axes.empty()) return;
    const size_t L = (size_t)a.axes.back();
    const int64_t esz = a.prec ? 16 : 8;
    const char *src = a.in;
    std::vector<int64_t> ssrc = a.sin;
    std::vector<int64_t> shape_in = a.shape;
    shape_in[L] = a.shape[L] / 2 + 1;
    Scratch *tmp = nullptr;
    struct Guard { Scratch *&p; ~Guard() { delete p; } } guard{tmp};
    if (a.axes.size() > 1) {
        uint64_t tbytes = 0;
        std::vector<int64_t> st = c_strides_padded(shape_in, esz, tbytes);
        tmp = new Scratch(tbytes, s);
        c2c_axes(a.prec, shape_in, a.sin, st, a.axes.data(), a.axes.size() - 1, a.in, (char *)tmp->p, forward, 1.0, s);
        src = (const char *)tmp->p;
        ssrc = st;
    }
    LineJob j;
    j.prec = a.prec;
    j.n = (uint64_t)a.shape[L];
    j.in = src;
    j.out = a.out;
    j.is = ssrc[L];
    j.os = a.sout[L];
    j.batch = batch_dims(a.shape, ssrc, a.sout, L);
    j.backward = !forward;
    j.fct = a.fct;
    j.load_mode = LD_HERM;
    j.store_mode = ST_REAL;
    run_lines(j, s);
}

// ---------------------------------------------------------------------------------------
// Zero-padded / cropped input: the `n` / `s` arguments of the numpy / scipy layer.  The reference
// materialises a padded copy on the host first (rocket_fft/overloads.py:575-609); here the padding is
// part of the first load of every line (LineJob::n_in) and lines that consist of padding only are
// never touched: `valid[d]` is the extent of real data along d, a pass over axis a runs only over the
// valid extents of the other dims and makes a's extent fully valid.
// ---------------------------------------------------------------------------------------
static std::vector<int64_t> valid_extents(const std::vector<int64_t> &shape_in, const std::vector<int64_t> &shape,
                                          const std::vector<uint64_t> &axes) {
    if (shape_in.size() != shape.size()) { set_error("input and transform shapes differ in rank"); throw Error(); }
    std::vector<int64_t> valid(shape.size());
    for (size_t d = 0; d < shape.size(); ++d) {
        bool is_axis = false;
        for (auto ax : axes) is_axis = is_axis || (ax == d);
        if (!is_axis && shape_in[d] != shape[d]) {
            set_error("input and transform shapes may differ only along transformed axes");
            throw Error();
        }
        valid[d] = std::min(shape_in[d], shape[d]);
    }
    return valid;
}

static void c2c_axes_pruned(int prec, const std::vector<int64_t> &shape, std::vector<int64_t> valid,
                            const std::vector<int64_t> &sin, const std::vector<int64_t> &sout, const uint64_t *axes,
                            size_t naxes, const char *in, char *out, bool forward, double fct, cudaStream_t s) {
    bool first = true;
    for (size_t i = 0; i < naxes; ++i) {
        const size_t ax = (size_t)axes[i];
        LineJob j;
        j.prec = prec;
        j.n = (uint64_t)shape[ax];
        j.n_in = (uint64_t)valid[ax];
        j.in = first ? in : out;
        j.out = out;
        const auto &si = first ? sin : sout;
        j.is = si[ax];
        j.os = sout[ax];
        j.batch = batch_dims(valid, si, sout, ax);
        j.backward = !forward;
        j.fct = first ? fct : 1.0;
        run_lines(j, s);
        valid[ax] = shape[ax];
        first = false;
    }
}

// a.shape: the transform (= output) shape
void op_c2c_pad(const NdArgs &a, const std::vector<int64_t> &shape_in, bool forward, cudaStream_t s) {
    if (any_zero(a.shape) || any_zero(shape_in) || a.axes.empty()) return;
    c2c_axes_pruned(a.prec, a.shape, valid_extents(shape_in, a.shape, a.axes), a.sin, a.sout, a.axes.data(),
                    a.axes.size(), a.in, a.out, forward, a.fct, s);
}

// a.shape: the (padded / cropped) real shape that is transformed; output extent along axes.back() is n/2+1
void op_r2c_pad(const NdArgs &a, const std::vector<int64_t> &shape_in, bool forward, cudaStream_t s) {
    if (any_zero(a.shape) || any_zero(shape_in) || a.axes.empty()) return;
    std::vector<int64_t> valid = valid_extents(shape_in, a.shape, a.axes);
    const size_t L = (size_t)a.axes.back();
    LineJob j;
    j.prec = a.prec;
    j.n = (uint64_t)a.shape[L];
    j.n_in = (uint64_t)valid[L];
    j.in = a.in;
    j.out = a.out;
    j.is = a.sin[L];
    j.os = a.sout[L];
    j.batch = batch_dims(valid, a.sin, a.sout, L);
    j.backward = !forward;
    j.fct = a.fct;
    j.load_mode = LD_REAL;
    j.store_mode = ST_HALF;
    run_lines(j, s);
    if (a.axes.size() > 1) {
        std::vector<int64_t> shape_out = a.shape;
        shape_out[L] = a.shape[L] / 2 + 1;
        valid[L] = shape_out[L];
        c2c_axes_pruned(a.prec, shape_out, valid, a.sout, a.sout, a.axes.data(), a.axes.size() - 1, a.out, a.out, forward,
                        1.0, s);
    }
}

// a.shape: the real output shape; shape_in: the complex input as it is (bins beyond it are zero,
// bins beyond n/2 along axes.back() are ignored)
void op_c2r_pad(const NdArgs &a, const std::vector<int64_t> &shape_in, bool forward, cudaStream_t s) {
    if (any_zero(a.shape) || any_zero(shape_in) || a.axes.empty()) return;
    const size_t L = (size_t)a.axes.back();
    const int64_t esz = a.prec ? 16 : 8;
    std::vector<int64_t> cshape = a.shape;
    cshape[L] = a.shape[L] / 2 + 1;
    std::vector<int64_t> valid = valid_extents(shape_in, cshape, a.axes);
    const char *src = a.in;
    std::vector<int64_t> ssrc = a.sin;
    Scratch *tmp = nullptr;
    struct Guard { Scratch *&p; ~Guard() { delete p; } } guard{tmp};
    if (a.axes.size() > 1) {
        // only the bins that exist along L are carried through the leading axes
        std::vector<int64_t> tshape = cshape;
        tshape[L] = valid[L];
        std::vector<int64_t> st = c_strides(tshape, esz);
        tmp = new Scratch(prod(tshape) * (uint64_t)esz, s);
        c2c_axes_pruned(a.prec, tshape, valid, a.sin, st, a.axes.data(), a.axes.size() - 1, a.in, (char *)tmp->p, forward,
                        1.0, s);
        for (size_t i = 0; i + 1 < a.axes.size(); ++i) valid[a.axes[i]] = tshape[a.axes[i]];
        src = (const char *)tmp->p;
        ssrc = st;
    }
    LineJob j;
    j.prec = a.prec;
    j.n = (uint64_t)a.shape[L];
    j.n_in = (uint64_t)valid[L];
    j.in = src;
    j.out = a.out;
    j.is = ssrc[L];
    j.os = a.sout[L];
    j.batch = batch_dims(a.shape, ssrc, a.sout, L);
    j.backward = !forward;
    j.fct = a.fct;
    j.load_mode = LD_HERM;
    j.store_mode = ST_REAL;
    run_lines(j, s);
}

// ---------------------------------------------------------------------------------------
// roll: out[(i + shift) mod n] = in[i] along every dim -- fftshift (shift = n/2), ifftshift
// (shift = -(n/2)) and np.roll of the layer above the path (reference: rocket_fft/overloads.py:752-855,
// 1221-1310).  One pass over HBM: iteration follows the output (last dim fastest), a thread moves one item.
// ---------------------------------------------------------------------------------------
struct RollGeom {
    int nd;
    uint32_t ext[8], back[8];  // back = (n - shift) mod n: source index = (j + back) mod n
    FastDiv d[8];
    int64_t si[8], so[8];
};

template <typename V>
__global__ void __launch_bounds__(256) roll_kernel(const RollGeom g, uint32_t total, const char *__restrict__ in,
                                                   char *__restrict__ out) {
    for (uint32_t f = blockIdx.x * 256u + threadIdx.x; f < total; f += gridDim.x * 256u) {
        uint32_t rem = f;
        int64_t src = 0, dst = 0;
#pragma unroll 1
        for (int d = g.nd - 1; d >= 0; --d) {
            uint32_t q, j;
            fdivmod(rem, g.d[d], q, j);
            rem = q;
            uint32_t i = j + g.back[d];
            if (i >= g.ext[d]) i -= g.ext[d];
            src += (int64_t)i * g.si[d];
            dst += (int64_t)j * g.so[d];
        }
        *reinterpret_cast<V *>(out + dst) = *reinterpret_cast<const V *>(in + src);
    }
}

void op_roll(int64_t item, const std::vector<int64_t> &shape, const std::vector<int64_t> &sin,
             const std::vector<int64_t> &sout, const std::vector<int64_t> &shift, const char *in, char *out,
             cudaStream_t s) {
    if (any_zero(shape)) return;
    if (item != 4 && item != 8 && item != 16) { set_error("roll: item size must be 4, 8 or 16 bytes"); throw Error(); }
    if (shape.size() != shift.size() || shape.size() != sin.size() || shape.size() != sout.size()) {
        set_error("roll: shape / strides / shift differ in rank");
        throw Error();
    }
    bool al = ((uintptr_t)in % item) == 0 && ((uintptr_t)out % item) == 0;
    for (size_t d = 0; d < shape.size(); ++d) al = al && (sin[d] % item) == 0 && (sout[d] % item) == 0;
    if (!al) { set_error("roll: arrays must be aligned to the item size"); throw Error(); }
    // drop unit dims; more than 8 dims or >= 2^32 items: peel the outermost dim on the host
    std::vector<int64_t> sh, si, so, sf;
    for (size_t d = 0; d < shape.size(); ++d)
        if (shape[d] > 1) {
            sh.push_back(shape[d]);
            si.push_back(sin[d]);
            so.push_back(sout[d]);
            sf.push_back(((shift[d] % shape[d]) + shape[d]) % shape[d]);
        }
    if (sh.empty()) { sh = {1}; si = {item}; so = {item}; sf = {0}; }
    if (sh.size() > 8 || prod(sh) >= (1ull << 32) || sh[0] >= (1ll << 32)) {
        if (sh.size() == 1) { set_error("roll: dimension too long"); throw Error(); }
        const int64_t n0 = sh[0];
        std::vector<int64_t> sh1(sh.begin() + 1, sh.end()), si1(si.begin() + 1, si.end()), so1(so.begin() + 1, so.end()),
            sf1(sf.begin() + 1, sf.end());
        for (int64_t j = 0; j < n0; ++j) {
            const int64_t i = (j + n0 - sf[0]) % n0;
            op_roll(item, sh1, si1, so1, sf1, in + i * si[0], out + j * so[0], s);
        }
        return;
    }
    RollGeom g;
    g.nd = (int)sh.size();
    for (int d = 0; d < g.nd; ++d) {
        g.ext[d] = (uint32_t)sh[d];
        g.back[d] = (uint32_t)((sh[d] - sf[d]) % sh[d]);
        g.d[d] = make_fastdiv((uint32_t)sh[d]);
        g.si[d] = si[d];
        g.so[d] = so[d];
    }
    const uint32_t total = (uint32_t)prod(sh);
    const unsigned blocks = (unsigned)std::min<uint64_t>(((uint64_t)total + 255) / 256, 148ull * 64);
    if (item == 4) roll_kernel<uint32_t><<<blocks, 256, 0, s>>>(g, total, in, out);
    else if (item == 8) roll_kernel<uint2><<<blocks, 256, 0, s>>>(g, total, in, out);
    else roll_kernel<uint4><<<blocks, 256, 0, s>>>(g, total, in, out);
    RFB_AFTER_LAUNCH2();
    count_launch();
}

// ---------------------------------------------------------------------------------------
// data[l][j] *= table[j] for contiguous lines (real x real or complex x complex): the spectral
// coefficient multiply of the fast Hankel transform and its bias factors (reference:
// rocket_fft/overloads.py:880-901, 1768-1780: `A *= u`, `a * exp(-bias ...)`)
// ---------------------------------------------------------------------------------------
template <typename T, bool CPLX>
__global__ void __launch_bounds__(256) scale_lines_kernel(uint64_t nlines, uint32_t n, const T *__restrict__ table,
                                                          T *__restrict__ data) {
    const uint32_t j = blockIdx.x * 256u + threadIdx.x;
    if (j >= n) return;
    if constexpr (CPLX) {
        using C = cx<T>;
        const C u = __ldg(reinterpret_cast<const C *>(table) + j);
        for (uint64_t l = blockIdx.y; l < nlines; l += gridDim.y) {
            C *p = reinterpret_cast<C *>(data) + l * n + j;
            *p = cmul(*p, u);
        }
    } else {
        const T u = __ldg(table + j);
        for (uint64_t l = blockIdx.y; l < nlines; l += gridDim.y) data[l * n + j] *= u;
    }
}

void op_scale_lines(int prec, bool cplx, uint64_t nlines, uint64_t n, const void *table, void *data, cudaStream_t s) {
    if (nlines == 0 || n == 0) return;
    if (n >= (1ull << 31)) { set_error("scale_lines: line too long"); throw Error(); }
    const dim3 grid((unsigned)((n + 255) / 256), (unsigned)std::min<uint64_t>(nlines, 16384));
    if (prec) {
        if (cplx) scale_lines_kernel<double, true><<<grid, 256, 0, s>>>(nlines, (uint32_t)n, (const double *)table, (double *)data);
        else scale_lines_kernel<double, false><<<grid, 256, 0, s>>>(nlines, (uint32_t)n, (const double *)table, (double *)data);
    } else {
        if (cplx) scale_lines_kernel<float, true><<<grid, 256, 0, s>>>(nlines, (uint32_t)n, (const float *)table, (float *)data);
        else scale_lines_kernel<float, false><<<grid, 256, 0, s>>>(nlines, (uint32_t)n, (const float *)table, (float *)data);
    }
    RFB_AFTER_LAUNCH2();
    count_launch();
}

// ---------------------------------------------------------------------------------------
// N-D index helper for the mirror / Hartley-combine kernels
// ---------------------------------------------------------------------------------------
struct NdIdx {
    int nd;
    uint32_t ext[8];    // iteration extents (dim nd-1 fastest)
    uint32_t full[8];   // full extents
    FastDiv d[8];
    int64_t sa[8], sb[8];
    int rev[8];         // index negated (mod full) along this dim when mirroring
    int L;              // the halved axis
    uint32_t l_first;   // first index along L covered by the iteration
};

// out[idx] = conj(out[mirror(idx)]) for all idx with idx_L > n_L/2
// (reference: numba_c2c_sym, _pocketfft_numba.cpp:123-129, rev_iter H:3383-3444)
template <typename T>
__global__ void mirror_fill_kernel(NdIdx ix, uint64_t total, char *out) {
    for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < total; f += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t rem = f;
        int64_t dst = 0, src = 0;
        for (int d = ix.nd - 1; d >= 0; --d) {
            uint32_t i = (uint32_t)(rem % ix.ext[d]);
            rem /= ix.ext[d];
            if (d == ix.L) i += ix.l_first;
            uint32_t m = ix.rev[d] ? (i == 0 ? 0 : ix.full[d] - i) : i;
            dst += (int64_t)i * ix.sa[d];
            src += (int64_t)m * ix.sa[d];
        }
        const T *p = reinterpret_cast<const T *>(out + src);
        T re = p[0], im = p[1];
        T *q = reinterpret_cast<T *>(out + dst);
        q[0] = re;
        q[1] = -im;
    }
}

// out[idx] = Re + Im of tmp[idx] (idx_L <= n_L/2) or Re - Im of tmp[mirror(idx)]
// (reference: r2r_genuine_hartley H:4078-4090)
template <typename T>
__global__ void hartley_combine_kernel(NdIdx ix, uint64_t total, const char *tmp, char *out) {
    for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < total; f += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t rem = f;
        int64_t dst = 0, src = 0, srcm = 0;
        bool upper = false;
        for (int d = ix.nd - 1; d >= 0; --d) {
            uint32_t i = (uint32_t)(rem % ix.ext[d]);
            rem /= ix.ext[d];
            uint32_t m = ix.rev[d] ? (i == 0 ? 0 : ix.full[d] - i) : i;
            if (d == ix.L && 2 * i > ix.full[d]) upper = true;
            dst += (int64_t)i * ix.sb[d];
            src += (int64_t)i * ix.sa[d];
            srcm += (int64_t)m * ix.sa[d];
        }
        const T *p = reinterpret_cast<const T *>(tmp + (upper ? srcm : src));
        *reinterpret_cast<T *>(out + dst) = upper ? (p[0] - p[1]) : (p[0] + p[1]);
    }
}

static NdIdx make_ndidx(const std::vector<int64_t> &shape, const std::vector<int64_t> &sa,
                        const std::vector<int64_t> &sb, const std::vector<uint64_t> &axes) {
    NdIdx ix;
    memset(&ix, 0, sizeof(ix));
    if (shape.size() > 8) { set_error("more than 8 array dimensions are not supported here"); throw Error(); }
    ix.nd = (int)shape.size();
    for (int d = 0; d < ix.nd; ++d) {
        ix.ext[d] = ix.full[d] = (uint32_t)shape[d];
        ix.sa[d] = sa[d];
        ix.sb[d] = sb.empty() ? 0 : sb[d];
    }
    for (auto a : axes) ix.rev[a] = 1;
    ix.L = (int)axes.back();
    return ix;
}

static unsigned ew_blocks(uint64_t total) { return (unsigned)std::min<uint64_t>((total + 255) / 256, 148 * 32); }

void op_c2c_sym(const NdArgs &a, bool forward, cudaStream_t s) {
    if (any_zero(a.shape) || a.axes.empty()) return;
    r2c_into(a.prec, a.shape, a.sin, a.sout, a.axes, a.in, a.out, forward, a.fct, s);
    NdIdx ix = make_ndidx(a.shape, a.sout, {}, a.axes);
    const uint32_t nL = (uint32_t)a.shape[ix.L];
    ix.l_first = nL / 2 + 1;
    if (ix.l_first >= nL) return;
    ix.ext[ix.L] = nL - ix.l_first;
    uint64_t total = 1;
    for (int d = 0; d < ix.nd; ++d) total *= ix.ext[d];
    if (a.prec) mirror_fill_kernel<double><<<ew_blocks(total), 256, 0, s>>>(ix, total, a.out);
    else mirror_fill_kernel<float><<<ew_blocks(total), 256, 0, s>>>(ix, total, a.out);
    count_launch();
    RFB_AFTER_LAUNCH2();
}

// ---------------------------------------------------------------------------------------
// Hartley  (reference: H:4042-4091)
// ---------------------------------------------------------------------------------------
void op_separable_hartley(const NdArgs &a, cudaStream_t s) {
    if (any_zero(a.shape)) return;
    bool first = true;
    for (auto axu : a.axes) {
        const size_t ax = (size_t)axu;
        LineJob j;
        j.prec = a.prec;
        j.n = (uint64_t)a.shape[ax];
        j.in = first ? a.in : a.out;
        j.out = a.out;
        const auto &si = first ? a.sin : a.sout;
        j.is = si[ax];
        j.os = a.sout[ax];
        j.batch = batch_dims(a.shape, si, a.sout, ax);
        j.fct = first ? a.fct : 1.0;
        j.load_mode = LD_REAL;
        j.store_mode = ST_HARTLEY;
        run_lines(j, s);
        first = false;
    }
}

void op_genuine_hartley(const NdArgs &a, cudaStream_t s) {
    if (any_zero(a.shape) || a.axes.empty()) return;
    if (a.axes.size() == 1) return op_separable_hartley(a, s);
    const int64_t esz = a.prec ? 16 : 8;
    const size_t L = (size_t)a.axes.back();
    std::vector<int64_t> tshape = a.shape;
    tshape[L] = a.shape[L] / 2 + 1;
    std::vector<int64_t> tst = c_strides(tshape, esz);
    Scratch tmp(prod(tshape) * (uint64_t)esz, s);
    r2c_into(a.prec, a.shape, a.sin, tst, a.axes, a.in, (char *)tmp.p, true, a.fct, s);
    NdIdx ix = make_ndidx(a.shape, tst, a.sout, a.axes);
    uint64_t total = prod(a.shape);
    if (a.prec) hartley_combine_kernel<double><<<ew_blocks(total), 256, 0, s>>>(ix, total, (const char *)tmp.p, a.out);
    else hartley_combine_kernel<float><<<ew_blocks(total), 256, 0, s>>>(ix, total, (const char *)tmp.p, a.out);
    count_launch();
    RFB_AFTER_LAUNCH2();
}

// ---------------------------------------------------------------------------------------
// r2r_fftpack  (reference: H:4025-4040 + ExecR2R H:3854-3873).  The reference executes the
// real plan in direction `forward` whatever `real2hermitian` says (H:3867), so:
//   (r2h, fwd) = (T,T): real -> halfcomplex, sign -         (T,F): hc -> real, then negate 2,4,..
//                (F,F): halfcomplex -> real, sign +         (F,T): negate 2,4,.. then real -> hc
// ---------------------------------------------------------------------------------------
void op_fftpack(const NdArgs &a, bool r2h, bool forward, cudaStream_t s) {
    if (any_zero(a.shape)) return;
    bool first = true;
    for (auto axu : a.axes) {
        const size_t ax = (size_t)axu;
        LineJob j;
        j.prec = a.prec;
        j.n = (uint64_t)a.shape[ax];
        j.in = first ? a.in : a.out;
        j.out = a.out;
        const auto &si = first ? a.sin : a.sout;
        j.is = si[ax];
        j.os = a.sout[ax];
        j.batch = batch_dims(a.shape, si, a.sout, ax);
        j.fct = first ? a.fct : 1.0;
        if (forward) {
            j.load_mode = LD_REAL;
            j.store_mode = ST_HC;
            j.backward = false;
            if (!r2h) j.flags |= FLAG_NEG_EVEN_IN;
        } else {
            j.load_mode = LD_HC;
            j.store_mode = ST_REAL;
            j.backward = true;
            if (r2h) j.flags |= FLAG_NEG_EVEN_OUT;
        }
        run_lines(j, s);
        first = false;
    }
}

// ---------------------------------------------------------------------------------------
// DCT / DST types 1-4  (reference: dct H:3891-3912, dst H:3914-3935, T_dct1 2918,
// T_dst1 2957, T_dcst23 2987, T_dcst4 3063).  Each line is embedded into a complex sequence
// of length Lf (2(N-1), 2(N+1) or 2N), transformed by the line engine, and read back with
// the type's phase factor:  see DESIGN.md "DCT/DST" for the identities.
// ---------------------------------------------------------------------------------------
struct DcstParams {
    uint32_t N, Lf;
    int type, cosine, ortho, quirk;
};

struct LinesIdx {
    int nd;
    uint32_t ext[8];
    FastDiv d[8];
    int64_t is[8], os[8];
};

__device__ __forceinline__ void lines_off(const LinesIdx &bi, uint32_t l, int64_t &oi, int64_t &oo) {
    oi = 0; oo = 0;
    for (int d = 0; d < bi.nd; ++d) {
        uint32_t q, r;
        fdivmod(l, bi.d[d], q, r);
        oi += (int64_t)r * bi.is[d];
        oo += (int64_t)r * bi.os[d];
        l = q;
    }
}

// exp(-i pi num / den) evaluated in double
__device__ __forceinline__ double2 cis_mpi(double num, double den) {
    double sn, cs;
    sincospi(num / den, &sn, &cs);
    return make_double2(cs, -sn);
}

template <typename T>
__global__ void dcst_pre_kernel(LinesIdx bi, uint32_t l0, uint32_t nlines, const char *in, int64_t sa, cx<T> *z, DcstParams p) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.Lf) return;
    const uint32_t N = p.N;
    const double SQ2 = 1.4142135623730951;
    for (uint32_t l = blockIdx.y; l < nlines; l += gridDim.y) {  // lines l0 .. l0 + nlines of the array, slab rows 0 .. nlines
        int64_t oi, oo;
        lines_off(bi, l0 + l, oi, oo);
        auto x = [&](uint32_t i) { return (double)*reinterpret_cast<const T *>(in + oi + (int64_t)i * sa); };
        double re = 0.0, im = 0.0;
        if (p.type == 1) {
            if (p.cosine) {
                const uint32_t i = (e <= N - 1) ? e : p.Lf - e;
                re = x(i);
                if (p.ortho && (i == 0 || i == N - 1)) re *= SQ2;
            } else {
                if (e == 0 || e == N + 1) re = 0.0;
                else if (e <= N) re = x(e - 1);
                else re = -x(p.Lf - e - 1);
            }
        } else if (e < N) {
            // sine variants read the line reversed (types 3, 4) or with alternating sign (type 2)
            if (p.type == 2) {
                re = x(e);
                if (!p.cosine && (e & 1)) re = -re;
            } else {
                const uint32_t i = p.cosine ? e : N - 1 - e;
                double v = x(i);
                if (p.type == 3) {
                    // coefficient 1 for the transform's element 0, 2 otherwise; ortho scales one
                    // element by sqrt(2): index 0 of the line as the caller sees it for the cosine
                    // transform -- and, reproducing the reference, for the sine transform as well
                    // (SciPy scales the caller's index N-1 there)
                    if (e != 0) v *= 2.0;
                    if (p.ortho) {
                        const uint32_t scaled = p.cosine ? 0u : (p.quirk ? 0u : N - 1);
                        if (i == scaled) v *= SQ2;
                    }
                }
                const double2 w = cis_mpi((double)e, 2.0 * (double)N);
                re = v * w.x;
                im = v * w.y;
            }
        }
        z[(size_t)l * p.Lf + e] = mk<T>((T)re, (T)im);
    }
}

template <typename T>
__global__ void dcst_post_kernel(LinesIdx bi, uint32_t l0, uint32_t nlines, const cx<T> *z, char *out, int64_t sa, DcstParams p) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t N = p.N;
    if (k >= N) return;
    const double RSQ2 = 0.70710678118654752;
    for (uint32_t l = blockIdx.y; l < nlines; l += gridDim.y) {
        int64_t oi, oo;
        lines_off(bi, l0 + l, oi, oo);
        const cx<T> *zl = z + (size_t)l * p.Lf;
        double y;
        if (p.type == 1) {
            if (p.cosine) {
                y = (double)zl[k].x;
                if (p.ortho && (k == 0 || k == N - 1)) y *= RSQ2;
            } else y = -(double)zl[k + 1].y;
        } else if (p.type == 2) {
            const uint32_t kk = p.cosine ? k : N - 1 - k;
            const double2 w = cis_mpi((double)kk, 2.0 * (double)N);
            const cx<T> v = zl[kk];
            y = 2.0 * ((double)v.x * w.x - (double)v.y * w.y);
            if (p.ortho) {
                const uint32_t scaled = p.cosine ? 0u : (p.quirk ? 0u : N - 1);
                if (k == scaled) y *= RSQ2;
            }
        } else if (p.type == 3) {
            y = (double)zl[k].x;
            if (!p.cosine && (k & 1)) y = -y;
        } else {
            const double2 w = cis_mpi((double)(2 * k + 1), 4.0 * (double)N);
            const cx<T> v = zl[k];
            y = 2.0 * ((double)v.x * w.x - (double)v.y * w.y);
            if (!p.cosine && (k & 1)) y = -y;
        }
        *reinterpret_cast<T *>(out + oo + (int64_t)k * sa) = (T)y;
    }
}

void op_dcst(const NdArgs &a, int type, bool ortho, bool cosine, cudaStream_t s, int quirk_mode) {
    // quirk_mode: -1 = the process-wide setting (rfb200_set_dst_ortho_quirk), 0 = SciPy's scaling, 1 = the reference's
    const bool dst_quirk = quirk_mode < 0 ? g_dst_quirk : quirk_mode != 0;
    if (any_zero(a.shape)) return;
    if (type < 1 || type > 4) { set_error("invalid DCT/DST type"); throw Error(); }
    const int64_t esz = a.prec ? 16 : 8;
    bool first = true;
    for (auto axu : a.axes) {
        const size_t ax = (size_t)axu;
        const uint64_t N = (uint64_t)a.shape[ax];
        DcstParams p;
        p.N = (uint32_t)N;
        p.type = type;
        p.cosine = cosine ? 1 : 0;
        p.ortho = ortho ? 1 : 0;
        p.quirk = dst_quirk ? 1 : 0;
        if (type == 1) {
            if (cosine && N < 2) { set_error("DCT-I needs at least two points (zero-length FFT requested)"); throw Error(); }
            p.Lf = (uint32_t)(cosine ? 2 * (N - 1) : 2 * (N + 1));
        } else p.Lf = (uint32_t)(2 * N);
        const auto &si = first ? a.sin : a.sout;
        std::vector<Dim> b = batch_dims(a.shape, si, a.sout, ax);
        if ((type == 2 || type == 3) && N % 2 == 0) {
            // power-of-two half length: one fused kernel (reorder + packed real FFT + phase factors)
            LineJob fj;
            fj.prec = a.prec;
            fj.n = N;
            fj.in = first ? a.in : a.out;
            fj.out = a.out;
            fj.is = si[ax];
            fj.os = a.sout[ax];
            fj.batch = b;
            fj.fct = first ? a.fct : 1.0;
            fj.load_mode = type == 2 ? LD_DCT2 : LD_DCT3;
            fj.store_mode = type == 2 ? ST_DCT2 : ST_DCT3;
            fj.flags = (cosine ? 0 : FLAG_SINE) | (ortho ? FLAG_ORTHO : 0) | (dst_quirk ? FLAG_QUIRK : 0);
            if (run_lines_pow2(fj, s)) { first = false; continue; }
        }
        {
            // Any other type / length: the type's reordering, extension and phase factors are the load and store stage of ONE
            // complex transform in the shared-memory tile kernel (line_io.cuh LD_G_* / ST_G_*) -- one launch per axis, no
            // work area.  Transform length: 2(N-1) / 2(N+1) (type I), N (types II, III), N/2 (type IV, N even) or 2N (odd N).
            LineJob gj;
            gj.prec = a.prec;
            gj.in = first ? a.in : a.out;
            gj.out = a.out;
            gj.is = si[ax];
            gj.os = a.sout[ax];
            gj.batch = b;
            gj.fct = first ? a.fct : 1.0;
            gj.flags = (cosine ? 0 : FLAG_SINE) | (ortho ? FLAG_ORTHO : 0) | (dst_quirk ? FLAG_QUIRK : 0);
            bool have = true;
            if (type == 1) {
                gj.n = cosine ? 2 * (N - 1) : 2 * (N + 1);
                gj.load_mode = cosine ? LD_G_DCT1 : LD_G_DST1;
                gj.store_mode = cosine ? ST_G_DCT1 : ST_G_DST1;
                gj.flags &= ~FLAG_SINE;
            } else if (type == 2) {
                gj.n = N;
                gj.load_mode = LD_G_DCT2;
                gj.store_mode = ST_G_DCT2;
                gj.aux_st = get_table(TAB_QUARTER, a.prec, N, 0);
            } else if (type == 3) {
                gj.n = N;
                gj.load_mode = LD_G_DCT3;
                gj.store_mode = ST_G_DCT3;
                gj.backward = true;
                gj.aux_ld = get_table(TAB_QUARTER, a.prec, N, 0);
            } else if (N % 2 == 0) {
                gj.n = N / 2;
                gj.load_mode = LD_G_DCT4;
                gj.store_mode = ST_G_DCT4;
                gj.aux_ld = gj.aux_st = get_table(TAB_QUARTER, a.prec, 2 * N, 0);
            } else {
                gj.n = 2 * N;
                gj.load_mode = LD_G_DCT4Z;
                gj.store_mode = ST_G_DCT4Z;
                gj.aux_ld = get_table(TAB_QUARTER, a.prec, N, 0);
                gj.aux_st = get_table(TAB_QUARTER, a.prec, 2 * N, 0);
            }
            if (have && gj.n >= 1 && run_lines_tile(gj, s)) { first = false; continue; }
        }
        // ---- lines too long for a shared-memory tile: embedding in a work area, three launches, in slabs of lines ----------
        LinesIdx bi;
        memset(&bi, 0, sizeof(bi));
        uint64_t nlines = 1;
        {
            int k = 0;
            for (auto &d : b) {
                if (d.n == 1) continue;
                if (k >= 8) { set_error("more than 9 array dimensions are not supported"); throw Error(); }
                bi.ext[k] = (uint32_t)d.n;
                bi.d[k] = make_fastdiv((uint32_t)d.n);
                bi.is[k] = d.is;
                bi.os[k] = d.os;
                nlines *= (uint64_t)d.n;
                ++k;
            }
            bi.nd = k;
        }
        if (nlines >= (1ull << 31)) { set_error("too many lines"); throw Error(); }
        // bounded work area: the lines go through in slabs of <= 1 GiB of scratch
        const uint64_t per_line = (uint64_t)p.Lf * (uint64_t)esz;
        const uint64_t slab = std::max<uint64_t>(1, std::min<uint64_t>(nlines, (1ull << 30) / per_line));
        Scratch sc(slab * per_line, s);
        const char *src = first ? a.in : a.out;
        for (uint64_t l0 = 0; l0 < nlines; l0 += slab) {
            const uint64_t cnt = std::min<uint64_t>(slab, nlines - l0);
            dim3 grid((unsigned)((p.Lf + 255) / 256), (unsigned)std::min<uint64_t>(cnt, 32768));
            if (a.prec) dcst_pre_kernel<double><<<grid, 256, 0, s>>>(bi, (uint32_t)l0, (uint32_t)cnt, src, si[ax], (double2 *)sc.p, p);
            else dcst_pre_kernel<float><<<grid, 256, 0, s>>>(bi, (uint32_t)l0, (uint32_t)cnt, src, si[ax], (float2 *)sc.p, p);
            count_launch("dcst_pre_kernel");
            RFB_AFTER_LAUNCH2();
            LineJob j;
            j.prec = a.prec;
            j.n = p.Lf;
            j.is = j.os = esz;
            j.batch.push_back(Dim{(int64_t)cnt, (int64_t)per_line, (int64_t)per_line, false});
            j.in = (const char *)sc.p;
            j.out = (char *)sc.p;
            j.fct = first ? a.fct : 1.0;
            run_lines(j, s);
            dim3 grid2((unsigned)((N + 255) / 256), (unsigned)std::min<uint64_t>(cnt, 32768));
            if (a.prec) dcst_post_kernel<double><<<grid2, 256, 0, s>>>(bi, (uint32_t)l0, (uint32_t)cnt, (const double2 *)sc.p, a.out, a.sout[ax], p);
            else dcst_post_kernel<float><<<grid2, 256, 0, s>>>(bi, (uint32_t)l0, (uint32_t)cnt, (const float2 *)sc.p, a.out, a.sout[ax], p);
            count_launch("dcst_post_kernel");
            RFB_AFTER_LAUNCH2();
        }
        first = false;
    }
}

}  // namespace rfb
