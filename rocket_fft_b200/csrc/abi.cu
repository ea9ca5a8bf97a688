// C ABI of librocketfft_b200.so: the ten numba_* drop-in symbols and the rfb200_* device entry
// points declared in include/rocketfft_b200.h.
#include <math.h>
#include <fcntl.h>
#include <pthread.h>
#include <unistd.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/rocketfft_b200.h"
#include "engine.h"
#include "tile_kernel.cuh"

using namespace rfb;

#define RFB_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

// ---- library stream (per device) ----------------------------------------------------------
std::mutex g_stream_mu;
cudaStream_t g_lib_stream[64] = {nullptr};
cudaStream_t g_pipe_stream[64][3] = {{nullptr}};  // host-array pipeline: H2D / kernels / D2H of different chunks overlap
thread_local bool g_user_stream_set = false;
thread_local cudaStream_t g_user_stream = nullptr;

cudaStream_t lib_stream() {
    if (g_user_stream_set) return g_user_stream;
    int dev = 0;
    RFB_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_stream_mu);
    if (!g_lib_stream[dev & 63]) RFB_CUDA_CHECK(cudaStreamCreateWithFlags(&g_lib_stream[dev & 63], cudaStreamNonBlocking));
    return g_lib_stream[dev & 63];
}

// ---- process state: fork detection, failures of the void numba_* entry points --------------------------------------
// A CUDA context does not survive fork() (the reference restarts its thread pool in the child, _pocketfft_hdronly.h:992-1006;
// there is no such remedy for a GPU context): a child forked AFTER this library touched CUDA gets a clear error from every
// transform instead of undefined behaviour.  A child forked before the first transform initialises CUDA itself and works.
std::atomic<bool> g_cuda_used{false}, g_forked_child{false};
std::atomic<uint64_t> g_failures{0};
void on_fork_child() {
    if (g_cuda_used.load()) {
        g_forked_child.store(true);
        plan_cache_forget();
    }
}
struct ForkInit {
    ForkInit() { pthread_atfork(nullptr, nullptr, on_fork_child); }
} g_fork_init;

// start of every transform entry point
void enter_call() {
    if (g_forked_child.load()) {
        set_error("this process was forked after the parent had used CUDA, and a CUDA context does not survive fork(): start worker "
                  "processes with the 'spawn' or 'forkserver' method, or fork before the first transform");
        throw Error();
    }
    g_cuda_used.store(true);
    check_async_error();
}

// Runs on the device that owns the arrays (not whatever device is current), restores the caller's device afterwards.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    static int device_of(const void *p) {
        if (!p) return -1;
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return -1; }
        return (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) ? at.device : -1;
    }
    DeviceGuard(const void *a, const void *b) {
        const int da = device_of(a), db = device_of(b);
        if (da >= 0 && db >= 0 && da != db) { set_error("input and output arrays live on different devices"); throw Error(); }
        const int want = da >= 0 ? da : db;
        if (want < 0) return;
        RFB_CUDA_CHECK(cudaGetDevice(&prev));
        if (prev != want) {
            RFB_CUDA_CHECK(cudaSetDevice(want));
            switched = true;
        }
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

// A numba_* entry point returns void: a failed call must not leave `aout` looking like a result.  Host output arrays are
// filled with NaN (element by element: the gaps of a strided array are not ours), the message goes to stderr and stays
// readable through rfb200_last_error, and rfb200_failure_count lets a caller detect failures after the fact.
void fill_nan_host(const rfb200_array_record *aout, uint64_t ndim, const std::vector<int64_t> &shape, int64_t item, int64_t scalar) {
    if (!aout || !aout->data) return;
    for (auto v : shape) if (v <= 0) return;
    std::vector<int64_t> idx(ndim, 0);
    const int64_t *st = aout->shape_and_strides + ndim;
    for (;;) {
        int64_t off = 0;
        for (uint64_t d = 0; d < ndim; ++d) off += idx[d] * st[d];
        char *q = (char *)aout->data + off;
        for (int64_t c = 0; c < item / scalar; ++c) {
            if (scalar == 8) ((double *)q)[c] = NAN;
            else ((float *)q)[c] = NAN;
        }
        uint64_t d = ndim;
        while (d > 0) {
            --d;
            if (++idx[d] < shape[d]) break;
            idx[d] = 0;
            if (d == 0) return;
        }
        if (ndim == 0) return;
    }
}

// Is [p, p + n) ordinary writable host memory?  Asked without CUDA (the forked child must not call it) and without a
// signal handler: read(2) from /dev/zero into the range fails with EFAULT instead of faulting.  (Overwrites the range
// with zeros -- only used on an output that is about to be filled with NaN.)
bool host_writable(void *p, size_t n) {
    if (!p) return false;
    const int fd = open("/dev/zero", O_RDONLY);
    if (fd < 0) return false;
    const ssize_t r = read(fd, p, n);
    close(fd);
    return r == (ssize_t)n;
}

enum OpKind { OP_C2C, OP_R2C, OP_C2R, OP_C2C_SYM, OP_DCT, OP_DST, OP_FFTPACK, OP_SEP_HARTLEY, OP_GEN_HARTLEY };

struct OpFlags {
    bool forward = true, ortho = false, r2h = false;
    int type = 2;
    int quirk = -1;  // DST-II/III ortho scaling: -1 process default, 0 SciPy's, 1 the reference's
};

bool in_is_complex(OpKind k) { return k == OP_C2C || k == OP_C2R; }
bool out_is_complex(OpKind k) { return k == OP_C2C || k == OP_R2C || k == OP_C2C_SYM; }

void dispatch(OpKind k, const NdArgs &a, const OpFlags &f, cudaStream_t s) {
    TableScope tables;  // the plan tables this call fetches stay leased until its kernels are enqueued
    for (auto ax : a.axes)
        if (ax >= a.shape.size()) { set_error("axis out of range"); throw Error(); }
    switch (k) {
        case OP_C2C: op_c2c(a, f.forward, s); break;
        case OP_R2C: op_r2c(a, f.forward, s); break;
        case OP_C2R: op_c2r(a, f.forward, s); break;
        case OP_C2C_SYM: op_c2c_sym(a, f.forward, s); break;
        case OP_DCT: op_dcst(a, f.type, f.ortho, true, s); break;
        case OP_DST: op_dcst(a, f.type, f.ortho, false, s, f.quirk); break;
        case OP_FFTPACK: op_fftpack(a, f.r2h, f.forward, s); break;
        case OP_SEP_HARTLEY: op_separable_hartley(a, s); break;
        case OP_GEN_HARTLEY: op_genuine_hartley(a, s); break;
    }
}

// byte span [lo, hi) covered by an array relative to its data pointer
void span_of(const std::vector<int64_t> &shape, const std::vector<int64_t> &st, int64_t itemsize, int64_t &lo, int64_t &hi) {
    lo = 0;
    hi = 0;
    for (size_t d = 0; d < shape.size(); ++d) {
        const int64_t ext = (shape[d] - 1) * st[d];
        if (ext < 0) lo += ext; else hi += ext;
    }
    hi += itemsize;
}

bool is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// One staging engine per device, created on first use with pageable memory; a call that finds it busy (another thread is
// inside a host-array call on the same device) lets the driver stage its copies instead.
std::mutex g_stager_mu[64];
Stager *g_stager[64] = {nullptr};

struct DevBuf {
    void *p = nullptr;
    cudaStream_t s;
    DevBuf(size_t n, cudaStream_t st) : s(st) { RFB_CUDA_CHECK(cudaMallocAsync(&p, n ? n : 16, st)); }
    ~DevBuf() { if (p) cudaFreeAsync(p, s); }
};

// The numba_* path: arrays described by Numba array records; host data is staged.
void run_record_op(OpKind k, uint64_t ndim, const rfb200_array_record *ain, rfb200_array_record *aout,
                   const rfb200_array_record *axes, double fct, const OpFlags &f) {
    clear_error();
    std::vector<int64_t> fail_shape;  // output extents, known once the arguments are parsed (for the NaN fill on failure)
    int64_t fail_item = 0, fail_scalar = 4;
    bool out_on_host = false;
    try {
        NdArgs a;
        const rfb200_array_record *shp_src = (k == OP_C2R) ? aout : ain;
        a.shape.assign(shp_src->shape_and_strides, shp_src->shape_and_strides + ndim);
        a.sin.assign(ain->shape_and_strides + ndim, ain->shape_and_strides + 2 * ndim);
        a.sout.assign(aout->shape_and_strides + ndim, aout->shape_and_strides + 2 * ndim);
        const uint64_t *ax = reinterpret_cast<const uint64_t *>(axes->data);
        a.axes.assign(ax, ax + axes->nitems);
        a.fct = fct;
        for (auto s : a.shape)
            if (s == 0) return;
        if (a.axes.empty()) return;
        // precision from the input's item size (reference: _pocketfft_numba.cpp:38, 58, 98, 152)
        const int64_t isz = ain->itemsize;
        const int64_t in_scalar = in_is_complex(k) ? isz / 2 : isz;
        a.prec = (in_scalar == 8) ? 1 : 0;
        const int64_t ssz = a.prec ? 8 : 4;
        const int64_t in_item = in_is_complex(k) ? 2 * ssz : ssz;
        const int64_t out_item = out_is_complex(k) ? 2 * ssz : ssz;
        const size_t L = (size_t)a.axes.back();
        std::vector<int64_t> shape_in = a.shape, shape_out = a.shape;
        if (k == OP_R2C) shape_out[L] = a.shape[L] / 2 + 1;
        if (k == OP_C2R) shape_in[L] = a.shape[L] / 2 + 1;

        fail_shape = shape_out;
        fail_item = out_item;
        fail_scalar = ssz;
        if (g_forked_child.load()) out_on_host = host_writable(aout->data, (size_t)ssz);  // (no CUDA query in the child)
        enter_call();
        const bool dev_in = is_device_ptr(ain->data), dev_out = is_device_ptr(aout->data);
        out_on_host = !dev_out;
        if (dev_in != dev_out) { set_error("input and output must both be host or both be device arrays"); throw Error(); }
        if (dev_in) {
            // Device (or managed) memory behind a numba_* call: the reference's entry points are synchronous, so is this
            // one.  Unless the caller chose a stream (rfb200_set_stream), the work goes to the legacy default stream, which
            // orders it after the kernels the caller has queued on blocking streams, and the call returns when it is done.
            DeviceGuard dg(ain->data, aout->data);
            cudaStream_t ds = g_user_stream_set ? g_user_stream : (cudaStream_t) nullptr;
            a.in = (const char *)ain->data;
            a.out = (char *)aout->data;
            dispatch(k, a, f, ds);
            RFB_CUDA_CHECK(cudaStreamSynchronize(ds));
            check_async_error();
            return;
        }
        cudaStream_t s = lib_stream();
        // ---- host arrays: H2D -> kernels -> D2H, synchronous for the caller ----
        // Batched work (an untransformed outer dim whose slices are disjoint in memory) is cut into
        // chunks that go round-robin over three streams, so that the H2D copy of one chunk, the kernels
        // of another and the D2H copy of a third overlap (PCIe is full duplex; the kernels are ~30x
        // faster than the link).
        size_t cd = a.shape.size();
        if (!g_user_stream_set) {
            int64_t best = 0;
            for (size_t d = 0; d < a.shape.size(); ++d) {
                bool is_axis = false;
                for (auto ax2 : a.axes) is_axis = is_axis || (ax2 == d);
                if (is_axis || a.shape[d] < 2 || a.sin[d] <= 0 || a.sout[d] <= 0) continue;
                std::vector<int64_t> si = shape_in, so = shape_out;
                si[d] = 1;
                so[d] = 1;
                int64_t l1, h1, l2, h2;
                span_of(si, a.sin, in_item, l1, h1);
                span_of(so, a.sout, out_item, l2, h2);
                if (h1 - l1 > a.sin[d] || h2 - l2 > a.sout[d]) continue;  // slices interleave in memory
                if (a.sin[d] > best) { best = a.sin[d]; cd = d; }
            }
        }
        int64_t tlo, thi, ulo, uhi;
        span_of(shape_in, a.sin, in_item, tlo, thi);
        span_of(shape_out, a.sout, out_item, ulo, uhi);
        const uint64_t total_bytes = (uint64_t)(thi - tlo) + (uint64_t)(uhi - ulo);
        int64_t nchunks = 1;
        if (cd < a.shape.size() && total_bytes >= (64ull << 20)) {
            nchunks = std::min<int64_t>(a.shape[cd], 8);
            while (nchunks > 1 && total_bytes / (uint64_t)nchunks < (16ull << 20)) --nchunks;
        }
        cudaStream_t streams[3] = {s, s, s};
        if (nchunks > 1) {
            int dev = 0;
            RFB_CUDA_CHECK(cudaGetDevice(&dev));
            std::lock_guard<std::mutex> lk(g_stream_mu);
            for (int i = 0; i < 3; ++i) {
                cudaStream_t &st = g_pipe_stream[dev & 63][i];
                if (!st) RFB_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
                streams[i] = st;
            }
        }
        const char *hin0 = (const char *)ain->data;
        char *hout0 = (char *)aout->data;
        // pageable host memory (a plain NumPy array): our own staging ring instead of the driver's (staging.cu)
        Stager *stg = nullptr;
        std::unique_lock<std::mutex> stg_lock;
        if (total_bytes >= (8ull << 20) && (host_memory_is_pageable(hin0) || host_memory_is_pageable(hout0))) {
            int dev = 0;
            RFB_CUDA_CHECK(cudaGetDevice(&dev));
            stg_lock = std::unique_lock<std::mutex>(g_stager_mu[dev & 63], std::try_to_lock);
            if (stg_lock.owns_lock()) {
                if (!g_stager[dev & 63]) g_stager[dev & 63] = new Stager();
                stg = g_stager[dev & 63];
            }
        }
        const bool stage_in = stg && host_memory_is_pageable(hin0), stage_out = stg && host_memory_is_pageable(hout0);
        std::vector<DevBuf *> bufs;
        struct G { std::vector<DevBuf *> &v; ~G() { for (auto p : v) delete p; } } guard{bufs};
        const int64_t ext = cd < a.shape.size() ? a.shape[cd] : 1;
        const int64_t per = (ext + nchunks - 1) / nchunks;
        for (int64_t c0 = 0, ci = 0; c0 < ext; c0 += per, ++ci) {
            cudaStream_t cs = streams[ci % 3];
            NdArgs sub = a;
            std::vector<int64_t> si = shape_in, so = shape_out;
            const char *hin_c = hin0;
            char *hout_c = hout0;
            if (nchunks > 1) {
                const int64_t cnt = std::min<int64_t>(per, ext - c0);
                sub.shape[cd] = cnt;
                si[cd] = cnt;
                so[cd] = cnt;
                hin_c += c0 * a.sin[cd];
                hout_c += c0 * a.sout[cd];
            }
            int64_t ilo, ihi, olo, ohi;
            span_of(si, a.sin, in_item, ilo, ihi);
            span_of(so, a.sout, out_item, olo, ohi);
            const char *hin = hin_c + ilo;
            char *hout = hout_c + olo;
            const size_t ibytes = (size_t)(ihi - ilo), obytes = (size_t)(ohi - olo);
            const bool same = (hin == hout) && (ibytes == obytes);
            DevBuf *din = new DevBuf(ibytes, cs);
            bufs.push_back(din);
            if (stage_in) stg->upload((char *)din->p, hin, ibytes, cs);
            else RFB_CUDA_CHECK(cudaMemcpyAsync(din->p, hin, ibytes, cudaMemcpyHostToDevice, cs));
            char *dout_base;
            if (same) dout_base = (char *)din->p;
            else {
                DevBuf *dout = new DevBuf(obytes, cs);
                bufs.push_back(dout);
                dout_base = (char *)dout->p;
                uint64_t dense = (uint64_t)out_item;
                for (auto v : so) dense *= (uint64_t)v;
                if (dense != obytes) {  // gaps between elements must survive the round trip
                    if (stage_out) stg->upload(dout_base, hout, obytes, cs);
                    else RFB_CUDA_CHECK(cudaMemcpyAsync(dout_base, hout, obytes, cudaMemcpyHostToDevice, cs));
                }
            }
            sub.in = (const char *)din->p - ilo;
            sub.out = dout_base - olo;
            dispatch(k, sub, f, cs);
            if (stage_out) stg->download(hout, dout_base, obytes, cs);
            else RFB_CUDA_CHECK(cudaMemcpyAsync(hout, dout_base, obytes, cudaMemcpyDeviceToHost, cs));
        }
        for (int i = 0; i < (nchunks > 1 ? 3 : 1); ++i) RFB_CUDA_CHECK(cudaStreamSynchronize(streams[i]));
        if (stg) stg->finish();
        check_async_error();
        return;
    } catch (const Error &) {
    } catch (const std::exception &e) {
        set_error(e.what());
    }
    // ---- failure: never return silently with an untouched output -----------------------------------------------------
    g_failures.fetch_add(1);
    fprintf(stderr, "rocketfft_b200: transform failed: %s\n", last_error());
    if (out_on_host && fail_item) fill_nan_host(aout, ndim, fail_shape, fail_item, fail_scalar);
}

int run_device_op(OpKind k, int precision, size_t ndim, const int64_t *shape, const int64_t *stride_in,
                  const int64_t *stride_out, size_t naxes, const uint64_t *axes, double fct, const void *d_in,
                  void *d_out, void *stream, const OpFlags &f) {
    clear_error();
    try {
        enter_call();
        DeviceGuard dg(d_in, d_out);
        NdArgs a;
        a.prec = precision ? 1 : 0;
        a.shape.assign(shape, shape + ndim);
        a.sin.assign(stride_in, stride_in + ndim);
        a.sout.assign(stride_out, stride_out + ndim);
        a.axes.assign(axes, axes + naxes);
        a.in = (const char *)d_in;
        a.out = (char *)d_out;
        a.fct = fct;
        dispatch(k, a, f, (cudaStream_t)stream);
        return 0;
    } catch (const Error &) {
        return 1;
    } catch (const std::exception &e) {
        set_error(e.what());
        return 1;
    }
}

OpFlags mk_flags(bool forward, int type = 2, bool ortho = false, bool r2h = false) {
    OpFlags f;
    f.forward = forward;
    f.type = type;
    f.ortho = ortho;
    f.r2h = r2h;
    return f;
}

}  // namespace

// ---- (1) numba_* -------------------------------------------------------------------------------
RFB_EXPORT uint64_t numba_good_size(uint64_t target, bool real) { return rfb::good_size(target, real); }

RFB_EXPORT void numba_c2c(uint64_t ndim, const rfb200_array_record *ain, rfb200_array_record *aout,
                          rfb200_array_record *axes, bool forward, double fct, uint64_t) {
    run_record_op(OP_C2C, ndim, ain, aout, axes, fct, mk_flags(forward));
}
RFB_EXPORT void numba_r2c(uint64_t ndim, const rfb200_array_record *ain, rfb200_array_record *aout,
                          rfb200_array_record *axes, bool forward, double fct, uint64_t) {
    run_record_op(OP_R2C, ndim, ain, aout, axes, fct, mk_flags(forward));
}
RFB_EXPORT void numba_c2r(uint64_t ndim, const rfb200_array_record *ain, rfb200_array_record *aout,
                          rfb200_array_record *axes, bool forward, double fct, uint64_t) {
    run_record_op(OP_C2R, ndim, ain, aout, axes, fct, mk_flags(forward));
}
RFB_EXPORT void numba_c2c_sym(uint64_t ndim, const rfb200_array_record *ain, rfb200_array_record *aout,
                              rfb200_array_record *axes, bool forward, double fct, uint64_t) {
    run_record_op(OP_C2C_SYM, ndim, ain, aout, axes, fct, mk_flags(forward));
}
RFB_EXPORT void numba_dct(uint64_t ndim, const rfb200_array_record *ain, rfb200_array_record *aout,
                          rfb200_array_record *axes, uint64_t type, double fct, bool ortho, uint64_t) {
    run_record_op(OP_DCT, ndim, ain, aout, axes, fct, mk_flags(true, (int)type, ortho));
}
RFB_EXPORT void numba_dst(uint64_t ndim, const rfb200_array_record *ain, rfb200_array_record *aout,
                          rfb200_array_record *axes, uint64_t type, double fct, bool ortho, uint64_t) {
    run_record_op(OP_DST, ndim, ain, aout, axes, fct, mk_flags(true, (int)type, ortho));
}
RFB_EXPORT void numba_r2r_fftpack(uint64_t ndim, const rfb200_array_record *ain, rfb200_array_record *aout,
                                  rfb200_array_record *axes, bool real2hermitian, bool forward, double fct, uint64_t) {
    run_record_op(OP_FFTPACK, ndim, ain, aout, axes, fct, mk_flags(forward, 2, false, real2hermitian));
}
RFB_EXPORT void numba_r2r_separable_hartley(uint64_t ndim, const rfb200_array_record *ain, rfb200_array_record *aout,
                                            rfb200_array_record *axes, double fct, uint64_t) {
    run_record_op(OP_SEP_HARTLEY, ndim, ain, aout, axes, fct, mk_flags(true));
}
RFB_EXPORT void numba_r2r_genuine_hartley(uint64_t ndim, const rfb200_array_record *ain, rfb200_array_record *aout,
                                          rfb200_array_record *axes, double fct, uint64_t) {
    run_record_op(OP_GEN_HARTLEY, ndim, ain, aout, axes, fct, mk_flags(true));
}

// ---- (2) rfb200_* --------------------------------------------------------------------------------
#define DEV_ARGS                                                                                               \
    int precision, size_t ndim, const int64_t *shape, const int64_t *stride_in, const int64_t *stride_out,    \
        size_t naxes, const uint64_t *axes
#define DEV_PASS precision, ndim, shape, stride_in, stride_out, naxes, axes

RFB_EXPORT int rfb200_c2c(DEV_ARGS, int forward, double fct, const void *d_in, void *d_out, void *stream) {
    return run_device_op(OP_C2C, DEV_PASS, fct, d_in, d_out, stream, mk_flags(forward != 0));
}
RFB_EXPORT int rfb200_r2c(DEV_ARGS, int forward, double fct, const void *d_in, void *d_out, void *stream) {
    return run_device_op(OP_R2C, DEV_PASS, fct, d_in, d_out, stream, mk_flags(forward != 0));
}
RFB_EXPORT int rfb200_c2r(DEV_ARGS, int forward, double fct, const void *d_in, void *d_out, void *stream) {
    return run_device_op(OP_C2R, DEV_PASS, fct, d_in, d_out, stream, mk_flags(forward != 0));
}
RFB_EXPORT int rfb200_c2c_sym(DEV_ARGS, int forward, double fct, const void *d_in, void *d_out, void *stream) {
    return run_device_op(OP_C2C_SYM, DEV_PASS, fct, d_in, d_out, stream, mk_flags(forward != 0));
}
RFB_EXPORT int rfb200_dct(DEV_ARGS, int type, double fct, int ortho, const void *d_in, void *d_out, void *stream) {
    return run_device_op(OP_DCT, DEV_PASS, fct, d_in, d_out, stream, mk_flags(true, type, ortho != 0));
}
RFB_EXPORT int rfb200_dst(DEV_ARGS, int type, double fct, int ortho, const void *d_in, void *d_out, void *stream) {
    // ortho: 0 = off, 1 = on with the process-wide DST-II/III scaling choice (default: the reference's quirk),
    // 2 = on with SciPy's scaling, 3 = on with the reference's, whatever the process-wide setting is
    OpFlags f = mk_flags(true, type, ortho != 0);
    if (ortho == 2) f.quirk = 0;
    if (ortho == 3) f.quirk = 1;
    return run_device_op(OP_DST, DEV_PASS, fct, d_in, d_out, stream, f);
}
RFB_EXPORT int rfb200_r2r_fftpack(DEV_ARGS, int real2hermitian, int forward, double fct, const void *d_in,
                                  void *d_out, void *stream) {
    return run_device_op(OP_FFTPACK, DEV_PASS, fct, d_in, d_out, stream, mk_flags(forward != 0, 2, false, real2hermitian != 0));
}
RFB_EXPORT int rfb200_r2r_separable_hartley(DEV_ARGS, double fct, const void *d_in, void *d_out, void *stream) {
    return run_device_op(OP_SEP_HARTLEY, DEV_PASS, fct, d_in, d_out, stream, mk_flags(true));
}
RFB_EXPORT int rfb200_r2r_genuine_hartley(DEV_ARGS, double fct, const void *d_in, void *d_out, void *stream) {
    return run_device_op(OP_GEN_HARTLEY, DEV_PASS, fct, d_in, d_out, stream, mk_flags(true));
}

RFB_EXPORT int rfb200_c2c_scatter(int precision, size_t ndim, const int64_t *shape, const int64_t *stride_in,
                                  const int64_t *stride_out, size_t axis, int forward, double fct, const void *d_in,
                                  size_t nparts, void *const *d_out_parts, void *stream) {
    clear_error();
    try {
        enter_call();
        DeviceGuard dg(d_in, nullptr);  // (the parts may be peer-mapped memory of other devices)
        TableScope tables;
        NdArgs a;
        a.prec = precision ? 1 : 0;
        a.shape.assign(shape, shape + ndim);
        a.sin.assign(stride_in, stride_in + ndim);
        a.sout.assign(stride_out, stride_out + ndim);
        a.in = (const char *)d_in;
        a.out = nullptr;
        a.fct = fct;
        if (axis >= ndim) { set_error("axis out of range"); throw Error(); }
        std::vector<char *> parts;
        for (size_t i = 0; i < nparts; ++i) parts.push_back((char *)d_out_parts[i]);
        op_c2c_scatter(a, axis, forward != 0, parts, (cudaStream_t)stream);
        return 0;
    } catch (const Error &) {
        return 1;
    } catch (const std::exception &e) {
        set_error(e.what());
        return 1;
    }
}

// ---- zero-padded / cropped input and index rotation (the layer directly above the path) ---------------
static int run_pad_op(int kind, int precision, size_t ndim, const int64_t *shape_in, const int64_t *shape,
                      const int64_t *stride_in, const int64_t *stride_out, size_t naxes, const uint64_t *axes, int forward,
                      double fct, const void *d_in, void *d_out, void *stream) {
    clear_error();
    try {
        enter_call();
        DeviceGuard dg(d_in, d_out);
        TableScope tables;
        NdArgs a;
        a.prec = precision ? 1 : 0;
        a.shape.assign(shape, shape + ndim);
        a.sin.assign(stride_in, stride_in + ndim);
        a.sout.assign(stride_out, stride_out + ndim);
        a.axes.assign(axes, axes + naxes);
        a.in = (const char *)d_in;
        a.out = (char *)d_out;
        a.fct = fct;
        for (auto ax : a.axes)
            if (ax >= ndim) { set_error("axis out of range"); throw Error(); }
        const std::vector<int64_t> sin_shape(shape_in, shape_in + ndim);
        if (kind == 0) op_c2c_pad(a, sin_shape, forward != 0, (cudaStream_t)stream);
        else if (kind == 1) op_r2c_pad(a, sin_shape, forward != 0, (cudaStream_t)stream);
        else op_c2r_pad(a, sin_shape, forward != 0, (cudaStream_t)stream);
        return 0;
    } catch (const Error &) {
        return 1;
    } catch (const std::exception &e) {
        set_error(e.what());
        return 1;
    }
}

#define PAD_ARGS                                                                                                   \
    int precision, size_t ndim, const int64_t *shape_in, const int64_t *shape, const int64_t *stride_in,           \
        const int64_t *stride_out, size_t naxes, const uint64_t *axes, int forward, double fct, const void *d_in, \
        void *d_out, void *stream
#define PAD_PASS precision, ndim, shape_in, shape, stride_in, stride_out, naxes, axes, forward, fct, d_in, d_out, stream
RFB_EXPORT int rfb200_c2c_pad(PAD_ARGS) { return run_pad_op(0, PAD_PASS); }
RFB_EXPORT int rfb200_r2c_pad(PAD_ARGS) { return run_pad_op(1, PAD_PASS); }
RFB_EXPORT int rfb200_c2r_pad(PAD_ARGS) { return run_pad_op(2, PAD_PASS); }

RFB_EXPORT int rfb200_roll(int itemsize, size_t ndim, const int64_t *shape, const int64_t *stride_in,
                           const int64_t *stride_out, const int64_t *shift, const void *d_in, void *d_out, void *stream) {
    clear_error();
    try {
        enter_call();
        DeviceGuard dg(d_in, d_out);
        op_roll(itemsize, std::vector<int64_t>(shape, shape + ndim), std::vector<int64_t>(stride_in, stride_in + ndim),
                std::vector<int64_t>(stride_out, stride_out + ndim), std::vector<int64_t>(shift, shift + ndim),
                (const char *)d_in, (char *)d_out, (cudaStream_t)stream);
        return 0;
    } catch (const Error &) {
        return 1;
    } catch (const std::exception &e) {
        set_error(e.what());
        return 1;
    }
}

RFB_EXPORT int rfb200_scale_lines(int precision, int complex_items, uint64_t nlines, uint64_t n, const void *d_table,
                                  void *d_data, void *stream) {
    clear_error();
    try {
        enter_call();
        DeviceGuard dg(d_table, d_data);
        op_scale_lines(precision ? 1 : 0, complex_items != 0, nlines, n, d_table, d_data, (cudaStream_t)stream);
        return 0;
    } catch (const Error &) {
        return 1;
    } catch (const std::exception &e) {
        set_error(e.what());
        return 1;
    }
}

// ---- housekeeping ---------------------------------------------------------------------------------
RFB_EXPORT const char *rfb200_last_error(void) { return rfb::last_error(); }
RFB_EXPORT void rfb200_clear_error(void) { rfb::clear_error(); }
RFB_EXPORT void rfb200_plan_cache_clear(void) { rfb::plan_cache_clear(); }
RFB_EXPORT void rfb200_set_stream(void *stream) {
    g_user_stream_set = true;
    g_user_stream = (cudaStream_t)stream;
}
RFB_EXPORT void rfb200_use_library_stream(void) { g_user_stream_set = false; }
RFB_EXPORT uint64_t rfb200_launch_count(void) { return rfb::launch_count(); }
RFB_EXPORT void rfb200_launch_trace(int enable) { rfb::launch_trace_enable(enable != 0); }
RFB_EXPORT const char *rfb200_launch_trace_get(void) {
    static thread_local std::string buf;
    buf = rfb::launch_trace_get();
    return buf.c_str();
}
RFB_EXPORT void rfb200_launch_count_reset(void) { rfb::launch_count_reset(); }
RFB_EXPORT void rfb200_set_dst_ortho_quirk(int enabled) { rfb::set_dst_ortho_quirk(enabled != 0); }
RFB_EXPORT uint64_t rfb200_failure_count(void) { return g_failures.load(); }
RFB_EXPORT int rfb200_host_pin(void *ptr, uint64_t bytes) {
    clear_error();
    try {
        enter_call();
        if (!ptr || !bytes) { set_error("rfb200_host_pin: empty range"); throw Error(); }
        RFB_CUDA_CHECK(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
        return 0;
    } catch (const Error &) {
    } catch (const std::exception &e) { set_error(e.what()); }
    cudaGetLastError();
    return 1;
}
RFB_EXPORT int rfb200_host_unpin(void *ptr) {
    clear_error();
    try {
        enter_call();
        RFB_CUDA_CHECK(cudaHostUnregister(ptr));
        return 0;
    } catch (const Error &) {
    } catch (const std::exception &e) { set_error(e.what()); }
    cudaGetLastError();
    return 1;
}
RFB_EXPORT void rfb200_plan_cache_stats(uint64_t *entries, uint64_t *bytes) { rfb::plan_cache_stats(entries, bytes); }
// host-array variant of numba_dst with an explicit DST-II/III ortho scaling (quirk: 0 SciPy's, 1 the reference's)
RFB_EXPORT void rfb200_host_dst(uint64_t ndim, const rfb200_array_record *ain, rfb200_array_record *aout, rfb200_array_record *axes,
                                uint64_t type, double fct, int ortho, int quirk) {
    OpFlags f = mk_flags(true, (int)type, ortho != 0);
    f.quirk = quirk ? 1 : 0;
    run_record_op(OP_DST, ndim, ain, aout, axes, fct, f);
}
RFB_EXPORT int64_t rfb200_debug_fuse4_unit(uint32_t unit, uint32_t nstrips, uint32_t lag) {
    bool stepB = false;
    uint32_t strip = 0;
    if (lag > nstrips || !rfb::fuse4_decode_unit(unit, nstrips, lag, stepB, strip)) return -1;
    return ((int64_t)(stepB ? 1 : 0) << 32) | (int64_t)strip;
}
RFB_EXPORT const char *rfb200_version(void) { return "rocketfft_b200 0.1.0 (sm_100a)"; }
