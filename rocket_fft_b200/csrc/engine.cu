// Line engine: batched 1-D complex DFTs over arbitrarily strided memory.
//   * lines that fit in shared memory      -> one fused tile kernel (tile_kernel.cuh)
//   * long smooth lines, n = n1*n2          -> "four-step": two tile-kernel launches with the
//                                             n1 x n2 twiddle fused into the first and the
//                                             transposition folded into the second's addressing
//   * lines with a prime factor > 64        -> Bluestein chirp-z on a power-of-two length
// (Counterpart of general_nd / pocketfft_c / fftblue in the reference,
//  _pocketfft_hdronly.h:3568-3607, 2834-2912, 2723-2828 -- different algorithms, new code.)
#include "engine.h"

#include <algorithm>
#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <stdlib.h>
#include <string.h>

#include "geom_fill.cuh"
#include "tile_kernel.cuh"

namespace rfb {

static std::atomic<uint64_t> g_launches{0};
uint64_t launch_count() { return g_launches.load(); }
void launch_count_reset() { g_launches.store(0); }
// Launch trace (rfb200_launch_trace): the names of the kernels launched since the last reset, so that a benchmark can name
// the kernel it times from what actually ran.
static std::mutex g_trace_mu;
static std::vector<std::string> g_trace;
static std::atomic<bool> g_trace_on{false};
void launch_trace_enable(bool on) {
    std::lock_guard<std::mutex> lk(g_trace_mu);
    g_trace_on.store(on);
    g_trace.clear();
}
std::string launch_trace_get() {
    std::lock_guard<std::mutex> lk(g_trace_mu);
    std::string r;
    for (auto &n : g_trace) { if (!r.empty()) r += ";"; r += n; }
    return r;
}
void count_launch(const char *name) {
    g_launches.fetch_add(1);
    if (g_trace_on.load() && name) {
        std::lock_guard<std::mutex> lk(g_trace_mu);
        if (g_trace.size() < 256) g_trace.push_back(name);
    }
}

#define RFB_AFTER_LAUNCH_N(name)                 \
    do {                                        \
        count_launch(name);                     \
        RFB_CUDA_CHECK(cudaGetLastError());     \
    } while (0)
#define RFB_AFTER_LAUNCH() RFB_AFTER_LAUNCH_N("engine helper kernel")

static const size_t MAX_SMEM = 227 * 1024;

static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

// ---------------------------------------------------------------------------------------
// scratch
// ---------------------------------------------------------------------------------------
static void init_pool_once() {
    static std::mutex mu;
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    if (dev < 64 && !done[dev]) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t thr = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        done[dev] = true;
    }
}

Scratch::Scratch(size_t bytes, cudaStream_t st) : s(st) {
    init_pool_once();
    RFB_CUDA_CHECK(cudaMallocAsync(&p, bytes ? bytes : 16, st));
}
Scratch::~Scratch() {
    if (p) cudaFreeAsync(p, s);
}

// ---------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------
static inline int64_t iabs64(int64_t v) { return v < 0 ? -v : v; }

struct BatchIdx {
    int nd;
    uint32_t ext[8];
    FastDiv d[8];
    int64_t is[8], os[8];
};

static uint64_t total_lines(const std::vector<Dim> &b) {
    uint64_t L = 1;
    for (auto &d : b) L *= (uint64_t)d.n;
    return L;
}

static BatchIdx make_batch_idx(const std::vector<Dim> &b) {
    BatchIdx bi;
    if (b.size() > 8) { set_error("more than 9 array dimensions are not supported"); throw Error(); }
    bi.nd = (int)b.size();
    for (int i = 0; i < bi.nd; ++i) {
        bi.ext[i] = (uint32_t)b[i].n;
        bi.d[i] = make_fastdiv((uint32_t)b[i].n);
        bi.is[i] = b[i].is;
        bi.os[i] = b[i].os;
    }
    return bi;
}

__device__ __forceinline__ void batch_offsets(const BatchIdx &bi, uint32_t l, int64_t &oi, int64_t &oo) {
    oi = 0; oo = 0;
    for (int d = 0; d < bi.nd; ++d) {
        uint32_t q, r;
        fdivmod(l, bi.d[d], q, r);
        oi += (int64_t)r * bi.is[d];
        oo += (int64_t)r * bi.os[d];
        l = q;
    }
}

// ---------------------------------------------------------------------------------------
// gather / scatter kernels for lines that go through scratch (long lines with a fused mode)
// grid: x over elements of a line, y over lines (grid-stride)
// ---------------------------------------------------------------------------------------
template <typename T, bool ALIGNED>
__global__ void gather_lines_kernel(BatchIdx bi, uint32_t L, const char *in, int64_t sa, cx<T> *scratch, uint32_t n,
                                    uint32_t n_in, int mode, int flags) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    for (uint32_t l = blockIdx.y; l < L; l += gridDim.y) {
        int64_t oi, oo;
        batch_offsets(bi, l, oi, oo);
        scratch[(size_t)l * n + e] = load_value<T, ALIGNED>(mode, flags, in + oi, sa, e, n, n_in);
    }
}

template <typename T, bool ALIGNED>
__global__ void scatter_lines_kernel(BatchIdx bi, uint32_t L, const cx<T> *scratch, char *out, int64_t sa, uint32_t n,
                                     uint32_t n_out, int mode, int flags) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_out) return;
    for (uint32_t l = blockIdx.y; l < L; l += gridDim.y) {
        int64_t oi, oo;
        batch_offsets(bi, l, oi, oo);
        const cx<T> v = scratch[(size_t)l * n + store_bin(mode, j)];
        store_value<T, ALIGNED>(mode, flags, out + oo, sa, j, v);
    }
}

// ---------------------------------------------------------------------------------------
// Bluestein kernels
// ---------------------------------------------------------------------------------------
// a[l][m] = x[l][m] * chirp[m] (m < n), 0 (n <= m < M); backward handled by the swap identity
template <typename T, bool ALIGNED>
__global__ void blue_pre_kernel(BatchIdx bi, uint32_t L, const char *in, int64_t sa, cx<T> *a, uint32_t n, uint32_t M,
                                const cx<T> *__restrict__ chirp, int backward) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    for (uint32_t l = blockIdx.y; l < L; l += gridDim.y) {
        cx<T> v = mk<T>(T(0), T(0));
        if (m < n) {
            int64_t oi, oo;
            batch_offsets(bi, l, oi, oo);
            v = ld_cx<T, ALIGNED>(in + oi + (int64_t)m * sa);
            if (backward) v = cswap(v);
            v = cmul(v, chirp[m]);
        }
        a[(size_t)l * M + m] = v;
    }
}

template <typename T>
__global__ void blue_mul_kernel(cx<T> *a, uint32_t L, uint32_t M, const cx<T> *__restrict__ bhat) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const cx<T> b = bhat[m];
    for (uint32_t l = blockIdx.y; l < L; l += gridDim.y) a[(size_t)l * M + m] = cmul(a[(size_t)l * M + m], b);
}

template <typename T, bool ALIGNED>
__global__ void blue_post_kernel(BatchIdx bi, uint32_t L, const cx<T> *y, char *out, int64_t sa, uint32_t n, uint32_t M,
                                 const cx<T> *__restrict__ chirp, T fct, int backward) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const cx<T> w = chirp[k];
    for (uint32_t l = blockIdx.y; l < L; l += gridDim.y) {
        int64_t oi, oo;
        batch_offsets(bi, l, oi, oo);
        cx<T> v = cscale(cmul(y[(size_t)l * M + k], w), fct);
        if (backward) v = cswap(v);
        st_cx<T, ALIGNED>(out + oo + (int64_t)k * sa, v);
    }
}

// b[m] = conj(chirp[|m|]) / M wrapped onto length M
template <typename T>
__global__ void blue_kernel_seq(cx<T> *b, uint32_t n, uint32_t M, const cx<T> *__restrict__ chirp) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    cx<T> v = mk<T>(T(0), T(0));
    const T inv = T(1) / T(M);
    if (m < n) v = mk<T>(chirp[m].x * inv, -chirp[m].y * inv);
    else if (M - m < n) v = mk<T>(chirp[M - m].x * inv, -chirp[M - m].y * inv);
    b[m] = v;
}

// four-step order of a length n1*n2 table: dst[k1*n2 + k2] = src[k2*n1 + k1]
template <typename T>
__global__ void table_to_fourstep_order(cx<T> *dst, const cx<T> *__restrict__ src, uint32_t n1, uint32_t n2) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n1 * n2) return;
    const uint32_t k1 = i / n2, k2 = i % n2;
    dst[i] = src[k2 * n1 + k1];
}

// ---------------------------------------------------------------------------------------
// tile launch
// ---------------------------------------------------------------------------------------
struct TilePlan {
    std::vector<uint32_t> sched;
    uint32_t pitch, padsh, W, threads;
    size_t smem;
    bool load_lf, store_lf;
};

static bool alignment_ok(const LineJob &job, const std::vector<Dim> &dims) {
    const int64_t esz = job.prec ? 16 : 8;
    auto ok = [&](int64_t v) { return (v % esz) == 0; };
    bool cin = (job.load_mode == LD_C2C || job.load_mode == LD_HERM);
    bool cout = (job.store_mode == ST_C2C || job.store_mode == ST_HALF);
    bool a = true;
    if (cin) {
        a = a && ok((int64_t)(uintptr_t)job.in) && ok(job.is);
        for (auto &d : dims) a = a && ok(d.is);
    }
    for (auto ptr : job.split_out) a = a && ok((int64_t)(uintptr_t)ptr);
    if (cout) {
        a = a && ok((int64_t)(uintptr_t)job.out) && ok(job.os);
        for (auto &d : dims) a = a && ok(d.os);
    }
    return a;
}

// Decide the tile shape for lines of length n; false if a line does not fit.
static bool plan_tile(const LineJob &job, const std::vector<Dim> &dims, TilePlan &tp) {
    const uint64_t n = job.n;
    const size_t esz = job.prec ? 16 : 8;
    if (n > (1u << 20)) return false;
    tp.sched = radix_schedule(n, RMAX_GENERIC);  // n == 1: a single radix-1 pass
    if (tp.sched.empty() || tp.sched.size() > (size_t)MAXP) return false;
    tp.padsh = job.prec ? 4 : 5;
    uint32_t pitch = (uint32_t)((n - 1) + ((n - 1) >> tp.padsh) + 1);
    pitch |= 1u;
    tp.pitch = pitch;
    const size_t line_bytes = (size_t)pitch * esz;
    if (line_bytes > MAX_SMEM) return false;
    const uint64_t e0 = dims.empty() ? 1 : (uint64_t)dims[0].n;
    tp.load_lf = !dims.empty() && iabs64(dims[0].is) < iabs64(job.is);
    tp.store_lf = !dims.empty() && iabs64(dims[0].os) < iabs64(job.os);
    const uint64_t wfit = MAX_SMEM / line_bytes;
    uint64_t W;
    if (tp.load_lf || tp.store_lf) {
        // neighbouring lines are adjacent in memory: take enough of them for 128-byte rows,
        // more when lines are short
        uint64_t wpref = 128 / esz;
        const size_t budget = 64 * 1024;
        while (wpref * 2 * line_bytes <= budget && wpref < 64) wpref *= 2;
        W = std::min<uint64_t>(wpref, wfit);
    } else {
        const size_t budget = 32 * 1024;  // several CTAs per SM for short lines
        W = std::max<uint64_t>(1, budget / line_bytes);
        W = std::min<uint64_t>(W, wfit);
        W = std::min<uint64_t>(W, 256);
    }
    W = std::max<uint64_t>(1, std::min<uint64_t>(W, e0));
    tp.W = (uint32_t)W;
    tp.smem = (size_t)W * line_bytes;
    uint64_t work = W * n;
    const uint64_t max_th = job.prec ? 512 : 1024;
    uint32_t th = (uint32_t)std::min<uint64_t>(max_th, std::max<uint64_t>(64, ((work / 4 + 31) / 32) * 32));
    tp.threads = th;
    return true;
}

template <typename T, bool ALIGNED>
static void launch_tile_typed(const LineJob &job, const std::vector<Dim> &dims, const TilePlan &tp, cudaStream_t s) {
    TileGeom<T> g;
    const uint64_t ntiles = fill_geom<T>(g, job, dims, tp.W, tp.load_lf, tp.store_lf);
    const uint32_t n = (uint32_t)job.n;
    g.npass = (uint32_t)tp.sched.size();
    uint32_t l1 = 1, twoff = 0;
    for (uint32_t i = 0; i < g.npass; ++i) {
        PassInfo &ps = g.pass[i];
        ps.R = tp.sched[i];
        ps.l1 = l1;
        ps.ido = n / (l1 * ps.R);
        ps.d_ido = make_fastdiv(ps.ido);
        ps.d_nbl = make_fastdiv(n / ps.R);
        ps.d_R = make_fastdiv(ps.R);
        ps.twoff = twoff;
        if (ps.ido > 1) twoff += (ps.R - 1) * ps.ido;
        l1 *= ps.R;
    }
    g.pitch = tp.pitch;
    g.padsh = tp.padsh;
    g.tw = (n > 1) ? (const cx<T> *)get_table(TAB_LINE, job.prec, n, 0) : nullptr;
    g.ptw = (n > 1) ? (const cx<T> *)get_table(TAB_TILE, job.prec, n, 0) : nullptr;
    auto kern = fft_tile_kernel<T, ALIGNED>;
    static thread_local int dev_set = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev_set != dev) {
        RFB_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MAX_SMEM));
        dev_set = dev;
    }
    kern<<<(unsigned)ntiles, tp.threads, tp.smem, s>>>(g);
    RFB_AFTER_LAUNCH_N(sizeof(T) == 8 ? "fft_tile_kernel<double>" : "fft_tile_kernel<float>");
}

static void launch_tile(const LineJob &job, const std::vector<Dim> &dims, const TilePlan &tp, cudaStream_t s) {
    const bool al = alignment_ok(job, dims);
    if (job.prec) {
        if (al) launch_tile_typed<double, true>(job, dims, tp, s);
        else launch_tile_typed<double, false>(job, dims, tp, s);
    } else {
        if (al) launch_tile_typed<float, true>(job, dims, tp, s);
        else launch_tile_typed<float, false>(job, dims, tp, s);
    }
}

// ---------------------------------------------------------------------------------------
// elementwise launch helpers
// ---------------------------------------------------------------------------------------
static dim3 ew_grid(uint64_t n_e, uint64_t L) {
    return dim3((unsigned)((n_e + 255) / 256), (unsigned)std::min<uint64_t>(L, 32768), 1);
}

// ---------------------------------------------------------------------------------------
// the three algorithms
// ---------------------------------------------------------------------------------------
static void run_fourstep(const LineJob &job, const std::vector<Dim> &dims, cudaStream_t s);
static void run_bluestein(const LineJob &job, const std::vector<Dim> &dims, cudaStream_t s);
static void run_via_scratch(const LineJob &job, const std::vector<Dim> &dims, cudaStream_t s);
static void run_norm(const LineJob &job, const std::vector<Dim> &dims, cudaStream_t s);

static bool choose_split(uint64_t n, int prec, uint64_t &n1, uint64_t &n2) {
    // divisor pair closest to sqrt(n) whose members both factor into radices <= RMAX_GENERIC
    auto pf = prime_factors(n);
    if (pf.empty() || pf.back() > RMAX_GENERIC) return false;
    std::vector<uint64_t> divs{1};
    for (size_t i = 0; i < pf.size();) {
        size_t j = i;
        while (j < pf.size() && pf[j] == pf[i]) ++j;
        size_t base = divs.size();
        uint64_t pw = 1;
        for (size_t e = 0; e < j - i; ++e) {
            pw *= pf[i];
            for (size_t k = 0; k < base; ++k) divs.push_back(divs[k] * pw);
        }
        i = j;
    }
    uint64_t best = 0;
    for (auto d : divs)
        if (d * d <= n && d > best) best = d;
    if (best <= 1) return false;
    n1 = best;
    n2 = n / best;
    const size_t esz = prec ? 16 : 8;
    // the longer factor must still fit a tile of at least two lines
    if ((n2 + (n2 >> 4) + 2) * esz * 2 > MAX_SMEM) return false;
    return true;
}

// drop unit dims, sort by stride; false if the job is empty
static bool normalise(LineJob &job, std::vector<Dim> &dims) {
    if (job.n == 0) return false;
    for (auto &d : job.batch) {
        if (d.n == 0) return false;
        if (d.n >= (1ll << 31)) { set_error("array dimension too large"); throw Error(); }
        if (d.n == 1) continue;
        dims.push_back(d);
    }
    if (job.n >= (1ull << 31)) { set_error("transform length too large"); throw Error(); }
    bool have_tw = false;
    for (auto &d : dims) have_tw = have_tw || d.tw;
    if (!have_tw) job.twN = 0;
    std::stable_sort(dims.begin(), dims.end(), [](const Dim &a, const Dim &b) {
        return std::min(iabs64(a.is), iabs64(a.os)) < std::min(iabs64(b.is), iabs64(b.os));
    });
    if (total_lines(dims) >= (1ull << 31)) { set_error("too many lines"); throw Error(); }
    return true;
}

void run_lines(const LineJob &job_in, cudaStream_t s) {
    init_pool_once();
    LineJob job = job_in;
    std::vector<Dim> dims;
    if (!normalise(job, dims)) return;
    run_norm(job, dims, s);
}

static bool pow2_try(const LineJob &job, const std::vector<Dim> &dims, cudaStream_t s) {
    if (dims.size() > (size_t)MAXB) return false;
    if (!alignment_ok(job, dims)) return false;
    const bool load_lf = !dims.empty() && iabs64(dims[0].is) < iabs64(job.is);
    const bool store_lf = !dims.empty() && iabs64(dims[0].os) < iabs64(job.os);
    return job.prec ? launch_pow2_f64(job, dims, load_lf, store_lf, s) : launch_pow2_f32(job, dims, load_lf, store_lf, s);
}

bool run_lines_pow2(const LineJob &job_in, cudaStream_t s) {
    LineJob job = job_in;
    std::vector<Dim> dims;
    if (!normalise(job, dims)) return true;
    return pow2_try(job, dims, s);
}

bool run_lines_tile(const LineJob &job_in, cudaStream_t s) {
    init_pool_once();
    LineJob job = job_in;
    std::vector<Dim> dims;
    if (!normalise(job, dims)) return true;
    if (dims.size() > (size_t)MAXB) return false;
    TilePlan tp;
    if (!plan_tile(job, dims, tp)) return false;
    launch_tile(job, dims, tp, s);
    return true;
}

// dims: extent > 1, sorted by stride (dims[0] = the tile dim)
static void run_norm(const LineJob &job, const std::vector<Dim> &dims, cudaStream_t s) {
    // more batch dims than one launch takes: peel the outermost ones on the host
    if (dims.size() > (size_t)MAXB) {
        std::vector<Dim> inner(dims.begin(), dims.end() - 1);
        const Dim &o = dims.back();
        if (o.tw && job.twN) { set_error("internal: four-step dim peeled"); throw Error(); }
        for (int64_t i = 0; i < o.n; ++i) {
            LineJob sub = job;
            sub.in = job.in + i * o.is;
            sub.out = job.out + i * o.os;
            run_norm(sub, inner, s);
        }
        return;
    }
    const bool simple = (job.load_mode == LD_C2C || job.load_mode == LD_REAL) && job.store_mode == ST_C2C &&
                        job.flags == 0 && (job.n_in == 0 || job.n_in == job.n);
    const size_t esz = job.prec ? 16 : 8;
    if (launch_pow2_stream_f32(job, dims, s)) return;
    // power-of-two lines: register-resident Stockham kernel
    static const bool use_pow2 = env_int("RFB200_NO_POW2", 0) == 0;
    if (use_pow2 && alignment_ok(job, dims)) {
        const bool load_lf = !dims.empty() && iabs64(dims[0].is) < iabs64(job.is);
        const bool store_lf = !dims.empty() && iabs64(dims[0].os) < iabs64(job.os);
        const bool done = job.prec ? launch_pow2_f64(job, dims, load_lf, store_lf, s)
                                   : launch_pow2_f32(job, dims, load_lf, store_lf, s);
        if (done) return;
    }
    if (job.conv) { set_error("internal: convolution rows need the power-of-two kernel"); throw Error(); }
    if (!job.split_out.empty()) {
        set_error("scatter output needs aligned arrays and a power-of-two axis length: 16..16384 for contiguous lines, 16..2048 for strided ones");
        throw Error();
    }
    // smooth non-power-of-two lines: register-resident mixed-radix kernel
    static const bool use_regmix = env_int("RFB200_NO_REGMIX", 0) == 0;
    if (use_regmix && !job.pre_tab && !job.post_tab) {
        const bool load_lf = !dims.empty() && iabs64(dims[0].is) < iabs64(job.is);
        const bool store_lf = !dims.empty() && iabs64(dims[0].os) < iabs64(job.os);
        const bool al = alignment_ok(job, dims);
        if (launch_spec_jit(job, dims, load_lf, store_lf, al, s)) return;
        if (launch_regmix(job, dims, load_lf, store_lf, al, s)) return;
    }
    if (job.pre_tab || job.post_tab) {
        // fused factors exist only in the power-of-two kernel: long (or strided long) lines split first
        uint64_t a, b;
        if (job.twN == 0 && job.g_mul == 1 && choose_split(job.n, job.prec, a, b)) {
            run_fourstep(job, dims, s);
            return;
        }
        set_error("internal: fused element-wise factors need the power-of-two kernel");
        throw Error();
    }
    TilePlan tp;
    bool tile_ok = plan_tile(job, dims, tp);
    if (tile_ok && simple && (tp.load_lf || tp.store_lf) && job.twN == 0) {
        // strided long lines: a tile of >= 32 bytes per row does not fit -> four-step keeps
        // the accesses coalesced
        const uint32_t wmin = (uint32_t)(32 / esz) * 2;
        uint64_t a, b;
        if (tp.W < wmin && tp.W < dims[0].n && job.n >= 1024 && choose_split(job.n, job.prec, a, b)) tile_ok = false;
    }
    if (tile_ok) {
        launch_tile(job, dims, tp, s);
        return;
    }
    if (!simple) {
        run_via_scratch(job, dims, s);
        return;
    }
    if (job.twN) { set_error("internal: nested four-step"); throw Error(); }
    uint64_t n1, n2;
    if (choose_split(job.n, job.prec, n1, n2)) run_fourstep(job, dims, s);
    else run_bluestein(job, dims, s);
}

static void run_via_scratch(const LineJob &job, const std::vector<Dim> &dims, cudaStream_t s) {
    const size_t esz = job.prec ? 16 : 8;
    const uint64_t L = total_lines(dims);
    const uint64_t n = job.n;
    Scratch sc(L * n * esz, s);
    BatchIdx bi = make_batch_idx(dims);
    const bool al = alignment_ok(job, dims);
    const uint32_t n_in = (uint32_t)(job.n_in ? job.n_in : n);
    dim3 grid = ew_grid(n, L);
    if (job.prec) {
        if (al) gather_lines_kernel<double, true><<<grid, 256, 0, s>>>(bi, (uint32_t)L, job.in, job.is, (double2 *)sc.p, (uint32_t)n, n_in, job.load_mode, job.flags);
        else gather_lines_kernel<double, false><<<grid, 256, 0, s>>>(bi, (uint32_t)L, job.in, job.is, (double2 *)sc.p, (uint32_t)n, n_in, job.load_mode, job.flags);
    } else {
        if (al) gather_lines_kernel<float, true><<<grid, 256, 0, s>>>(bi, (uint32_t)L, job.in, job.is, (float2 *)sc.p, (uint32_t)n, n_in, job.load_mode, job.flags);
        else gather_lines_kernel<float, false><<<grid, 256, 0, s>>>(bi, (uint32_t)L, job.in, job.is, (float2 *)sc.p, (uint32_t)n, n_in, job.load_mode, job.flags);
    }
    RFB_AFTER_LAUNCH();
    LineJob mid;
    mid.prec = job.prec;
    mid.n = n;
    mid.is = mid.os = (int64_t)esz;
    mid.batch.push_back(Dim{(int64_t)L, (int64_t)(n * esz), (int64_t)(n * esz), false});
    mid.in = (const char *)sc.p;
    mid.out = (char *)sc.p;
    mid.backward = job.backward;
    mid.fct = job.fct;
    run_lines(mid, s);
    const uint32_t n_out = (job.store_mode == ST_HALF) ? (uint32_t)(n / 2 + 1) : (uint32_t)n;
    grid = ew_grid(n_out, L);
    if (job.prec) {
        if (al) scatter_lines_kernel<double, true><<<grid, 256, 0, s>>>(bi, (uint32_t)L, (const double2 *)sc.p, job.out, job.os, (uint32_t)n, n_out, job.store_mode, job.flags);
        else scatter_lines_kernel<double, false><<<grid, 256, 0, s>>>(bi, (uint32_t)L, (const double2 *)sc.p, job.out, job.os, (uint32_t)n, n_out, job.store_mode, job.flags);
    } else {
        if (al) scatter_lines_kernel<float, true><<<grid, 256, 0, s>>>(bi, (uint32_t)L, (const float2 *)sc.p, job.out, job.os, (uint32_t)n, n_out, job.store_mode, job.flags);
        else scatter_lines_kernel<float, false><<<grid, 256, 0, s>>>(bi, (uint32_t)L, (const float2 *)sc.p, job.out, job.os, (uint32_t)n, n_out, job.store_mode, job.flags);
    }
    RFB_AFTER_LAUNCH();
}

static void run_fourstep_plain(const LineJob &job, const std::vector<Dim> &dims, cudaStream_t s, void *work = nullptr);

static void run_fourstep(const LineJob &job, const std::vector<Dim> &dims, cudaStream_t s) {
    uint64_t n1, n2;
    choose_split(job.n, job.prec, n1, n2);
    // Long strided complex64 lines of arrays beyond the L2 cache: both steps in ONE warp-specialised persistent kernel fed by
    // the copy engine, the intermediate in an L2-resident ring (fused4v2_kernel.cuh).  RFB200_FUSE4=0 (read per call, so that
    // tests can compare the paths) forces the two-launch form.
    const char *f4 = getenv("RFB200_FUSE4");
    if ((!f4 || atoi(f4) != 0) && launch_fourstep_fused2_f32(job, dims, s)) return;
    run_fourstep_plain(job, dims, s);
}

static void run_fourstep_plain(const LineJob &job, const std::vector<Dim> &dims, cudaStream_t s, void *work) {
    const int64_t esz = job.prec ? 16 : 8;
    uint64_t n1, n2;
    choose_split(job.n, job.prec, n1, n2);
    // Scratch layout: when the lines are strided (neighbouring lines adjacent in memory) the
    // scratch keeps that neighbour dim fastest, [..][n][dim0], so that both steps read and write
    // rows of adjacent lines -- rows padded to a multiple of 128 bytes (with 8193 neighbouring lines dense rows would each
    // start at another 8-byte phase: no 16-byte accesses, every row segment straddling one more sector);
    // contiguous lines use [..][dim0][n].
    const bool lf = !dims.empty() && (iabs64(dims[0].is) < iabs64(job.is) || iabs64(dims[0].os) < iabs64(job.os));
    std::vector<int64_t> sstr(dims.size());
    int64_t s_axis, acc;
    if (lf) {
        sstr[0] = esz;
        s_axis = dims[0].n * esz;
        if (!work && s_axis >= 1024) s_axis = (s_axis + 127) & ~(int64_t)127;
        acc = s_axis * (int64_t)job.n;
        for (size_t i = 1; i < dims.size(); ++i) { sstr[i] = acc; acc *= dims[i].n; }
    } else {
        s_axis = esz;
        acc = (int64_t)job.n * esz;
        for (size_t i = 0; i < dims.size(); ++i) { sstr[i] = acc; acc *= dims[i].n; }
    }
    std::unique_ptr<Scratch> own;
    if (!work) {
        own.reset(new Scratch((uint64_t)acc, s));
        work = own->p;
    }
    // A: for every residue j0 (mod n2) an n1-point DFT over j1 of x[j1*n2 + j0], times
    //    exp(-2 pi i j0 k1 / n), stored at scratch[k1*n2 + j0]
    LineJob A;
    A.prec = job.prec;
    A.n = n1;
    A.is = (int64_t)n2 * job.is;
    A.os = (int64_t)n2 * s_axis;
    for (size_t i = 0; i < dims.size(); ++i) A.batch.push_back(Dim{dims[i].n, dims[i].is, sstr[i], false});
    A.batch.push_back(Dim{(int64_t)n2, job.is, s_axis, true});
    A.in = job.in;
    A.out = (char *)work;
    A.backward = job.backward;
    A.fct = 1.0;
    A.load_mode = job.load_mode;
    A.twN = job.n;
    if (job.pre_tab) {  // global element index j1*n2 + j0
        A.pre_tab = job.pre_tab;
        A.pre_bound = job.pre_bound;
        A.pre_swap = job.pre_swap;
        A.g_mul = n2;
    }
    run_lines(A, s);
    // B: for every k1 an n2-point DFT over j0 of scratch[k1*n2 + j0] -> X[k2*n1 + k1]
    LineJob B;
    B.prec = job.prec;
    B.n = n2;
    B.is = s_axis;
    B.os = (int64_t)n1 * job.os;
    for (size_t i = 0; i < dims.size(); ++i) B.batch.push_back(Dim{dims[i].n, sstr[i], dims[i].os, false});
    B.batch.push_back(Dim{(int64_t)n1, (int64_t)n2 * s_axis, job.os, job.post_tab != nullptr});
    B.in = (const char *)work;
    B.out = job.out;
    B.backward = job.backward;
    B.fct = job.fct;
    if (job.post_tab) {  // global bin k2*n1 + k1
        B.post_tab = job.post_tab;
        B.post_bound = job.post_bound;
        B.post_swap = job.post_swap;
        B.g_mul = n1;
    }
    run_lines(B, s);
}

static std::mutex g_blue_mu;

static void run_bluestein(const LineJob &job, const std::vector<Dim> &dims, cudaStream_t s) {
    if (job.load_mode != LD_C2C) {
        // real input: go through the gather path first
        run_via_scratch(job, dims, s);
        return;
    }
    const size_t esz = job.prec ? 16 : 8;
    const uint64_t n = job.n;
    uint64_t M = 1;
    while (M < 2 * n - 1) M <<= 1;
    const uint64_t L = total_lines(dims);
    const void *chirp = get_table(TAB_CHIRP, job.prec, n, 0);
    const void *bhat;
    {
        std::lock_guard<std::mutex> lk(g_blue_mu);
        bool created = false;
        void *w = nullptr;
        bhat = get_table(TAB_CHIRP_FFT, job.prec, n, M, &created, &w);
        if (created) {
            // the table's contents are computed here; if that fails the (unfilled) entry must not stay in the cache
            try {
                dim3 grid((unsigned)((M + 255) / 256));
                if (job.prec) blue_kernel_seq<double><<<grid, 256, 0, s>>>((double2 *)w, (uint32_t)n, (uint32_t)M, (const double2 *)chirp);
                else blue_kernel_seq<float><<<grid, 256, 0, s>>>((float2 *)w, (uint32_t)n, (uint32_t)M, (const float2 *)chirp);
                RFB_AFTER_LAUNCH();
                LineJob f;
                f.prec = job.prec;
                f.n = M;
                f.is = f.os = (int64_t)esz;
                f.in = (const char *)w;
                f.out = (char *)w;
                run_lines(f, s);
                RFB_CUDA_CHECK(cudaStreamSynchronize(s));
            } catch (...) {
                discard_table(TAB_CHIRP_FFT, job.prec, n, M);
                throw;
            }
        }
    }
    // ---- fused pipeline: two FFT_M jobs, chirp / padding / spectrum multiply / truncation ride on
    //      their loads and stores (needs the power-of-two kernel for every launch) ----------------
    {
        int logM = 0;
        while ((1ull << logM) < M) ++logM;
        bool fusable = alignment_ok(job, dims) && env_int("RFB200_NO_FUSED_BLUESTEIN", 0) == 0 && logM >= 4;
        if (M > 2048) fusable = fusable && logM <= 22 && dims.size() + 1 <= (size_t)MAXB;
        else fusable = fusable && dims.size() <= (size_t)MAXB;
        const uint64_t max_lines_f = std::max<uint64_t>(1, (8ull << 30) / (M * esz));
        if (fusable && L <= max_lines_f) {
            const bool lf = !dims.empty() && (iabs64(dims[0].is) < iabs64(job.is) || iabs64(dims[0].os) < iabs64(job.os));
            uint64_t n1 = 0, n2 = 0;
            if (!lf && M > (1ull << 14) && job.is == (int64_t)esz && job.os == (int64_t)esz) {
                // rows of n2 = 4096 points (the contiguous kernel's sweet spot), columns of n1 = M / n2 >= 16 points:
                // short columns mean wide tiles (many neighbouring columns per CTA) for the strided passes
                int l2 = 12;
                while (logM - l2 < 4) --l2;
                while (logM - l2 > 11) ++l2;
                n2 = 1ull << l2;
                n1 = M >> l2;
                // Long contiguous lines: M = n1*n2 needs two launches per transform anyway.  A convolution does not
                // need its spectrum in natural order, so the forward transform leaves it in four-step order
                // ([k1][k2], bin k2*n1 + k1) and the backward transform consumes that order: of the four launches
                // only the two n1-point column passes touch strided memory, the two n2-point row passes run in
                // place on contiguous lines, and one work buffer suffices.
                //   A : n1-point DFTs down the columns of x (chirp and zero padding on the load), times w_M^(j0 k1)
                //   B : n2-point DFTs along the rows, times B^ (stored in four-step order)
                //   B': n2-point backward DFTs along the rows, times conj w_M^(j0 k1)
                //   A': n1-point backward DFTs down the columns -> natural order, chirp / fct / truncation on the store
                // (B and B' are one kernel, MODE 5 of the power-of-two kernel: three passes over HBM in total)
                const void *bhat_t;
                {
                    std::lock_guard<std::mutex> lk(g_blue_mu);
                    bool created = false;
                    void *w = nullptr;
                    bhat_t = get_table(TAB_CHIRP_FFT_T, job.prec, n, M, &created, &w);
                    if (created) {
                        try {
                            dim3 grid((unsigned)((M + 255) / 256));
                            if (job.prec) table_to_fourstep_order<double><<<grid, 256, 0, s>>>((double2 *)w, (const double2 *)bhat, (uint32_t)n1, (uint32_t)n2);
                            else table_to_fourstep_order<float><<<grid, 256, 0, s>>>((float2 *)w, (const float2 *)bhat, (uint32_t)n1, (uint32_t)n2);
                            RFB_AFTER_LAUNCH();
                            RFB_CUDA_CHECK(cudaStreamSynchronize(s));
                        } catch (...) {
                            discard_table(TAB_CHIRP_FFT_T, job.prec, n, M);
                            throw;
                        }
                    }
                }
                std::vector<int64_t> st(dims.size());
                int64_t acc2 = (int64_t)(M * esz);
                for (size_t i = 0; i < dims.size(); ++i) { st[i] = acc2; acc2 *= dims[i].n; }
                Scratch s1(L * M * esz, s);
                const int64_t e = (int64_t)esz, row = (int64_t)n2 * e;
                LineJob A;
                A.prec = job.prec;
                A.n = n1;
                A.is = row;
                A.os = row;
                for (size_t i = 0; i < dims.size(); ++i) A.batch.push_back(Dim{dims[i].n, dims[i].is, st[i], false});
                A.batch.push_back(Dim{(int64_t)n2, e, e, true});
                A.in = job.in;
                A.out = (char *)s1.p;
                A.twN = M;
                A.pre_tab = chirp;
                A.pre_bound = n;
                A.pre_swap = job.backward;
                A.g_mul = n2;
                run_lines(A, s);
                LineJob B;
                B.prec = job.prec;
                B.n = n2;
                B.is = B.os = e;
                for (size_t i = 0; i < dims.size(); ++i) B.batch.push_back(Dim{dims[i].n, st[i], st[i], false});
                B.batch.push_back(Dim{(int64_t)n1, row, row, true});
                B.in = (const char *)s1.p;
                B.out = (char *)s1.p;
                B.post_tab = bhat_t;
                B.post_bound = M;
                B.g_mul = 1;
                B.c_mul = n2;
                // B and B' work on the same rows: one kernel does both without leaving the SM when it can
                LineJob Cv = B;
                Cv.conv = true;
                Cv.pre_tab = bhat_t;
                Cv.pre_bound = M;
                Cv.post_tab = nullptr;
                Cv.post_bound = 0;
                Cv.twN = M;
                static const bool use_conv = env_int("RFB200_NO_CONV_ROW", 0) == 0;
                if (!(use_conv && run_lines_pow2(Cv, s))) {
                    run_lines(B, s);
                    LineJob Bi = B;
                    Bi.post_tab = nullptr;
                    Bi.post_bound = 0;
                    Bi.c_mul = 1;
                    Bi.backward = true;
                    Bi.twN = M;
                    run_lines(Bi, s);
                }
                LineJob Ai;
                Ai.prec = job.prec;
                Ai.n = n1;
                Ai.is = row;
                Ai.os = row;
                for (size_t i = 0; i < dims.size(); ++i) Ai.batch.push_back(Dim{dims[i].n, st[i], dims[i].os, false});
                Ai.batch.push_back(Dim{(int64_t)n2, e, e, true});
                Ai.in = (const char *)s1.p;
                Ai.out = job.out;
                Ai.backward = true;
                Ai.fct = job.fct;
                Ai.post_tab = chirp;
                Ai.post_bound = n;
                Ai.post_swap = job.backward;
                Ai.g_mul = n2;
                run_lines(Ai, s);
                return;
            }
            std::vector<int64_t> sstr(dims.size());
            int64_t s_axis, acc;
            if (lf) {
                sstr[0] = (int64_t)esz;
                s_axis = dims[0].n * (int64_t)esz;
                acc = s_axis * (int64_t)M;
                for (size_t i = 1; i < dims.size(); ++i) { sstr[i] = acc; acc *= dims[i].n; }
            } else {
                s_axis = (int64_t)esz;
                acc = (int64_t)(M * esz);
                for (size_t i = 0; i < dims.size(); ++i) { sstr[i] = acc; acc *= dims[i].n; }
            }
            Scratch s1(L * M * esz, s);
            LineJob f;
            f.prec = job.prec;
            f.n = M;
            f.is = job.is;
            f.os = s_axis;
            for (size_t i = 0; i < dims.size(); ++i) f.batch.push_back(Dim{dims[i].n, dims[i].is, sstr[i], false});
            f.in = job.in;
            f.out = (char *)s1.p;
            f.pre_tab = chirp;
            f.pre_bound = n;
            f.pre_swap = job.backward;
            f.post_tab = bhat;
            f.post_bound = M;
            run_lines(f, s);
            LineJob b;
            b.prec = job.prec;
            b.n = M;
            b.is = s_axis;
            b.os = job.os;
            for (size_t i = 0; i < dims.size(); ++i) b.batch.push_back(Dim{dims[i].n, sstr[i], dims[i].os, false});
            b.in = (const char *)s1.p;
            b.out = job.out;
            b.backward = true;
            b.fct = job.fct;
            b.post_tab = chirp;
            b.post_bound = n;
            b.post_swap = job.backward;
            run_lines(b, s);
            return;
        }
    }
    // process the lines in chunks so the padded work area stays bounded (<= ~6 GiB)
    const uint64_t max_lines = std::max<uint64_t>(1, (6ull << 30) / (M * esz));
    if (L > max_lines && dims.size() >= 1) {
        // split along the outermost batch dim
        std::vector<Dim> inner(dims.begin(), dims.end() - 1);
        const Dim &o = dims.back();
        const uint64_t inner_lines = total_lines(inner);
        const uint64_t step = std::max<uint64_t>(1, max_lines / std::max<uint64_t>(1, inner_lines));
        if (step < (uint64_t)o.n) {
            for (int64_t i = 0; i < o.n; i += (int64_t)step) {
                LineJob sub = job;
                sub.batch = inner;
                sub.batch.push_back(Dim{std::min<int64_t>((int64_t)step, o.n - i), o.is, o.os, false});
                sub.in = job.in + i * o.is;
                sub.out = job.out + i * o.os;
                run_lines(sub, s);
            }
            return;
        }
    }
    Scratch sc(L * M * esz, s);
    BatchIdx bi = make_batch_idx(dims);
    const bool al = alignment_ok(job, dims);
    dim3 grid = ew_grid(M, L);
    const int bw = job.backward ? 1 : 0;
    if (job.prec) {
        if (al) blue_pre_kernel<double, true><<<grid, 256, 0, s>>>(bi, (uint32_t)L, job.in, job.is, (double2 *)sc.p, (uint32_t)n, (uint32_t)M, (const double2 *)chirp, bw);
        else blue_pre_kernel<double, false><<<grid, 256, 0, s>>>(bi, (uint32_t)L, job.in, job.is, (double2 *)sc.p, (uint32_t)n, (uint32_t)M, (const double2 *)chirp, bw);
    } else {
        if (al) blue_pre_kernel<float, true><<<grid, 256, 0, s>>>(bi, (uint32_t)L, job.in, job.is, (float2 *)sc.p, (uint32_t)n, (uint32_t)M, (const float2 *)chirp, bw);
        else blue_pre_kernel<float, false><<<grid, 256, 0, s>>>(bi, (uint32_t)L, job.in, job.is, (float2 *)sc.p, (uint32_t)n, (uint32_t)M, (const float2 *)chirp, bw);
    }
    RFB_AFTER_LAUNCH();
    LineJob f;
    f.prec = job.prec;
    f.n = M;
    f.is = f.os = (int64_t)esz;
    f.batch.push_back(Dim{(int64_t)L, (int64_t)(M * esz), (int64_t)(M * esz), false});
    f.in = (const char *)sc.p;
    f.out = (char *)sc.p;
    run_lines(f, s);
    if (job.prec) blue_mul_kernel<double><<<grid, 256, 0, s>>>((double2 *)sc.p, (uint32_t)L, (uint32_t)M, (const double2 *)bhat);
    else blue_mul_kernel<float><<<grid, 256, 0, s>>>((float2 *)sc.p, (uint32_t)L, (uint32_t)M, (const float2 *)bhat);
    RFB_AFTER_LAUNCH();
    f.backward = true;
    run_lines(f, s);
    grid = ew_grid(n, L);
    if (job.prec) {
        if (al) blue_post_kernel<double, true><<<grid, 256, 0, s>>>(bi, (uint32_t)L, (const double2 *)sc.p, job.out, job.os, (uint32_t)n, (uint32_t)M, (const double2 *)chirp, (double)job.fct, bw);
        else blue_post_kernel<double, false><<<grid, 256, 0, s>>>(bi, (uint32_t)L, (const double2 *)sc.p, job.out, job.os, (uint32_t)n, (uint32_t)M, (const double2 *)chirp, (double)job.fct, bw);
    } else {
        if (al) blue_post_kernel<float, true><<<grid, 256, 0, s>>>(bi, (uint32_t)L, (const float2 *)sc.p, job.out, job.os, (uint32_t)n, (uint32_t)M, (const float2 *)chirp, (float)job.fct, bw);
        else blue_post_kernel<float, false><<<grid, 256, 0, s>>>(bi, (uint32_t)L, (const float2 *)sc.p, job.out, job.os, (uint32_t)n, (uint32_t)M, (const float2 *)chirp, (float)job.fct, bw);
    }
    RFB_AFTER_LAUNCH();
}

}  // namespace rfb
