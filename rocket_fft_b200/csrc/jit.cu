// Run-time specialisation of the register-resident Stockham kernel (spec_kernel.cuh) for the length at hand.
//
// For smooth non-power-of-two lengths the generic run-time-radix kernel (regmix_kernel.cuh) spends about
// half of its issue slots on index arithmetic and guards that are compile-time constants once n is known
// (measured: n = 1000: 27.8 % -> 54.2 % of HBM peak with the specialised instantiation).  The set of lengths is
// open-ended, so the instantiation happens on first use: NVRTC compiles `fft_spec_kernel<T, Plan, true>` from
// the same headers the build uses (-I csrc), for sm_100a, and the cubin is loaded with cudaLibraryLoadData.
// Plans are cached per (device, precision, n, threads-per-line, lines-per-CTA).  If NVRTC cannot be loaded or a
// compilation fails the caller silently stays on the build-time kernels (regmix / tile) -- still CUDA, never CPU.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "geom_fill.cuh"

namespace rfb {

namespace {

// ---- minimal NVRTC binding through dlopen ---------------------------------------------------------------
typedef struct _nvrtcProgram *nvrtcProgram;
struct Nvrtc {
    void *h = nullptr;
    int (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *);
    int (*DestroyProgram)(nvrtcProgram *);
    int (*CompileProgram)(nvrtcProgram, int, const char *const *);
    int (*GetProgramLogSize)(nvrtcProgram, size_t *);
    int (*GetProgramLog)(nvrtcProgram, char *);
    int (*AddNameExpression)(nvrtcProgram, const char *);
    int (*GetLoweredName)(nvrtcProgram, const char *, const char **);
    int (*GetCUBINSize)(nvrtcProgram, size_t *);
    int (*GetCUBIN)(nvrtcProgram, char *);
    bool ok = false;
};

Nvrtc &nvrtc() {
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *cands[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so",
                               "/usr/local/cuda/lib64/libnvrtc.so"};
        for (auto c : cands) {
            n.h = dlopen(c, RTLD_NOW | RTLD_LOCAL);
            if (n.h) break;
        }
        if (!n.h) return;
#define RFB_SYM(field, name)                          \
    *(void **)(&n.field) = dlsym(n.h, name);          \
    if (!n.field) return;
        RFB_SYM(CreateProgram, "nvrtcCreateProgram")
        RFB_SYM(DestroyProgram, "nvrtcDestroyProgram")
        RFB_SYM(CompileProgram, "nvrtcCompileProgram")
        RFB_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
        RFB_SYM(GetProgramLog, "nvrtcGetProgramLog")
        RFB_SYM(AddNameExpression, "nvrtcAddNameExpression")
        RFB_SYM(GetLoweredName, "nvrtcGetLoweredName")
        RFB_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
        RFB_SYM(GetCUBIN, "nvrtcGetCUBIN")
#undef RFB_SYM
        n.ok = true;
    });
    return n;
}

// directory of the kernel headers: <dir of this shared object>/csrc
std::string header_dir() {
    Dl_info info;
    if (dladdr((void *)&header_dir, &info) && info.dli_fname) {
        std::string p(info.dli_fname);
        size_t k = p.rfind('/');
        return (k == std::string::npos ? std::string(".") : p.substr(0, k)) + "/csrc";
    }
    return "csrc";
}

struct SpecKey {
    int dev, prec;
    uint64_t n;
    uint32_t tpl, w;
    int kmode;
    bool operator<(const SpecKey &o) const {
        return std::tie(dev, prec, n, tpl, w, kmode) < std::tie(o.dev, o.prec, o.n, o.tpl, o.w, o.kmode);
    }
};
struct SpecEntry {
    cudaKernel_t kern = nullptr;  // nullptr: compilation failed, do not retry
    size_t smem = 0;
    uint32_t cap = 16;
};
std::mutex g_mu;
std::map<SpecKey, SpecEntry> g_cache;

// ---- on-disk cache of compiled cubins (plan persistence across processes) ----------------------------------
// $RFB200_CACHE_DIR, else $XDG_CACHE_HOME/rocketfft_b200, else $HOME/.cache/rocketfft_b200; RFB200_CACHE_DIR=""
// disables it.  A file is `<lowered kernel name>\n<cubin bytes>`, named by a hash of the generated source, the
// compile options and the size + mtime of every kernel header, so a rebuilt library never picks up stale code.
uint64_t fnv1a(const std::string &s, uint64_t h = 1469598103934665603ull) {
    for (unsigned char c : s) h = (h ^ c) * 1099511628211ull;
    return h;
}

std::string cache_dir() {
    const char *e = getenv("RFB200_CACHE_DIR");
    std::string d;
    if (e) d = e;
    else if ((e = getenv("XDG_CACHE_HOME")) && *e) d = std::string(e) + "/rocketfft_b200";
    else if ((e = getenv("HOME")) && *e) d = std::string(e) + "/.cache/rocketfft_b200";
    if (d.empty()) return d;
    std::string part;
    for (size_t i = 0; i <= d.size(); ++i) {  // mkdir -p
        if (i == d.size() || d[i] == '/') {
            if (!part.empty()) mkdir(part.c_str(), 0755);
        }
        if (i < d.size()) part += d[i];
    }
    return d;
}

std::string header_stamp() {
    static const std::string stamp = [] {
        std::string s;
        const char *files[] = {"spec_kernel.cuh", "common.cuh", "line_io.cuh", "radix.cuh", "tile_kernel.cuh", "modes.h"};
        const std::string dir = header_dir();
        for (auto f : files) {
            struct stat st;
            if (stat((dir + "/" + f).c_str(), &st) == 0)
                s += std::string(f) + ":" + std::to_string((long long)st.st_size) + ":" +
                     std::to_string((long long)st.st_mtime) + ";";
        }
        return s;
    }();
    return stamp;
}

bool cache_read(const std::string &path, std::string &lname, std::vector<char> &cubin) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    std::vector<char> all;
    char buf[65536];
    size_t k;
    while ((k = fread(buf, 1, sizeof(buf), f)) > 0) all.insert(all.end(), buf, buf + k);
    fclose(f);
    auto nl = std::find(all.begin(), all.end(), '\n');
    if (nl == all.end() || all.end() - nl < 64) return false;
    lname.assign(all.begin(), nl);
    cubin.assign(nl + 1, all.end());
    return true;
}

void cache_write(const std::string &path, const std::string &lname, const std::vector<char> &cubin) {
    const std::string tmp = path + "." + std::to_string((long long)getpid());
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return;
    bool ok = fwrite(lname.data(), 1, lname.size(), f) == lname.size() && fputc('\n', f) != EOF &&
              fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
    ok = (fclose(f) == 0) && ok;
    if (ok) rename(tmp.c_str(), path.c_str());  // atomic: concurrent processes see a whole file or none
    else remove(tmp.c_str());
}

// NVRTC -> cubin for one plan; *spill = bytes of spill stores ptxas reported
bool nvrtc_build(const std::string &src, const std::string &name, uint64_t n, std::string &lname, std::vector<char> &cubin,
                 long *spill) {
    Nvrtc &rt = nvrtc();
    if (!rt.ok) return false;
    nvrtcProgram prog = nullptr;
    if (rt.CreateProgram(&prog, src.c_str(), "rfb_spec.cu", 0, nullptr, nullptr) != 0) return false;
    rt.AddNameExpression(prog, name.c_str());
    const std::string inc = "-I" + header_dir();
    const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", inc.c_str(), "-default-device", "-lineinfo",
                          "--ptxas-options=-v"};
    const int rc = rt.CompileProgram(prog, 6, opts);
    size_t ls = 0;
    rt.GetProgramLogSize(prog, &ls);
    std::string log(ls + 1, '\0');
    rt.GetProgramLog(prog, &log[0]);
    if (rc != 0) {
        if (getenv("RFB200_JIT_VERBOSE"))
            fprintf(stderr, "rocketfft_b200: NVRTC failed for n=%llu:\n%s\n", (unsigned long long)n, log.c_str());
        rt.DestroyProgram(&prog);
        return false;
    }
    *spill = 0;
    const size_t sp = log.find(" bytes spill stores");
    if (sp != std::string::npos) {
        size_t b = sp;
        while (b > 0 && isdigit((unsigned char)log[b - 1])) --b;
        *spill = atol(log.substr(b, sp - b).c_str());
    }
    const char *lowered = nullptr;
    size_t cs = 0;
    bool ok = rt.GetLoweredName(prog, name.c_str(), &lowered) == 0 && lowered && rt.GetCUBINSize(prog, &cs) == 0 && cs > 0;
    cubin.resize(cs);
    ok = ok && rt.GetCUBIN(prog, cubin.data()) == 0;
    lname = lowered ? lowered : "";
    rt.DestroyProgram(&prog);
    return ok;
}

bool compile_spec(int prec, int kmode, uint64_t n, uint32_t tpl, uint32_t w, uint32_t minb,
                  const std::vector<uint32_t> &sched, SpecEntry &out) {
    const char *T = prec ? "double" : "float";
    const bool verbose = getenv("RFB200_JIT_VERBOSE") != nullptr;
    std::string rad;
    for (size_t i = 0; i < sched.size(); ++i) rad += (i ? ", " : "") + std::to_string(sched[i]);
    const std::string name = std::string("rfb::fft_spec_kernel<") + T + ", rfb::PJ, true, " + std::to_string(kmode) + ">";
    std::string lname;
    std::vector<char> cubin;
    const std::string dir = cache_dir();
    char key[64];
    snprintf(key, sizeof(key), "%016llx",
             (unsigned long long)fnv1a(std::string(T) + "|" + std::to_string(kmode) + "|" + std::to_string(n) + "|" + std::to_string(tpl) + "|" +
                                       std::to_string(w) + "|" + std::to_string(minb) + "|" + rad + "|" + header_stamp()));
    const std::string path = dir.empty() ? std::string() : dir + "/spec_" + key + ".cubin";
    for (int attempt = 0; attempt < 2; ++attempt) {
        bool have = attempt == 0 && !path.empty() && cache_read(path, lname, cubin);
        const bool cached = have;
        if (have && verbose)
            fprintf(stderr, "rocketfft_b200: jit n=%llu %s from %s\n", (unsigned long long)n, T, path.c_str());
        // a kernel that spills its register tile is slower than one at half the occupancy target: retry once or twice
        for (uint32_t mb = minb; !have; mb = mb / 2) {
            char src[2048];
            snprintf(src, sizeof(src),
                     "#include \"spec_kernel.cuh\"\n"
                     "namespace rfb {\n"
                     "struct PJ {\n"
                     "    static constexpr int N = %llu, TPL = %u, W = %u, NPASS = %zu, MINB = %u;\n"
                     "    __host__ __device__ static constexpr int radix(int s) { constexpr int r[%zu] = {%s}; return r[s]; }\n"
                     "};\n"
                     "template __global__ void fft_spec_kernel<%s, PJ, true, %d>(const TileGeom<%s>);\n"
                     "}\n",
                     (unsigned long long)n, tpl, w, sched.size(), mb, sched.size(), rad.c_str(), T, kmode, T);
            long spill = 0;
            if (!nvrtc_build(src, name, n, lname, cubin, &spill)) return false;
            if (verbose)
                fprintf(stderr, "rocketfft_b200: jit n=%llu %s mode=%d tpl=%u w=%u minb=%u [%s] spill=%ld B\n",
                        (unsigned long long)n, T, kmode, tpl, w, mb, rad.c_str(), spill);
            if (spill <= 64 || mb <= 1) {
                if (!path.empty()) cache_write(path, lname, cubin);
                have = true;
            }
        }
        cudaLibrary_t lib = nullptr;
        cudaKernel_t k = nullptr;
        if (cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) == cudaSuccess &&
            cudaLibraryGetKernel(&k, lib, lname.c_str()) == cudaSuccess && k) {
            out.kern = k;
            return true;
        }
        cudaGetLastError();
        if (!cached) return false;
        remove(path.c_str());  // unreadable cache entry: compile afresh
    }
    return false;
}

}  // namespace

// Returns true if the job was launched on a run-time specialised kernel.
bool launch_spec_jit(const LineJob &job, const std::vector<Dim> &dims, bool load_lf, bool store_lf, bool aligned,
                     cudaStream_t s) {
    static const int mode = [] {
        const char *v = getenv("RFB200_JIT");
        return v ? atoi(v) : 1;  // 0: off, 1: for large batches, 2: always
    }();
    if (mode == 0 || !aligned || dims.size() > (size_t)MAXB) return false;
    // packed real transform (kernel MODE 1): contiguous even-length real line -> half spectrum through an n/2-point
    // complex transform; the two reals of a point are read as one complex value, which needs complex alignment
    const int64_t rsz = job.prec ? 8 : 4;
    bool packed = job.load_mode == LD_REAL && job.store_mode == ST_HALF && job.flags == 0 && job.twN == 0 && job.n % 2 == 0 &&
                  job.n >= 12 && !load_lf && !store_lf && job.is == rsz && ((uintptr_t)job.in % (2 * rsz)) == 0;
    for (auto &d : dims) packed = packed && (d.is % (2 * rsz)) == 0;
    if (packed && regmix_schedule(job.n / 2, job.prec ? 8 : 16).empty() && regmix_schedule(job.n / 2, 16).empty()) packed = false;
    // packed inverse (kernel MODE 2): Hermitian bins (any stride) -> even-length real line
    bool packed_inv = job.load_mode == LD_HERM && job.store_mode == ST_REAL && job.flags == 0 && job.twN == 0 &&
                      job.n % 2 == 0 && job.n >= 12 && !load_lf && !store_lf && (job.n_in == 0 || job.n_in == job.n);
    if (packed_inv && regmix_schedule(job.n / 2, job.prec ? 8 : 16).empty() && regmix_schedule(job.n / 2, 16).empty())
        packed_inv = false;
    bool out_vec = ((uintptr_t)job.out % (2 * rsz)) == 0;
    for (auto &d : dims) out_vec = out_vec && (d.os % (2 * rsz)) == 0;
    const int kmode = packed ? 1 : (packed_inv ? 2 : 0);
    packed = packed || packed_inv;  // from here on: geometry in half-length complex points
    const uint64_t n = packed ? job.n / 2 : job.n;
    if (n < 6 || n > 32768 || (n & (n - 1)) == 0) return false;
    if (job.store_mode == ST_HC || job.load_mode >= LD_DCT2 || job.store_mode >= ST_DCT2) return false;
    if (!job.split_out.empty() || job.pre_tab || job.post_tab) return false;
    uint64_t lines = 1;
    for (auto &d : dims) lines *= (uint64_t)d.n;
    // a compilation costs ~1 s: only for work that repays it (or when forced)
    if (mode == 1 && lines * n < (1ull << 21)) return false;
    const size_t esz = job.prec ? 16 : 8;
    uint32_t cap = job.prec ? 8 : 16;
    std::vector<uint32_t> sched = regmix_schedule(n, cap);
    if (sched.empty() && job.prec) {
        cap = 16;
        sched = regmix_schedule(n, cap);
    }
    if (sched.empty() || sched.size() > 12) return false;
    // points per thread: small budget first (more threads, fewer registers), the double budget for long lines
    const uint32_t e1 = job.prec ? 8 : 16;
    const bool lf = load_lf || store_lf;
    uint32_t tpl = 0, eb = 0;
    for (uint32_t e : {e1, 2 * e1}) {
        bool ok = true;
        uint32_t t = 1;
        for (auto R : sched) {
            if (R > e) { ok = false; break; }
            const uint32_t nb = (uint32_t)(n / R), jmax = e / R;
            t = std::max(t, (nb + jmax - 1) / jmax);
        }
        // the double budget runs at <= 88 registers per thread: 704 threads still fit one SM's register file
        const uint32_t tmax = (e == e1) ? 1024u : (lf ? 512u : 704u);
        if (ok && t <= (lf ? tmax / 4 : tmax)) { tpl = t; eb = e; break; }
    }
    if (!tpl) return false;
    const uint32_t tmax = (eb == e1) ? 1024u : (lf ? 512u : 704u);
    uint32_t w;
    if (lf) {
        w = (uint32_t)(128 / esz);
        while (w > 1 && w * tpl > tmax) w /= 2;
        if (w * esz < 32) return false;
    } else {
        w = std::max<uint32_t>(1, 256 / tpl);
        const uint64_t e0 = dims.empty() ? 1 : (uint64_t)dims[0].n;
        w = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(w, e0));
    }
    const uint32_t threads = w * tpl;
    if (threads > tmax) return false;
    const uint32_t pitch = (w == 1) ? (uint32_t)n : ((uint32_t)n | 1u);
    const size_t smem = (size_t)w * pitch * esz;
    if (smem > 227 * 1024) return false;
    // occupancy target: 64 registers with the small budget, 128 with the double one
    const uint32_t minb = std::max<uint32_t>(1, std::min<uint32_t>(8, tmax / threads));
    int dev = 0;
    cudaGetDevice(&dev);
    SpecEntry ent;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        SpecKey key{dev, job.prec, n, tpl, w, kmode};
        auto it = g_cache.find(key);
        if (it == g_cache.end()) {
            SpecEntry e;
            e.smem = smem;
            e.cap = cap;
            if (compile_spec(job.prec, kmode, n, tpl, w, minb, sched, e) && e.kern) {
                if (cudaFuncSetAttribute((const void *)e.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
                    cudaSuccess) {
                    cudaGetLastError();
                    e.kern = nullptr;
                }
            } else e.kern = nullptr;
            it = g_cache.emplace(key, e).first;
        }
        ent = it->second;
    }
    if (!ent.kern) return false;
    const void *ptw = get_table(TAB_REGMIX, job.prec, n, cap);
    auto launch = [&](auto tag) {
        using T = decltype(tag);
        TileGeom<T> g;
        LineJob j2 = job;
        if (packed) j2.n = n;  // geometry in complex points
        const uint64_t ntiles = fill_geom<T>(g, j2, dims, w, load_lf, store_lf);
        set_prefetch_by_mode<T>(g, job, dims, w);
        if (kmode == 2) g.flags = out_vec ? 0 : 1;
        if (kmode == 1) g.n_in = (uint32_t)(job.n_in ? job.n_in : job.n);  // real samples present
        if (packed) {
            g.twA = (const cx<T> *)get_table(TAB_LINE, job.prec, job.n, 0);
        }
        g.ptw = (const cx<T> *)ptw;
        void *args[] = {&g};
        RFB_CUDA_CHECK(cudaLaunchKernel((const void *)ent.kern, dim3((unsigned)ntiles), dim3(threads), args, smem, s));
    };
    if (job.prec) launch(double());
    else launch(float());
    {
        char nm[64];
        snprintf(nm, sizeof nm, "fft_spec_kernel<%s,n=%llu> (NVRTC)", job.prec ? "double" : "float", (unsigned long long)job.n);
        count_launch(nm);
    }
    RFB_CUDA_CHECK(cudaGetLastError());
    return true;
}

}  // namespace rfb
