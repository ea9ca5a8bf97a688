#include "plan.h"

#include <math.h>

#include <stdlib.h>

#include <condition_variable>
#include <map>
#include <mutex>
#include <tuple>

namespace rfb {

// ---------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------
static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
const char *last_error() { return g_err.c_str(); }
void clear_error() { g_err.clear(); }

// ---------------------------------------------------------------------------------------
// good_size: smallest {2,3,5,7,11}-smooth (complex) / {2,3,5}-smooth (real) integer >= target.
// Same function of (target, real) as numba_good_size in the reference
// (_pocketfft_numba.cpp:25-29 -> _pocketfft_hdronly.h:589-650); small targets are returned
// unchanged there (<= 12 complex, <= 6 real) and therefore here.  Depth-first enumeration of
// smooth numbers with branch-and-bound on the best candidate so far.
// ---------------------------------------------------------------------------------------
namespace {
struct SmoothSearch {
    unsigned __int128 target, best;
    const uint32_t *primes;
    int nprimes;
    void go(int idx, unsigned __int128 val) {
        if (val >= target) {
            if (val < best) best = val;
            return;
        }
        if (idx == nprimes) return;
        for (unsigned __int128 v = val; v < best; v *= primes[idx]) go(idx + 1, v);
    }
};
}  // namespace

uint64_t good_size(uint64_t target, bool real) {
    static const uint32_t pc[5] = {11, 7, 5, 3, 2};
    static const uint32_t pr[3] = {5, 3, 2};
    if (target <= (real ? 6u : 12u)) return target;
    SmoothSearch s;
    s.target = target;
    s.best = 1;
    while (s.best < s.target) s.best *= 2;
    s.primes = real ? pr : pc;
    s.nprimes = real ? 3 : 5;
    s.go(0, 1);
    return (uint64_t)s.best;
}

// ---------------------------------------------------------------------------------------
// factorisation
// ---------------------------------------------------------------------------------------
std::vector<uint64_t> prime_factors(uint64_t n) {
    std::vector<uint64_t> f;
    while (n > 1 && (n & 1) == 0) { f.push_back(2); n >>= 1; }
    for (uint64_t p = 3; p * p <= n; p += 2)
        while (n % p == 0) { f.push_back(p); n /= p; }
    if (n > 1) f.push_back(n);
    return f;
}

uint64_t largest_prime_factor(uint64_t n) {
    auto f = prime_factors(n);
    return f.empty() ? 1 : f.back();
}

std::vector<uint32_t> radix_schedule(uint64_t n, uint32_t rmax) {
    std::vector<uint32_t> sched;
    auto f = prime_factors(n);
    int e2 = 0;
    std::vector<uint32_t> odd;
    for (auto p : f) {
        if (p == 2) ++e2;
        else {
            if (p > rmax) return {};
            odd.push_back((uint32_t)p);
        }
    }
    // group the twos: as many 16s as possible, remainder folded into 8/4 where that
    // keeps every pass at radix >= 4
    int n16 = e2 / 4, rem = e2 % 4;
    if (rem == 1 && n16 > 0) { --n16; for (int i = 0; i < n16; ++i) sched.push_back(16); sched.push_back(8); sched.push_back(4); }
    else {
        for (int i = 0; i < n16; ++i) sched.push_back(16);
        if (rem == 1) sched.push_back(2);
        if (rem == 2) sched.push_back(4);
        if (rem == 3) sched.push_back(8);
    }
    // odd primes ascending, except that the largest one leads: the first pass reads global memory
    // (more independent loads per butterfly) and the last one writes it (fewer digit reversals)
    if (odd.size() > 2) {
        sched.push_back(odd.back());
        odd.pop_back();
    }
    for (auto p : odd) sched.push_back(p);
    if (sched.empty()) sched.push_back(1);  // n == 1
    return sched;
}

namespace {
// fewest radices (all <= cap, from the kernel's set) whose product is n; non-increasing search
void regmix_search(uint64_t n, uint32_t maxr, const uint32_t *rad, int nrad, std::vector<uint32_t> &cur,
                   std::vector<uint32_t> &best) {
    if (n == 1) {
        if (best.empty() || cur.size() < best.size()) best = cur;
        return;
    }
    if (!best.empty() && cur.size() + 1 >= best.size()) return;
    for (int i = 0; i < nrad; ++i) {
        const uint32_t R = rad[i];
        if (R > maxr || n % R) continue;
        cur.push_back(R);
        regmix_search(n / R, R, rad, nrad, cur, best);
        cur.pop_back();
    }
}
}  // namespace

std::vector<uint32_t> regmix_schedule(uint64_t n, uint32_t cap) {
    static const uint32_t all[14] = {16, 15, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2};
    for (auto p : prime_factors(n))
        if (p > 13 || p > cap) return {};
    std::vector<uint32_t> cur, best;
    regmix_search(n, cap, all, 14, cur, best);
    if (best.empty()) return {};
    // the kernel runs the radices in this fixed order (powers of two first: their reads have the
    // longest unit-stride runs; odd radices last: conflict-free at small inner strides)
    static const uint32_t order[14] = {16, 8, 4, 2, 12, 10, 6, 15, 13, 11, 9, 7, 5, 3};
    std::vector<uint32_t> sched;
    for (auto R : order)
        for (auto r : best)
            if (r == R) sched.push_back(R);
    return sched;
}

// ---------------------------------------------------------------------------------------
// exact trigonometry
// ---------------------------------------------------------------------------------------
void sincos_2pi(uint64_t num, uint64_t den, long double &c, long double &s) {
    static const long double PI_4 = 0.78539816339744830961566084581987572L;
    num %= den;
    unsigned __int128 a = (unsigned __int128)num * 8u;
    uint64_t oct = (uint64_t)(a / den);
    uint64_t rem = (uint64_t)(a - (unsigned __int128)oct * den);
    long double bc, bs;
    uint64_t quarter;
    if ((oct & 1) == 0) {
        long double th = PI_4 * ((long double)rem / (long double)den);
        bc = cosl(th); bs = sinl(th);
        quarter = oct / 2;
    } else {
        long double th = PI_4 * ((long double)(den - rem) / (long double)den);
        bc = cosl(th); bs = -sinl(th);
        quarter = (oct + 1) / 2;
    }
    switch (quarter & 3) {
        case 0: c = bc; s = bs; break;
        case 1: c = -bs; s = bc; break;
        case 2: c = -bc; s = -bs; break;
        default: c = bs; s = -bc; break;
    }
}

uint32_t split_size(uint64_t n) {
    uint32_t s = 1;
    while ((uint64_t)s * s < n) s <<= 1;
    return s;
}

// ---------------------------------------------------------------------------------------
// device table cache
// ---------------------------------------------------------------------------------------
namespace {
using Key = TableKey;
struct KeyLess {
    bool operator()(const Key &a, const Key &o) const {
        return std::tie(a.dev, a.kind, a.prec, a.n, a.param) < std::tie(o.dev, o.kind, o.prec, o.n, o.param);
    }
};
// An entry is BUILDING while one thread computes and uploads it (outside the cache mutex; other threads asking for the same
// table wait on the condition variable), then READY.  `leases` counts the TableScope's (one per library call in flight on
// the host) that have fetched it: such an entry is never evicted, so a pointer handed to a call stays valid until that
// call has enqueued its kernels; a table evicted later is released with cudaFree, which waits for the device to finish the
// kernels that may still read it.
struct Entry {
    void *d = nullptr;
    size_t bytes = 0;
    uint64_t last_use = 0;
    int leases = 0;
    bool ready = false;
};
std::mutex g_mu;
std::condition_variable g_cv;
std::map<Key, Entry, KeyLess> g_tables;
uint64_t g_clock = 0;
size_t g_bytes = 0;
thread_local std::vector<Key> *g_scope = nullptr;  // tables leased by the library call running on this thread

size_t cache_max_entries() {
    static const size_t v = [] { const char *e = getenv("RFB200_PLAN_CACHE_ENTRIES"); return e ? (size_t)atoll(e) : (size_t)256; }();
    return v;
}
size_t cache_max_bytes() {
    static const size_t v = [] { const char *e = getenv("RFB200_PLAN_CACHE_MB"); return (e ? (size_t)atoll(e) : (size_t)1024) << 20; }();
    return v;
}

// g_mu held.  Least recently used first; entries in use (leased or being built) stay.
void evict_locked(std::vector<void *> &to_free) {
    while (g_tables.size() > cache_max_entries() || g_bytes > cache_max_bytes()) {
        auto victim = g_tables.end();
        for (auto it = g_tables.begin(); it != g_tables.end(); ++it)
            if (it->second.ready && it->second.leases == 0 && (victim == g_tables.end() || it->second.last_use < victim->second.last_use))
                victim = it;
        if (victim == g_tables.end()) break;
        to_free.push_back(victim->second.d);
        g_bytes -= victim->second.bytes;
        g_tables.erase(victim);
    }
}

template <typename T>
void fill(std::vector<T> &h, TableKind kind, uint64_t n, uint64_t param, size_t count) {
    h.resize(2 * count);
    for (size_t t = 0; t < count; ++t) {
        long double c, s;
        switch (kind) {
            case TAB_LINE:
            case TAB_SPLIT_B: sincos_2pi(t, n, c, s); break;
            case TAB_SPLIT_A: sincos_2pi((uint64_t)(((unsigned __int128)t * param) % n), n, c, s); break;
            case TAB_CHIRP: {
                // exp(-i pi t^2 / n) = exp(-2 pi i (t^2 mod 2n) / (2n)), exact integer reduction
                uint64_t m = (uint64_t)(((unsigned __int128)t * t) % (2 * (unsigned __int128)n));
                sincos_2pi(m, 2 * n, c, s);
                break;
            }
            case TAB_QUARTER: sincos_2pi(t, 4 * n, c, s); break;
            default: c = 0; s = 0; break;
        }
        h[2 * t] = (T)c;
        h[2 * t + 1] = (T)(-s);
    }
}
// Pass-major, q-major twiddles of the register-resident power-of-two kernel: pass p with
// radix R_p and inner stride ido_p > 1 owns (R_p-1)*ido_p entries exp(-2 pi i i q/(R_p ido_p))
// stored at [(q-1)*ido_p + i].  Pass structure as in P2<LOGN> (pow2_kernel.cuh): one radix
// 2/4/8 pass first when log2 n is not a multiple of 4, then radix-16 passes.
template <typename T>
void fill_stockham(std::vector<T> &h, uint64_t n, int which, uint64_t param) {
    std::vector<uint32_t> rad;
    if (which == 1) rad = radix_schedule(n, 64);
    else if (which == 2) rad = regmix_schedule(n, (uint32_t)param);
    else {
        int logn = 0;
        while ((1ull << logn) < n) ++logn;
        if (logn % 4) rad.push_back(1u << (logn % 4));
        for (int i = 0; i < logn / 4; ++i) rad.push_back(16);
    }
    h.clear();
    uint64_t l1 = 1;
    for (auto R : rad) {
        const uint64_t ido = n / (l1 * R);
        if (ido > 1)
            for (uint32_t q = 1; q < R; ++q)
                for (uint64_t i = 0; i < ido; ++i) {
                    long double c, s;
                    sincos_2pi(i * q, R * ido, c, s);
                    h.push_back((T)c);
                    h.push_back((T)(-s));
                }
        l1 *= R;
    }
    if (h.empty()) { h.push_back(T(1)); h.push_back(T(0)); }
}
}  // namespace

TableScope::TableScope() : prev_(g_scope) { g_scope = &keys_; }
TableScope::~TableScope() {
    g_scope = prev_;
    if (keys_.empty()) return;
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto &k : keys_) {
        auto it = g_tables.find(k);
        if (it != g_tables.end() && it->second.leases > 0) --it->second.leases;
    }
}

const void *get_table(TableKind kind, int prec, uint64_t n, uint64_t param, bool *created, void **writable) {
    int dev = 0;
    RFB_CUDA_CHECK(cudaGetDevice(&dev));
    Key key{dev, (int)kind, prec, n, param};
    if (created) *created = false;
    {
        std::unique_lock<std::mutex> lk(g_mu);
        for (;;) {
            auto it = g_tables.find(key);
            if (it == g_tables.end()) break;
            if (!it->second.ready) { g_cv.wait(lk); continue; }  // another thread is building it
            it->second.last_use = ++g_clock;
            if (g_scope) { ++it->second.leases; g_scope->push_back(key); }
            if (writable) *writable = it->second.d;
            return it->second.d;
        }
        g_tables[key] = Entry();  // BUILDING: claimed by this thread
    }
    // ---- build and upload outside the mutex ------------------------------------------------------------------------
    void *d = nullptr;
    size_t bytes = 0;
    try {
        size_t count = 0;
        switch (kind) {
            case TAB_LINE: count = n; break;
            case TAB_SPLIT_A: count = (n + param - 1) / param + 1; break;
            case TAB_SPLIT_B: count = param; break;
            case TAB_CHIRP: count = n; break;
            case TAB_CHIRP_FFT:
            case TAB_CHIRP_FFT_T: count = param; break;
            case TAB_QUARTER: count = n + 1; break;
            case TAB_STOCKHAM:
            case TAB_REGMIX:
            case TAB_TILE: count = n; break;  // upper bound: sum (R-1)*ido < n
        }
        const size_t esz = prec ? 16 : 8;
        bytes = count * esz > 0 ? count * esz : esz;
        RFB_CUDA_CHECK(cudaMalloc(&d, bytes));
        if (kind == TAB_STOCKHAM || kind == TAB_TILE || kind == TAB_REGMIX) {
            if (prec) {
                std::vector<double> h;
                fill_stockham(h, n, kind == TAB_TILE ? 1 : (kind == TAB_REGMIX ? 2 : 0), param);
                RFB_CUDA_CHECK(cudaMemcpy(d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
            } else {
                std::vector<float> h;
                fill_stockham(h, n, kind == TAB_TILE ? 1 : (kind == TAB_REGMIX ? 2 : 0), param);
                RFB_CUDA_CHECK(cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
            }
        } else if (kind != TAB_CHIRP_FFT && kind != TAB_CHIRP_FFT_T) {
            if (prec) {
                std::vector<double> h;
                fill(h, kind, n, param, count);
                RFB_CUDA_CHECK(cudaMemcpy(d, h.data(), count * esz, cudaMemcpyHostToDevice));
            } else {
                std::vector<float> h;
                fill(h, kind, n, param, count);
                RFB_CUDA_CHECK(cudaMemcpy(d, h.data(), count * esz, cudaMemcpyHostToDevice));
            }
        }
    } catch (...) {
        if (d) cudaFree(d);
        {
            std::lock_guard<std::mutex> lk(g_mu);
            g_tables.erase(key);
        }
        g_cv.notify_all();
        throw;
    }
    std::vector<void *> to_free;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        Entry &e = g_tables[key];
        e.d = d;
        e.bytes = bytes;
        e.ready = true;
        e.last_use = ++g_clock;
        if (g_scope) { ++e.leases; g_scope->push_back(key); }
        else e.leases = 0;
        g_bytes += bytes;
        const bool keep = !g_scope;  // no scope: protect the new entry from its own eviction pass
        if (keep) ++e.leases;
        evict_locked(to_free);
        if (keep) --g_tables[key].leases;
    }
    g_cv.notify_all();
    for (void *q : to_free) cudaFree(q);  // implicit device synchronisation: no kernel is still reading it afterwards
    if (created) *created = true;
    if (writable) *writable = d;
    return d;
}

// A table whose contents the engine fills itself (TAB_CHIRP_FFT*) and whose fill failed: forget it.
void discard_table(TableKind kind, int prec, uint64_t n, uint64_t param) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    Key key{dev, (int)kind, prec, n, param};
    void *d = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_tables.find(key);
        if (it == g_tables.end() || !it->second.ready) return;
        d = it->second.d;
        g_bytes -= it->second.bytes;
        g_tables.erase(it);
    }
    if (d) cudaFree(d);
}

void plan_cache_clear() {
    std::vector<void *> to_free;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        int cur = 0;
        cudaGetDevice(&cur);
        for (auto it = g_tables.begin(); it != g_tables.end();) {
            if (it->first.dev == cur && it->second.ready && it->second.leases == 0) {
                to_free.push_back(it->second.d);
                g_bytes -= it->second.bytes;
                it = g_tables.erase(it);
            } else ++it;
        }
    }
    for (void *q : to_free) cudaFree(q);
}

void plan_cache_stats(uint64_t *entries, uint64_t *bytes) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (entries) *entries = g_tables.size();
    if (bytes) *bytes = g_bytes;
}

// Drops the host-side records without touching the device (the forked child of a process that used CUDA must not).
void plan_cache_forget() {
    g_tables.clear();
    g_bytes = 0;
}

}  // namespace rfb
