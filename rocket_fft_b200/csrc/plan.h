// Host-side planning: factorisation, good_size, exact twiddle generation and the per-device
// cache of device-resident tables.  (Counterpart of util::*, sincos_2pibyn, cfftp::factorize,
// get_plan in the reference: _pocketfft_hdronly.h:467-725, 1750-1829, 3169-3223.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace rfb {

uint64_t good_size(uint64_t target, bool real);

// Prime factorisation, ascending.
std::vector<uint64_t> prime_factors(uint64_t n);
uint64_t largest_prime_factor(uint64_t n);

// Radix schedule for an in-shared-memory line transform: products of 2s grouped into
// 16/8/4/2, then odd primes ascending.  Empty if some prime factor exceeds `rmax`.
std::vector<uint32_t> radix_schedule(uint64_t n, uint32_t rmax);
// Schedule of the register-resident mixed-radix kernel: the fewest radices from
// {16,15,13,12,11,10,9,8,7,6,5,4,3,2} (each <= cap) with product n, in the kernel's fixed order
// 16,8,4,2,12,10,6,15,13,11,9,7,5,3.  Empty if n has a prime factor > min(13, cap).
std::vector<uint32_t> regmix_schedule(uint64_t n, uint32_t cap);

// cos/sin of 2*pi*num/den with octant reduction in long double.
void sincos_2pi(uint64_t num, uint64_t den, long double &c, long double &s);

enum TableKind : int {
    TAB_LINE = 0,      // exp(-2 pi i t / n), t in [0, n)
    TAB_SPLIT_A = 1,   // exp(-2 pi i (t*S) / n), t in [0, ceil(n/S)]   (S = split)
    TAB_SPLIT_B = 2,   // exp(-2 pi i t / n), t in [0, S)
    TAB_CHIRP = 3,     // exp(-i pi t^2 / n), t in [0, n)
    TAB_CHIRP_FFT = 4, // forward DFT_M of the wrapped conjugate chirp, divided by M  (param = M)
    TAB_CHIRP_FFT_T = 9, // the same in four-step order: entry k1*n2 + k2 holds bin k2*n1 + k1  (param = M)
    TAB_QUARTER = 5,   // exp(-2 pi i t / (4 n)), t in [0, n]   (DCT/DST and real pre/post factors)
    TAB_REGMIX = 8,    // pass-major twiddles for regmix_schedule(n, param)  (param = radix cap)
    TAB_TILE = 7,      // pass-major twiddles of the generic tile kernel for radix_schedule(n, 64)
    TAB_STOCKHAM = 6,  // per-pass twiddles of the power-of-two register kernel (n = 2^k), see pow2_kernel.cuh
};

// Returns a device pointer to the table (complex<float> or complex<double> by prec),
// creating and uploading it on first use.  Thread-safe; tables live until
// plan_cache_clear().  TAB_CHIRP_FFT is filled in by the engine (it needs an FFT).
const void *get_table(TableKind kind, int prec, uint64_t n, uint64_t param, bool *created = nullptr,
                      void **writable = nullptr);
void plan_cache_clear();
// The cache is bounded (RFB200_PLAN_CACHE_ENTRIES, default 256 tables; RFB200_PLAN_CACHE_MB, default 1024): least recently
// used tables are released once a bound is exceeded (the reference keeps 16 plans, _pocketfft_hdronly.h:3169-3223).
void plan_cache_stats(uint64_t *entries, uint64_t *bytes);
void plan_cache_forget();
void discard_table(TableKind kind, int prec, uint64_t n, uint64_t param);
// One per library call (constructed at the ABI entry points): the tables fetched by the call on this thread are leased
// until the scope ends, i.e. they cannot be evicted between get_table and the kernel launches that use them.
struct TableKey {
    int dev, kind, prec;
    uint64_t n, param;
};
class TableScope {
  public:
    TableScope();
    ~TableScope();
    TableScope(const TableScope &) = delete;
    TableScope &operator=(const TableScope &) = delete;

    std::vector<TableKey> keys_;

  private:
    std::vector<TableKey> *prev_;
};

uint32_t split_size(uint64_t n);  // S used by the SPLIT tables: ceil(sqrt(n)) rounded up to a power of two

// error reporting (thread-local)
void set_error(const std::string &msg);
const char *last_error();
void clear_error();

#define RFB_CUDA_CHECK(expr)                                                                   \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            rfb::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));               \
            throw rfb::Error();                                                                \
        }                                                                                      \
    } while (0)

struct Error {};

}  // namespace rfb
