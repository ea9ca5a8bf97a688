// Internal interface of the transform engine (host side).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "modes.h"
#include "plan.h"

namespace rfb {

struct Dim {
    int64_t n;    // extent
    int64_t is;   // input stride, bytes
    int64_t os;   // output stride, bytes
    bool tw;      // this dim's coordinate drives the four-step factor
};

// One batched 1-D complex DFT over strided memory.
struct LineJob {
    int prec = 0;        // 0: float, 1: double
    uint64_t n = 0;      // transform length
    int64_t is = 0, os = 0;
    std::vector<Dim> batch;
    const char *in = nullptr;
    char *out = nullptr;
    bool backward = false;
    double fct = 1.0;
    int load_mode = LD_C2C, store_mode = ST_C2C;
    int flags = 0;
    uint64_t n_in = 0;   // input elements present per line (0: n); the rest is zero padding
    uint64_t twN = 0;    // four-step factor exp(-2 pi i c k / twN) on the output (0: none)
    // scatter: output bin k goes to split_out[k / split_blk] (a base pointer already adjusted so that
    // the usual offset arithmetic with the full index k applies); empty: everything goes to `out`
    std::vector<char *> split_out;
    uint64_t split_blk = 0;
    // element-wise factors fused into the load / store (Bluestein): with g = e * g_mul + c * c_mul, c the
    // coordinate along the batch dim flagged `tw` (0 if none),
    //   load : x[e] * pre_tab[g] if g < pre_bound else 0          (pre_swap: swap re/im of x first)
    //   store: bin skipped if g >= post_bound, else value * post_tab[g]   (post_swap: swap re/im last)
    const void *pre_tab = nullptr, *post_tab = nullptr;
    uint64_t pre_bound = 0, post_bound = 0, g_mul = 1, c_mul = 1;
    bool pre_swap = false, post_swap = false;
    // circular-convolution line (power-of-two kernel only, contiguous lines): forward transform, times pre_tab[g]
    // with g indexed by the BIN, backward transform; twN / fct apply to the final store
    bool conv = false;
    // tables of the DCT / DST load / store modes LD_G_* / ST_G_* (generic tile kernel only)
    const void *aux_ld = nullptr, *aux_st = nullptr;
};

void run_lines(const LineJob &job, cudaStream_t stream);
// Only the register-resident power-of-two kernel; false (nothing launched) if it does not take the job.
bool run_lines_pow2(const LineJob &job, cudaStream_t stream);
// Only the generic shared-memory tile kernel (one launch); false (nothing launched) if a line does not fit.
bool run_lines_tile(const LineJob &job, cudaStream_t stream);

// N-D array description as it arrives through the ABI.
struct NdArgs {
    int prec;
    std::vector<int64_t> shape, sin, sout;
    std::vector<uint64_t> axes;
    const char *in;
    char *out;
    double fct;
};

void op_c2c(const NdArgs &a, bool forward, cudaStream_t s);
// c2c along ONE axis whose output index is cut into parts.size() equal blocks, block h written to
// parts[h] (arrays of shape `shape` with the axis extent divided by the number of parts)
void op_c2c_scatter(const NdArgs &a, size_t axis, bool forward, const std::vector<char *> &parts, cudaStream_t s);
void op_r2c(const NdArgs &a, bool forward, cudaStream_t s);      // shape = real input shape
void op_c2r(const NdArgs &a, bool forward, cudaStream_t s);      // shape = real output shape
void op_c2c_sym(const NdArgs &a, bool forward, cudaStream_t s);
// zero-padded / cropped input fused into the first load (shape_in: the array as it is; a.shape: what is transformed)
void op_c2c_pad(const NdArgs &a, const std::vector<int64_t> &shape_in, bool forward, cudaStream_t s);
void op_r2c_pad(const NdArgs &a, const std::vector<int64_t> &shape_in, bool forward, cudaStream_t s);
void op_c2r_pad(const NdArgs &a, const std::vector<int64_t> &shape_in, bool forward, cudaStream_t s);
// out[(i + shift) mod n] = in[i] along every dim (fftshift / ifftshift / roll); item: 4, 8 or 16 bytes
void op_roll(int64_t item, const std::vector<int64_t> &shape, const std::vector<int64_t> &sin,
             const std::vector<int64_t> &sout, const std::vector<int64_t> &shift, const char *in, char *out, cudaStream_t s);
void op_dcst(const NdArgs &a, int type, bool ortho, bool cosine, cudaStream_t s, int quirk_mode = -1);
void op_fftpack(const NdArgs &a, bool r2h, bool forward, cudaStream_t s);
void op_separable_hartley(const NdArgs &a, cudaStream_t s);
void op_genuine_hartley(const NdArgs &a, cudaStream_t s);

// data[l][j] *= table[j] over `nlines` contiguous lines of n items (real or complex, same precision as the table)
void op_scale_lines(int prec, bool cplx, uint64_t nlines, uint64_t n, const void *table, void *data, cudaStream_t s);

// sticky report of a kernel-side failure (fused4v2_launch.cu): throws Error if a kernel flagged one since the last check
void check_async_error();
uint64_t launch_count();
void launch_trace_enable(bool on);
std::string launch_trace_get();
void launch_count_reset();
void set_dst_ortho_quirk(bool on);

// staged copies of pageable host memory (staging.cu)
class Stager {
  public:
    Stager();
    ~Stager();
    void upload(char *dst_dev, const char *src_host, size_t n, cudaStream_t s);    // returns when the host data has been read
    void download(char *dst_host, const char *src_dev, size_t n, cudaStream_t s);  // asynchronous: finish() waits for it
    void finish();
    struct Impl;

  private:
    Impl *impl_;
};
bool host_memory_is_pageable(const void *p);

// stream-ordered scratch
struct Scratch {
    void *p = nullptr;
    cudaStream_t s = nullptr;
    Scratch(size_t bytes, cudaStream_t st);
    ~Scratch();
    Scratch(const Scratch &) = delete;
    Scratch &operator=(const Scratch &) = delete;
};

}  // namespace rfb
