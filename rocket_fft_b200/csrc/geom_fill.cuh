// Host-side construction of the kernel geometry shared by the tile kernel and the
// power-of-two kernel launchers.
#pragma once
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "engine.h"
#include "tile_kernel.cuh"

namespace rfb {

void count_launch(const char *name = nullptr);

template <typename T>
inline uint64_t fill_geom(TileGeom<T> &g, const LineJob &job, const std::vector<Dim> &dims, uint32_t W, bool load_lf,
                          bool store_lf) {
    memset(&g, 0, sizeof(g));
    const uint32_t n = (uint32_t)job.n;
    g.n = n;
    g.W = W;
    g.d_n = make_fastdiv(n);
    g.d_W = make_fastdiv(W);
    g.load_line_fast = load_lf ? 1 : 0;
    g.store_line_fast = store_lf ? 1 : 0;
    g.load_mode = job.load_mode;
    g.store_mode = job.store_mode;
    g.flags = job.flags;
    g.n_in = (uint32_t)(job.n_in ? job.n_in : job.n);
    g.n_out = (job.store_mode == ST_HALF) ? n / 2 + 1 : n;
    g.d_nout = make_fastdiv(g.n_out);
    g.backward = job.backward ? 1 : 0;
    g.in_sa = job.is;
    g.out_sa = job.os;
    g.tw_dim = -1;
    g.c_dim = -1;
    for (int d = 0; d < MAXB; ++d) {
        if (d < (int)dims.size()) {
            g.bext[d] = (uint32_t)dims[d].n;
            g.in_bs[d] = dims[d].is;
            g.out_bs[d] = dims[d].os;
            if (dims[d].tw) g.c_dim = d;
            if (dims[d].tw && job.twN) g.tw_dim = d;
        } else {
            g.bext[d] = 1;
            g.in_bs[d] = 0;
            g.out_bs[d] = 0;
        }
    }
    const uint32_t tiles0 = (g.bext[0] + W - 1) / W;
    g.d_t0 = make_fastdiv(tiles0);
    g.d_e1 = make_fastdiv(g.bext[1]);
    const uint64_t ntiles = (uint64_t)tiles0 * g.bext[1] * g.bext[2];
    if (ntiles >= (1ull << 31)) { set_error("too many tiles in one launch"); throw Error(); }
    g.in = job.in;
    g.out = job.out;
    g.g_mul = (uint32_t)job.g_mul;
    g.c_mul = (uint32_t)job.c_mul;
    g.pre_tab = (const cx<T> *)job.pre_tab;
    g.post_tab = (const cx<T> *)job.post_tab;
    g.pre_bound = (uint32_t)job.pre_bound;
    g.post_bound = (uint32_t)job.post_bound;
    g.aux_ld = (const cx<T> *)job.aux_ld;
    g.aux_st = (const cx<T> *)job.aux_st;
    g.pre_swap = job.pre_swap ? 1 : 0;
    g.post_swap = job.post_swap ? 1 : 0;
    if (!job.split_out.empty()) {
        if (job.split_out.size() > 16) { set_error("at most 16 scatter destinations"); throw Error(); }
        g.split_blk = (uint32_t)job.split_blk;
        g.d_split = make_fastdiv(g.split_blk);
        for (size_t i = 0; i < job.split_out.size(); ++i) g.split_base[i] = job.split_out[i];
    }
    g.fct = (T)job.fct;
    if (g.tw_dim >= 0) {
        const uint32_t S = split_size(job.twN);
        g.d_twS = make_fastdiv(S);
        g.twA = (const cx<T> *)get_table(TAB_SPLIT_A, job.prec, job.twN, S);
        g.twB = (const cx<T> *)get_table(TAB_SPLIT_B, job.prec, job.twN, S);
    }
    return ntiles;
}

// L2 prefetch of a later tile's input (prefetch_later_tile in tile_kernel.cuh): for element-fast tiles whose W lines of
// `n_items` items of `item` bytes form one (nearly) dense run of bytes.  RFB200_PF = distance in tiles, 0 = off.
template <typename T>
inline void set_prefetch(TileGeom<T> &g, const LineJob &job, const std::vector<Dim> &dims, uint32_t W, int64_t item,
                         uint64_t n_items) {
    static const int pf = [] {
        const char *v = getenv("RFB200_PF");
        return v ? atoi(v) : 148;
    }();
    if (pf <= 0 || job.is != item || job.pre_tab) return;
    const uint64_t line_bytes = n_items * (uint64_t)item;
    uint64_t span = line_bytes;
    if (W > 1) {
        if (dims.empty() || dims[0].is < (int64_t)line_bytes) return;
        span = (uint64_t)(W - 1) * (uint64_t)dims[0].is + line_bytes;
        if (span > (uint64_t)W * line_bytes * 5 / 4) return;
    }
    if (span < 512 || span > (1u << 20)) return;
    g.pf_dist = (uint32_t)pf;
    g.pf_bytes = (uint32_t)span;
}

// line-fast tiles (lines strided, W neighbours adjacent): one prefetch per row of W items.
// RFB200_PF_LF = distance in tiles (default 32), 0 = off.
template <typename T>
inline void set_prefetch_rows(TileGeom<T> &g, const LineJob &job, const std::vector<Dim> &dims, uint32_t W, int64_t item,
                              uint64_t n_items) {
    static const int pf = [] {
        const char *v = getenv("RFB200_PF_LF");
        return v ? atoi(v) : 32;
    }();
    // rows further apart than 64 KiB each sit on their own page: the prefetches then cost more (TLB) than they
    // save (measured: 1024^3 c64 axis 1, rows 8 KiB apart: 58 % -> 68 %; axis 0, rows 8 MiB apart: 57 % -> 47 %)
    const int64_t sa = job.is < 0 ? -job.is : job.is;
    const int64_t max_stride = 65536;
    if (pf <= 0 || dims.empty() || dims[0].is != item || n_items > 65536 || sa > max_stride) return;
    // zero-padded load (Bluestein): rows whose first element is already past the bound are never read
    if (job.pre_tab && job.g_mul) n_items = std::min<uint64_t>(n_items, (job.pre_bound + job.g_mul - 1) / job.g_mul);
    g.pf_dist = (uint32_t)pf;
    g.pf_rows = (uint32_t)n_items;
    g.pf_bytes = (uint32_t)(W * (uint64_t)item);
}

// the same, with the memory layout of a line derived from the job's load mode (complex line of n points)
template <typename T>
inline void set_prefetch_by_mode(TileGeom<T> &g, const LineJob &job, const std::vector<Dim> &dims, uint32_t W) {
    const uint64_t n_in = job.n_in ? job.n_in : job.n;
    if (g.load_line_fast) {
        if (job.load_mode == LD_C2C) set_prefetch_rows<T>(g, job, dims, W, sizeof(cx<T>), n_in);
        else if (job.load_mode == LD_REAL) set_prefetch_rows<T>(g, job, dims, W, sizeof(T), n_in);
        return;
    }
    switch (job.load_mode) {
        case LD_C2C: set_prefetch<T>(g, job, dims, W, sizeof(cx<T>), n_in); break;
        case LD_REAL: set_prefetch<T>(g, job, dims, W, sizeof(T), n_in); break;
        case LD_HERM: set_prefetch<T>(g, job, dims, W, sizeof(cx<T>), std::min<uint64_t>(n_in, job.n / 2 + 1)); break;
        case LD_HC: set_prefetch<T>(g, job, dims, W, sizeof(T), job.n); break;
        default: break;
    }
}

// pow2_launch_*.cu: returns false when the job is not one the register kernel takes
bool launch_pow2_f32(const LineJob &job, const std::vector<Dim> &dims, bool load_lf, bool store_lf, cudaStream_t s);
bool launch_pow2_f64(const LineJob &job, const std::vector<Dim> &dims, bool load_lf, bool store_lf, cudaStream_t s);
// fused4v2_launch.cu: both four-step passes of 16384-point strided complex64 lines in one warp-specialised persistent kernel
bool launch_fourstep_fused2_f32(const LineJob &job, const std::vector<Dim> &dims, cudaStream_t s);
// pow2_stream_launch.cu: strided 512 / 1024-point complex64 lines, persistent kernel fed by the copy engine
bool launch_pow2_stream_f32(const LineJob &job, const std::vector<Dim> &dims, cudaStream_t s);
// jit.cu: run-time specialised kernel for smooth non-power-of-two lengths (NVRTC); false if not taken
bool launch_spec_jit(const LineJob &job, const std::vector<Dim> &dims, bool load_lf, bool store_lf, bool aligned,
                     cudaStream_t s);
// regmix_launch.cu: smooth lengths (prime factors <= 13) that fit a register tile
bool launch_regmix(const LineJob &job, const std::vector<Dim> &dims, bool load_lf, bool store_lf, bool aligned,
                   cudaStream_t s);

}  // namespace rfb
