// Host-side construction of the kernel geometry shared by the tile kernel and the
// power-of-two kernel launchers.
#pragma once
#include <string.h>

#include <vector>

#include "engine.h"
#include "tile_kernel.cuh"

namespace rfb {

void count_launch();

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
    g.pre_tab = (const cx<T> *)job.pre_tab;
    g.post_tab = (const cx<T> *)job.post_tab;
    g.pre_bound = (uint32_t)job.pre_bound;
    g.post_bound = (uint32_t)job.post_bound;
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

// pow2_launch_*.cu: returns false when the job is not one the register kernel takes
bool launch_pow2_f32(const LineJob &job, const std::vector<Dim> &dims, bool load_lf, bool store_lf, cudaStream_t s);
bool launch_pow2_f64(const LineJob &job, const std::vector<Dim> &dims, bool load_lf, bool store_lf, cudaStream_t s);
// jit.cu: run-time specialised kernel for smooth non-power-of-two lengths (NVRTC); false if not taken
bool launch_spec_jit(const LineJob &job, const std::vector<Dim> &dims, bool load_lf, bool store_lf, bool aligned,
                     cudaStream_t s);
// regmix_launch.cu: smooth lengths (prime factors <= 13) that fit a register tile
bool launch_regmix(const LineJob &job, const std::vector<Dim> &dims, bool load_lf, bool store_lf, bool aligned,
                   cudaStream_t s);

}  // namespace rfb
