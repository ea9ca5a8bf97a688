// Strided power-of-two lines of 512 / 1024 complex64 points (the outer axes of a volume) as a PERSISTENT kernel fed by the copy
// engine, with two warp groups taking turns on one exchange buffer.
//
// The register kernel for such lines (pow2_pair_kernel.cuh, 16 lines x 1024 points = 128 KiB per tile) fills an SM with ONE
// CTA: its load, butterfly and store phases run one after the other and HBM idles while it computes (measured, 1024^3: axis 0
// 57 %, axis 1 68 % of the HBM copy peak; 8.6 - 10.5 us per tile against 5.8 us of HBM time).  Here
//   * one CTA per SM walks over its tiles; while tile i is being transformed the copy engine (TMA, cp.async.bulk.tensor,
//     boxes of 256 rows x 128 B, 128-byte swizzle) brings tile i + 1 into a landing buffer -- the load latency disappears
//     behind the butterflies;
//   * the CTA is two warp groups; group g owns the lines 8 g .. 8 g + 7 of the tile (two neighbouring lines per thread, as in
//     pow2_pair_kernel.cuh).  Both groups share ONE exchange buffer of half a tile and take turns on it (named barriers
//     A -> B -> A -> ...): while one group exchanges, the other computes its butterflies.  Landing buffer 128 KiB + exchange
//     buffer 68 KiB fit the 227 KiB of an SM;
//   * the landing buffer is free again as soon as both groups hold their points in registers (group B has passed its first
//     turn on the exchange buffer): its thread 0 then asks for the next tile.
// Thread mapping inside a group: lanes 0-3 = the 4 line pairs, then the butterfly index with bits 0 and 2 swapped -- a
// quarter warp (one 128-byte shared-memory wavefront of 16-byte accesses) then reads rows r and r + 4 of the swizzled landing
// buffer (different halves of the 128-byte bank space) and writes / reads exchange slots 4 apart of 4 lines whose pitch is odd:
// every shared-memory access of the kernel is conflict free.
// (Counterpart of general_nd + copy_input / copy_output for strided axes in the reference, _pocketfft_hdronly.h:3496-3607;
// different algorithm, new code.)
#pragma once
#include <cuda.h>

#include "fused4v2_kernel.cuh"  // mbarrier / TMA helpers (namespace f4v2)
#include "pow2_pair_kernel.cuh"

namespace rfb {

struct StreamParams {
    char *out;
    int64_t out_sa, out_bs1, out_bs2;  // byte strides: transform axis, batch dims 1 and 2 (dim 0 = neighbouring lines, 8 B apart)
    uint32_t bext0;                    // extent of dim 0
    uint32_t ntiles;
    FastDiv d_t0, d_e1;                // tile -> (t0, i1, i2)
    int backward;
    float fct;
};

template <int LOGN, int W>
struct StreamBody {
    using C = float2;
    using PL = P2<LOGN>;
    using PB = PairBody<LOGN, W>;
    static constexpr int N = PL::N, TPL = PL::TPL, WP = W / 2, WPG = WP / 2, GT = WPG * TPL, NT = 2 * GT;
    static constexpr int PITCH = (N + PL::PAD) | 1;  // exchange pitch in 16-byte pairs
    static constexpr int ROWB = W * 8;               // landing row: W lines x 8 B
    static constexpr int LBYTES = N * ROWB;
    static constexpr int XBYTES = ((WPG * PITCH * 16) + 127) & ~127;
    static constexpr int BOXROWS = N < 256 ? N : 256, NBOX = N / BOXROWS;
    static constexpr int SMEM = LBYTES + XBYTES + 128 + 1024;  // + barrier + alignment slack
    static_assert(W == 16 && WPG == 4, "the swizzled landing layout and the thread mapping assume 16 lines per tile");
    static_assert(TPL >= 8 && GT % 32 == 0, "a group is a whole number of warps");

    static __device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
    static __device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

    // One turn on the exchange buffer: registers (outputs of pass P-1) -> shared -> registers (inputs of pass P).
    // Barrier 1: "A has left the buffer" (A arrives, B waits); barrier 2: "B has left the buffer"; 3 + g: inside group g.
    template <int P>
    static __device__ __forceinline__ void exchange_turn(C *a, C *b, float4 *line, int t, int g, bool wait_other, bool signal_other) {
        constexpr int Rp = PL::radix(P - 1), NBp = 16 / Rp;
        constexpr int ido = PL::ido(P);
        if (wait_other) { if (g == 0) bar_sync(2, NT); else bar_sync(1, NT); }
#pragma unroll
        for (int j = 0; j < NBp; ++j)
#pragma unroll
            for (int q = 0; q < Rp; ++q)
                line[p2_phys<LOGN, P>(t + j * TPL + q * (N / Rp))] =
                    make_float4(a[j * Rp + q].x, a[j * Rp + q].y, b[j * Rp + q].x, b[j * Rp + q].y);
        if (g == 0) bar_sync(3, GT); else bar_sync(4, GT);
        const int i = t % ido, k = t / ido;
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const float4 u = line[p2_phys<LOGN, P>(i + ido * (m + 16 * k))];
            a[m] = make_float2(u.x, u.y);
            b[m] = make_float2(u.z, u.w);
        }
        if (signal_other) { if (g == 0) bar_arrive(1, NT); else bar_arrive(2, NT); }
    }
};

// Measured on B200, 1024^3 complex64 (profiles/r02n_*, r02o_*): axis 1 (rows 8 KiB apart) 3.83 -> 3.70 ms against the register
// kernel; axis 0 (rows 8 MiB apart, every row of a box on its own page) 4.62 -> 5.61 ms -- the copy engine then takes ~20
// cycles per row -- so the launcher keeps rows more than 64 KiB apart on the register kernel.  Fetching the next tile with
// 16-byte cp.async from all threads instead (same layout, completion counted on the same mbarrier) was slower on both axes
// (5.80 / 4.16 ms) and was removed: per tile the butterflies and exchanges take about as long as HBM needs for the tile's
// bytes, and one CTA of 16 warps per SM does not overlap the two well however the loads are issued.
template <int LOGN, int W>
__global__ void __launch_bounds__(StreamBody<LOGN, W>::NT, 512 / StreamBody<LOGN, W>::NT)
    fft_pow2_stream_kernel(const StreamParams p, const __grid_constant__ CUtensorMap map, const float2 *__restrict__ stw) {
    using SB = StreamBody<LOGN, W>;
    using PL = typename SB::PL;
    using PB = typename SB::PB;
    using C = float2;
    constexpr int N = SB::N, TPL = SB::TPL, WPG = SB::WPG, GT = SB::GT;
    extern __shared__ __align__(1024) unsigned char smem_strm[];
    unsigned char *L = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_strm) + 1023) & ~(uintptr_t)1023);
    float4 *X = reinterpret_cast<float4 *>(L + SB::LBYTES);
    uint64_t *full = reinterpret_cast<uint64_t *>(L + SB::LBYTES + SB::XBYTES);
    const int tid = threadIdx.x, g = tid / GT, lt = tid % GT, wpl = lt % WPG, tq = lt / WPG;
    const int t = (tq & ~5) | ((tq & 1) << 2) | ((tq >> 2) & 1);  // butterfly index: bits 0 and 2 of the lane order swapped
    const int wp = g * WPG + wpl;                                  // pair of lines 2 wp, 2 wp + 1 of the tile
    if (tid == 0) {
        f4v2::mbar_init(full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint64_t pol = l2_policy_stream();
    auto issue = [&](uint32_t tile) {
        uint32_t t0, i1, i2, rest;
        fdivmod(tile, p.d_t0, rest, t0);
        fdivmod(rest, p.d_e1, i2, i1);
        f4v2::mbar_arrive_tx(full, (uint32_t)SB::LBYTES);
#pragma unroll
        for (int bx = 0; bx < SB::NBOX; ++bx)
            f4v2::tma_load_4d(L + bx * SB::BOXROWS * SB::ROWB, &map, (int)(t0 * W), bx * SB::BOXROWS, (int)i1, (int)i2, full, pol);
    };
    const bool issuer = tid == GT;  // group B's first thread
    if (issuer && blockIdx.x < p.ntiles) issue(blockIdx.x);
    float4 *line = X + wpl * SB::PITCH;
    uint32_t it = 0;
    for (uint32_t tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
        const bool last_tile = tile + gridDim.x >= p.ntiles;
        uint32_t t0, i1, i2, rest;
        fdivmod(tile, p.d_t0, rest, t0);
        fdivmod(rest, p.d_e1, i2, i1);
        C a[16], b[16];
        // ---- pass 0 operands: landing buffer -> registers (rows t + j TPL + m ido; the pair's 16 bytes sit in chunk wp ^ (row & 7)) ----
        f4v2::mbar_wait(full, it & 1u);
        {
            constexpr int R = PL::radix(0), NB = 16 / R, ido = PL::ido(0);
#pragma unroll
            for (int j = 0; j < NB; ++j)
#pragma unroll
                for (int m = 0; m < R; ++m) {
                    const int r = t + j * TPL + m * ido;
                    const float4 u = *reinterpret_cast<const float4 *>(L + r * SB::ROWB + ((wp ^ (r & 7)) << 4));
                    a[j * R + m] = make_float2(u.x, u.y);
                    b[j * R + m] = make_float2(u.z, u.w);
                }
            if (p.backward) {
#pragma unroll
                for (int i = 0; i < 16; ++i) { a[i] = cswap(a[i]); b[i] = cswap(b[i]); }
            }
        }
        PB::template compute2<0>(a, b, t, stw);
        // ---- first turn on the exchange buffer; once group B is through, every thread of the CTA holds its points of this
        //      tile in registers: the landing buffer can take the next tile ------------------------------------------------------
        if (g == 0) SB::template exchange_turn<1>(a, b, line, t, 0, it != 0, true);
        else {
            SB::bar_sync(1, SB::NT);
            if (issuer && !last_tile) {
                f4v2::fence_proxy_async();
                issue(tile + gridDim.x);
            }
            SB::template exchange_turn<1>(a, b, line, t, 1, false, PL::NPASS > 2 || !last_tile);
        }
        PB::template compute2<1>(a, b, t, stw);
        if constexpr (PL::NPASS > 2) {
            SB::template exchange_turn<2>(a, b, line, t, g, true, g == 0 || PL::NPASS > 3 || !last_tile);
            PB::template compute2<2>(a, b, t, stw);
        }
        if constexpr (PL::NPASS > 3) {
            SB::template exchange_turn<3>(a, b, line, t, g, true, g == 0 || !last_tile);
            PB::template compute2<3>(a, b, t, stw);
        }
        // ---- thread t holds bins t + j TPL + q N / RL of both lines: 16-byte stores (the pair is adjacent in memory) ------------------
        const uint32_t w_first = t0 * (uint32_t)W;
        const int wvalid = (int)min((uint32_t)W, p.bext0 - w_first);
        const bool ok0 = 2 * wp < wvalid, ok1 = 2 * wp + 1 < wvalid;
        if (ok0) {
            constexpr int RL = PL::radix(PL::NPASS - 1), NBL = 16 / RL;
            const float f = p.fct;
            const bool bw = p.backward != 0;
            char *pj = p.out + (int64_t)(w_first + 2 * wp) * 8 + (int64_t)i1 * p.out_bs1 + (int64_t)i2 * p.out_bs2 + (int64_t)t * p.out_sa;
            const int64_t step_q = (int64_t)(N / RL) * p.out_sa, step_j = (int64_t)TPL * p.out_sa;
#pragma unroll
            for (int j = 0; j < NBL; ++j) {
                char *pq = pj;
#pragma unroll
                for (int q = 0; q < RL; ++q) {
                    C va = cscale(a[j * RL + q], f), vb = cscale(b[j * RL + q], f);
                    if (bw) { va = cswap(va); vb = cswap(vb); }
                    if (ok1) asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(pq), "f"(va.x), "f"(va.y), "f"(vb.x), "f"(vb.y) : "memory");
                    else *reinterpret_cast<C *>(pq) = va;
                    pq += step_q;
                }
                pj += step_j;
            }
        }
    }
}

}  // namespace rfb
