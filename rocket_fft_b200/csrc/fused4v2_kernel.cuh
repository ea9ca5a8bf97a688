// Four-step transform of long strided lines (n = 128 x 128 complex64 points, e.g. the 16384-point columns of a
// 16384 x 8193 half spectrum) as ONE persistent, warp-specialised kernel: both steps in a single launch, the
// intermediate in a ring of L2-resident scratch slots, every global <-> shared transfer done by the copy engine
// (TMA: cp.async.bulk.tensor boxes of 128 rows x 256 B, SASS UTMALDG / UTMASTG; cp.async.bulk / UBLKCP for the
// contiguous ring reads) under mbarrier control so that the next tile is in flight while the current one is being
// transformed.
//
//   strip  = 32 neighbouring lines (columns); its intermediate, 16384 rows x 256 B = 4 MiB, lives in ring slot (strip % ring)
//   A(s,j0): 128-point DFTs over the rows j1*128 + j0 of the strip, times w_16384^(j0 k1), stored to slot rows k1*128 + j0
//   B(s,k1): 128-point DFTs over the slot rows k1*128 + j0 (one contiguous 32 KiB block), stored to array rows k2*128 + k1
//   G adjacent strips form a group; a unit = one step of one group = 128 x G tiles in the order (j0 or k1) major, strip
//   minor, so that tiles in flight at the same time touch G x 256 contiguous bytes of every array row they share (DRAM page
//   locality: measured 0.65 -> see DESIGN.md).  Unit order (fuse4_decode_unit over groups): A(0) .. A(lag-1) | A(lag) B(0) |
//   A(lag+1) B(1) | ...  B(s) needs all 128 tiles of A(s) (counter doneA[s]); A(s) needs the slot's previous tenant
//   B(s - ring) read (doneB); lag and ring count groups, strips are numbered group * G + g (numbers past the array's last
//   strip are skipped).
//
// CTA = 4 transform warps + 1 copy warp, STAGES shared-memory stages of 34.25 KiB.  The copy warp draws tickets (an atomic
// counter: tiles are handed out in unit order, so every tile a CTA waits for has been taken by a running CTA earlier --
// no co-residency assumption, no deadlock), checks the tile's dependency, and issues its loads; when the transform
// warps have left a stage it stores the staged result of an A tile to the ring (bulk stores), and publishes the completion
// counters.  The transform warps never touch a global counter and never wait for global memory: they wait on the stage's
// `full` mbarrier, transform (two neighbouring lines per thread, as pow2_pair_kernel.cuh), and stage the result in the
// same shared-memory stage, from where one TMA box store takes it to the ring (A) or to the array (B).
// The array side works for rows that are only 8-byte aligned (row pitch 8193 x 8 B, which TMA's 16-byte rules forbid):
// even and odd rows get a tensor map each -- two rows apart the addresses ARE multiples of 16 bytes -- and a tile only
// ever touches rows of one parity (rows j1*128 + j0 with j0 fixed).  A parity class whose rows start 8 bytes past a
// 16-byte boundary is LOADED through a box that starts one element early (34 elements wide; measured: a box whose first
// byte is not 16-byte aligned raises an illegal-instruction fault) and read 8 bytes in; such rows are STORED from
// registers, since a wider box would overwrite the neighbouring strips' columns.  Partial strips cost nothing on the TMA
// paths: the box is clipped at the array's edge by the copy engine (zero fill on the way in, nothing written on the way out).
// (Counterpart of general_nd + copy_input/copy_output for strided axes in the reference,
// _pocketfft_hdronly.h:3496-3607; different algorithm, new code.)
#pragma once
#include <cuda.h>

#include "pow2_kernel.cuh"

namespace rfb {

struct F4v2Params {
    char *out;             // output array (rows stored from registers when their parity class is not 16-byte aligned)
    int64_t out_pitch, out_outer;
    uint32_t cols;         // neighbouring lines per outer item
    char *ring_mem;        // ring * 4 MiB of scratch
    uint32_t *ctr;         // [0] ticket, [1] error, [2 .. 2+S) doneA, [2+S .. 2+2S) doneB
    uint32_t *host_err;    // pinned, mapped: set when a dependency never completed
    uint32_t nstrips, spo; // strip numbers in total (ngroups * G) / real strips per outer item
    uint32_t ngroups, gpo; // groups of G adjacent strips in total / per outer item
    uint32_t glog;         // log2 G
    uint32_t ring, lag;    // in groups; ring slots = ring * G
    uint32_t total_items;  // 2 * ngroups * G * 128
    uint32_t mis_in[2], mis_out[2];  // per row-parity class: 1 if its rows start 8 bytes past a 16-byte boundary
    FastDiv d_gpo, d_ring;  // group -> (outer, group in outer); strip number -> ring slot (mod ring * G)
    int backward;
    float fct;
    const float2 *stw;       // Stockham twiddles of the 128-point line transform
    const float2 *twA, *twB; // exp(-2 pi i t / 16384) = twA[t / S] * twB[t % S]
    FastDiv d_twS;
    uint32_t max_idle;       // bound on the copy thread's idle polls (never hang the GPU)
    uint32_t dbg_only;       // measurement aid: 1 = step A tiles only, 2 = step B tiles only (no dependencies; wrong results)
};

// tensor maps (kernel parameters): [0] / [1] input rows of even / odd parity, [2] / [3] output rows, [4] ring (A's store box)
struct F4v2Maps {
    CUtensorMap m[5];
};

namespace f4v2 {

constexpr int N = 128, W = 32, WP = 16, TPL = 8, NCOMP = 128, NTHREADS = 160;
constexpr int XPITCH = 137;                    // exchange: float4 per pair of lines ((128 + 8) | 1)
constexpr int STAGE_BYTES = WP * XPITCH * 16;  // 35072 = 274 x 128: holds the 32 KiB tile and the padded exchange
constexpr int ROW = 256;                       // tile row: 32 lines x 8 B
constexpr int TILE_BYTES = N * ROW;
constexpr int ROW_MIS = 272;                   // landing row of a misaligned class: 34 elements, the tile starts 8 bytes in
constexpr uint32_t KIND_A = 0, KIND_B = 1, KIND_STOP = 2;
constexpr uint32_t FLAG_SHIFT = 1, FLAG_STG = 2;  // tile landed 8 bytes into 272-byte rows / result stored from registers
constexpr uint32_t FLAG_AREG = 4;                 // A tile stored to the ring from registers and published by the transform warps

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t *b, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    } while (!ok);
}
// global -> shared: one box of a 4-D tensor map, completion counted in bytes on an mbarrier
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint64_t *bar, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
// shared -> global: one box, tracked by the issuing thread's bulk async-groups
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *map, int c0, int c1, int c2, int c3, const void *src, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2, %3, %4}], [%5], %6;"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(src)), "l"(pol) : "memory");
}
// L2 prefetch of one box
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap *map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// global -> shared contiguous bulk copy
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int NPEND> __device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NPEND) : "memory");
}
template <int NPEND> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(NPEND) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
// Completion counters.  The data they guard is moved by the copy engine, which reads and writes L2 directly: a counter is
// incremented only after the engine's writes have completed (cp.async.bulk.wait_group) and the dependent copies are
// issued only after the counter's value has been seen (control dependency), so relaxed GPU-scope accesses suffice --
// measured: with ld.acquire / red.release the copy thread spent most of its time in the L1 invalidate (CCTL.IVALL) and
// the memory barriers they imply, and the transform warps starved.
__device__ __forceinline__ uint32_t ld_acq(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release(uint32_t *p, uint32_t v) {
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// strip number -> (outer item, strip within it); false for the numbers past the last strip of an outer item
__device__ __forceinline__ bool decode_strip(const F4v2Params &p, uint32_t vs, uint32_t &outer, uint32_t &sw) {
    uint32_t gin;
    fdivmod(vs >> p.glog, p.d_gpo, outer, gin);
    sw = (gin << p.glog) + (vs & ((1u << p.glog) - 1u));
    return sw < p.spo;
}

using PL = P2<7>;  // 128 = 8 x 16: a radix-8 pass, one exchange, a radix-16 pass

template <int P>
__device__ __forceinline__ void compute2(float2 *a, float2 *b, int t, const float2 *__restrict__ stw) {
    constexpr int R = PL::radix(P), NB = 16 / R, ido = PL::ido(P);
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        Dft<float, R>::run(a + j * R);
        Dft<float, R>::run(b + j * R);
    }
    if constexpr (ido > 1) {
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const int i = (t + j * TPL) % ido;
            const float2 *tw = stw + PL::twoff(P) + i;
#pragma unroll
            for (int q = 1; q < R; ++q) {
                const float2 w = __ldg(tw + (q - 1) * ido);
                a[j * R + q] = cmul(a[j * R + q], w);
                b[j * R + q] = cmul(b[j * R + q], w);
            }
        }
    }
}

// the transform warps' work on one tile: stage = 128 rows x 256 B in, the same out
template <bool DBG_COPY>
__device__ __forceinline__ void transform_tile(const F4v2Params &p, unsigned char *stage, uint64_t *ready_bar, uint32_t kind,
                                               uint32_t strip, uint32_t tile, uint32_t flags, int ctid) {
    const int wp = ctid & (WP - 1), t = ctid >> 4;  // pair of lines, butterfly
    float2 a[16], b[16];
    // ---- pass 0 (radix 8): element e = t + 8 j + 16 m of lines 2 wp, 2 wp + 1 ---------------------------------------------
    {
        constexpr int R = 8, NB = 2, ido = 16;
        if (!(flags & FLAG_SHIFT)) {
            const float4 *base = reinterpret_cast<const float4 *>(stage) + t * (ROW / 16) + wp;
#pragma unroll
            for (int j = 0; j < NB; ++j)
#pragma unroll
                for (int m = 0; m < R; ++m) {
                    const float4 u = base[(j * TPL + m * ido) * (ROW / 16)];
                    a[j * R + m] = make_float2(u.x, u.y);
                    b[j * R + m] = make_float2(u.z, u.w);
                }
        } else {
            // 8 bytes into 272-byte rows: 16-byte accesses would be misaligned, and 8-byte accesses 16 bytes apart collide
            // two by two in the banks -- so this tile's thread owns lines wp and wp + 16 (lanes read consecutive elements)
            const unsigned char *base = stage + t * ROW_MIS + 8 + wp * 8;
#pragma unroll
            for (int j = 0; j < NB; ++j)
#pragma unroll
                for (int m = 0; m < R; ++m) {
                    const unsigned char *q = base + (j * TPL + m * ido) * ROW_MIS;
                    a[j * R + m] = *reinterpret_cast<const float2 *>(q);
                    b[j * R + m] = *reinterpret_cast<const float2 *>(q + 128);
                }
        }
        // backward = swap . forward . swap over the WHOLE four-step transform: swap on the way in (A) and out (B) only
        if (p.backward && kind == KIND_A) {
#pragma unroll
            for (int i = 0; i < 16; ++i) { a[i] = cswap(a[i]); b[i] = cswap(b[i]); }
        }
    }
    if (!DBG_COPY) compute2<0>(a, b, t, p.stw);
    // ---- exchange through the stage (the tile is dead once every thread has read its elements) ------------------------------
    {
        float4 *line = reinterpret_cast<float4 *>(stage) + wp * XPITCH;
        constexpr int Rp = 8, NBp = 2;
        bar_compute();
#pragma unroll
        for (int j = 0; j < NBp; ++j)
#pragma unroll
            for (int q = 0; q < Rp; ++q)
                line[p2_phys<7, 1>(t + j * TPL + q * (N / Rp))] = make_float4(a[j * Rp + q].x, a[j * Rp + q].y, b[j * Rp + q].x, b[j * Rp + q].y);
        bar_compute();
        // pass 1: ido = 1, butterfly k = t: elements 16 t + m
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const float4 u = line[p2_phys<7, 1>(m + 16 * t)];
            a[m] = make_float2(u.x, u.y);
            b[m] = make_float2(u.z, u.w);
        }
    }
    if (!DBG_COPY) compute2<1>(a, b, t, p.stw);
    // ---- thread t holds bins k = t + 8 q, q = 0..15, of both lines ------------------------------------------------------------
    if (DBG_COPY) {
    } else if (kind == KIND_A) {
        // times exp(-2 pi i j0 k / 16384): exact two-level look-up for every 4th bin, recurrence in between
        const uint32_t c = tile;
        auto lookup = [&](uint32_t x) {
            uint32_t hi, lo;
            fdivmod(x, p.d_twS, hi, lo);
            return cmul(__ldg(p.twA + hi), __ldg(p.twB + lo));
        };
        const float2 step = lookup(c * 8u);
        float2 wq = make_float2(1.f, 0.f);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            if ((q & 3) == 0) wq = lookup(c * (uint32_t)(t + 8 * q));
            else wq = cmul(wq, step);
            a[q] = cmul(a[q], wq);
            b[q] = cmul(b[q], wq);
        }
    } else {
        const float f = p.fct;
        const bool bw = p.backward != 0;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            a[q] = cscale(a[q], f);
            b[q] = cscale(b[q], f);
            if (bw) { a[q] = cswap(a[q]); b[q] = cswap(b[q]); }
        }
    }
    if (flags & FLAG_STG) {
        // B tiles leave from registers: nothing waits for these stores (no counter to publish), and the stage is free for the
        // next load one staging pass + one copy-engine read-out earlier.  16-byte stores where the rows allow it.
        mbar_arrive(ready_bar);
        uint32_t outer, sw;
        decode_strip(p, strip, outer, sw);
        const uint32_t ext = min((uint32_t)W, p.cols - sw * (uint32_t)W);
        const bool ok0 = 2u * (uint32_t)wp < ext, ok1 = 2u * (uint32_t)wp + 1u < ext;
        if (!ok0) return;
        char *o = p.out + (int64_t)outer * p.out_outer + (int64_t)(sw * (uint32_t)W + 2u * (uint32_t)wp) * 8 +
                  (int64_t)((uint32_t)t * (uint32_t)N + tile) * p.out_pitch;
        const int64_t step_q = (int64_t)(8 * N) * p.out_pitch;
        const uint64_t pol = l2_policy_stream();
        if (ok1 && ((reinterpret_cast<uintptr_t>(o) | (uintptr_t)step_q) & 15) == 0) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(o), "f"(a[q].x), "f"(a[q].y),
                             "f"(b[q].x), "f"(b[q].y), "l"(pol) : "memory");
                o += step_q;
            }
            return;
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            st_policy(reinterpret_cast<float2 *>(o), a[q], pol);
            if (ok1) st_policy(reinterpret_cast<float2 *>(o + 8), b[q], pol);
            o += step_q;
        }
        return;
    }
    if (flags & FLAG_AREG) {
        // A tile to the ring from registers: the stage is free now; the tile is published (doneA) once every thread's stores
        // are ordered before one thread's GPU-scope fence (barrier + fence: the grid-synchronisation pattern)
        mbar_arrive(ready_bar);
        uint32_t rq, slot;
        fdivmod(strip, p.d_ring, rq, slot);
        char *o = p.ring_mem + (int64_t)slot * ((int64_t)N * TILE_BYTES) + (int64_t)((uint32_t)t * (uint32_t)N + tile) * ROW;
        const uint64_t pol = l2_policy_keep();
        if (flags & FLAG_SHIFT) {
            o += wp * 8;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                st_policy(reinterpret_cast<float2 *>(o), a[q], pol);
                st_policy(reinterpret_cast<float2 *>(o + 128), b[q], pol);
                o += 8 * N * ROW;
            }
        } else {
            o += wp * 16;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(o), "f"(a[q].x), "f"(a[q].y),
                             "f"(b[q].x), "f"(b[q].y), "l"(pol) : "memory");
                o += 8 * N * ROW;
            }
        }
        bar_compute();
        if (ctid == 0) {
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            red_release(p.ctr + 2 + strip, 1u);
        }
        return;
    }
    bar_compute();  // every thread has read its exchange elements: the stage becomes the staging area, rows k of 256 B
    if (flags & FLAG_SHIFT) {
        float2 *st = reinterpret_cast<float2 *>(stage) + wp;  // lines wp and wp + 16 (see the load)
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            st[(t + 8 * q) * (ROW / 8)] = a[q];
            st[(t + 8 * q) * (ROW / 8) + 16] = b[q];
        }
    } else {
        float4 *st = reinterpret_cast<float4 *>(stage) + wp;
#pragma unroll
        for (int q = 0; q < 16; ++q) st[(t + 8 * q) * (ROW / 16)] = make_float4(a[q].x, a[q].y, b[q].x, b[q].y);
    }
    fence_proxy_async_smem();  // generic-proxy writes before the copy engine reads them
    mbar_arrive(ready_bar);
}

}  // namespace f4v2

template <int STAGES, bool B_TMA_STORE, bool A_REG, bool DBG_COPY = false>
__global__ void __launch_bounds__(f4v2::NTHREADS, STAGES == 2 ? 3 : 2)
    fft_fourstep_fused2_kernel(const F4v2Params p, const __grid_constant__ F4v2Maps maps) {
    using namespace f4v2;
    extern __shared__ __align__(128) unsigned char smem_f4v2[];
    unsigned char *stages = smem_f4v2;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_f4v2 + STAGES * STAGE_BYTES);
    uint64_t *ready = full + STAGES;
    uint4 *items = reinterpret_cast<uint4 *>(ready + STAGES);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(full + i, 1); mbar_init(ready + i, NCOMP); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp < 4) {
        // ------------------------------------------------ transform warps ------------------------------------------------
        for (uint32_t m = 0;; ++m) {
            const uint32_t b = m % STAGES, par = (m / STAGES) & 1u;
            mbar_wait(full + b, par);
            const uint4 it = items[b];
            if (it.x == KIND_STOP) break;
            transform_tile<DBG_COPY>(p, stages + b * STAGE_BYTES, ready + b, it.x, it.y, it.z, it.w, threadIdx.x);
        }
        return;
    }
    if (threadIdx.x != 4 * 32) return;
    // ---------------------------------------------------- copy thread ------------------------------------------------------
    const uint64_t pol_keep = l2_policy_keep(), pol_stream = l2_policy_stream();
    const uint32_t S = p.nstrips;
    uint32_t *doneA = p.ctr + 2, *doneB = p.ctr + 2 + S;
    uint32_t nload = 0, nret = 0, idle = 0;
    bool exhausted = false, stop_posted = false, aborted = false;
    int pending = -1;       // strip of the most recent A tile whose store to the ring is not yet published
    uint32_t readyA = ~0u;  // last strip seen with doneA complete
    uint32_t freeB = ~0u;   // last strip seen with doneB complete
    uint32_t ticket = atomicAdd(p.ctr, 1u);  // always one ticket ahead: the atomic's latency hides behind the current tile
    bool next_ok = false, nextB = false;     // the ticket is decoded and its dependency satisfied
    uint32_t next_strip = 0, next_tile = 0;
    // The stores of an A tile are complete (wait_group) when this runs; completed copy-engine writes sit in L2, where every
    // later reader (copy engine or ld.acquire) looks, and the release orders the counter behind them.
    auto publish_pending = [&]() {
        red_release(doneA + pending, 1u);
        pending = -1;
    };
    for (;;) {
        bool did = false;
        // ---- resolve the next tile's dependency ahead of time (needs no stage) ---------------------------------------------
        if (!exhausted && !next_ok) {
            if (ticket >= p.total_items) exhausted = true;
            else {
                // ticket -> (unit, tile-in-unit) -> (step, group) x (j0 or k1, strip in group)
                const uint32_t within = ticket & ((128u << p.glog) - 1u);
                uint32_t grp;
                fuse4_decode_unit(ticket >> (7 + p.glog), p.ngroups, p.lag, nextB, grp);
                if (p.dbg_only) { nextB = p.dbg_only == 2; grp = ticket >> (7 + p.glog); }
                next_strip = (grp << p.glog) + (within & ((1u << p.glog) - 1u));
                next_tile = within >> p.glog;
                next_ok = true;
                uint32_t o_, s_;
                if (!decode_strip(p, next_strip, o_, s_)) {  // past the last strip: nothing to do for this ticket
                    ticket = atomicAdd(p.ctr, 1u);
                    next_ok = false;
                    continue;
                }
                if (nextB && next_strip != readyA && !p.dbg_only) {
                    next_ok = ld_acq(doneA + next_strip) >= (uint32_t)N;
                    if (next_ok) readyA = next_strip;
                }
                const uint32_t rs = p.ring << p.glog;
                if (A_REG && next_ok && !nextB && next_strip >= rs && freeB != next_strip - rs && !p.dbg_only &&
                    decode_strip(p, next_strip - rs, o_, s_)) {
                    // the transform warps store this tile to the ring themselves: the slot must be free before the tile starts
                    next_ok = ld_acq(doneB + (next_strip - rs)) >= (uint32_t)N;
                    if (next_ok) freeB = next_strip - rs;
                }
            }
        }
        // ---- retire the oldest tile whose stage the transform warps have left: store it -----------------------------------
        if (nret < nload) {
            const uint32_t b = nret % STAGES, par = (nret / STAGES) & 1u;
            if (mbar_test(ready + b, par)) {
                const uint4 it = items[b];
                const uint32_t strip = it.y, tile = it.z;
                const unsigned char *src = stages + b * STAGE_BYTES;
                if (it.x == KIND_B) {
                    red_release(doneB + strip, 1u);  // the slot's rows of this tile were consumed when the tile landed
                    if (!(it.w & FLAG_STG)) {
                        uint32_t outer, sw;
                        decode_strip(p, strip, outer, sw);
                        tma_store_4d(&maps.m[2 + (tile & 1u)], (int)(sw * W), (int)(tile >> 1), 0, (int)outer, src, pol_stream);
                        bulk_commit();
                        if (pending >= 0) { bulk_wait<1>(); publish_pending(); }
                        bulk_wait_read<0>();  // the copy engine has read the stage: it can be refilled
                    }
                    ++nret;
                    did = true;
                } else if (A_REG) {
                    ++nret;  // stored and published by the transform warps
                    did = true;
                } else {
                    bool ok = true;
                    const uint32_t rs = p.ring << p.glog;  // ring slots
                    uint32_t o_, s_;
                    if (strip >= rs && !aborted && freeB != strip - rs && !p.dbg_only && decode_strip(p, strip - rs, o_, s_)) {
                        ok = ld_acq(doneB + (strip - rs)) >= (uint32_t)N;
                        if (ok) freeB = strip - rs;
                    }
                    if (ok) {
                        uint32_t rq, slot;
                        fdivmod(strip, p.d_ring, rq, slot);
                        tma_store_4d(&maps.m[4], 0, (int)tile, 0, (int)slot, src, pol_keep);
                        bulk_commit();
                        if (pending >= 0) { bulk_wait<1>(); publish_pending(); }
                        bulk_wait_read<0>();
                        pending = (int)strip;
                        ++nret;
                        did = true;
                    }
                }
            }
        }
        // ---- load the next tile into a free stage -------------------------------------------------------------------------
        if (!stop_posted && nload - nret < (uint32_t)STAGES) {
            const uint32_t b = nload % STAGES;
            if (exhausted) {
                items[b] = make_uint4(KIND_STOP, 0, 0, 0);
                mbar_arrive(full + b);
                stop_posted = true;
                did = true;
            } else if (next_ok) {
                const uint32_t strip = next_strip, tile = next_tile, odd = tile & 1u;
                unsigned char *dst = stages + b * STAGE_BYTES;
                uint32_t outer, sw;
                decode_strip(p, strip, outer, sw);
                uint32_t flags;
                // B_TMA_STORE: B tiles of 16-byte aligned rows and full strips staged and stored by the copy engine (a store
                // box must not reach past the strip); default: every B tile is stored from registers
                if (nextB) flags = (!B_TMA_STORE || p.mis_out[odd] || (sw + 1u) * (uint32_t)W > p.cols) ? FLAG_STG : 0u;
                else flags = (p.mis_in[odd] ? FLAG_SHIFT : 0u) | (A_REG ? FLAG_AREG : 0u);
                items[b] = make_uint4(nextB ? KIND_B : KIND_A, strip, tile, flags);
                mbar_arrive_tx(full + b, (flags & FLAG_SHIFT) ? (uint32_t)(N * ROW_MIS) : (uint32_t)TILE_BYTES);
                if (nextB) {
                    uint32_t rq, slot;
                    fdivmod(strip, p.d_ring, rq, slot);
                    const char *src = p.ring_mem + (int64_t)slot * ((int64_t)N * TILE_BYTES) + (int64_t)tile * TILE_BYTES;
                    bulk_g2s(dst, src, TILE_BYTES, full + b, pol_keep);
                } else {
                    // (a misaligned class: the map starts 8 bytes early and its box is 34 wide, so x = sw * W is aligned)
                    tma_load_4d(dst, &maps.m[odd], (int)(sw * W), (int)(tile >> 1), 0, (int)outer, full + b, pol_stream);
                }
                ++nload;
                ticket = atomicAdd(p.ctr, 1u);
                next_ok = false;
                did = true;
            }
        }
        if (stop_posted && nret == nload) break;
        if (did) { idle = 0; continue; }
        if (pending >= 0) { bulk_wait<0>(); publish_pending(); continue; }
        __nanosleep(32);
        if (++idle > p.max_idle && !aborted) {
            // a dependency never completed (a bug, or a foreign fault): flag it and run the queue dry without waiting
            aborted = true;
            atomicExch(p.ctr + 1, 1u);
            if (p.host_err) *reinterpret_cast<volatile uint32_t *>(p.host_err) = 1u;
            exhausted = true;
            next_ok = false;
        }
    }
    if (pending >= 0) { bulk_wait<0>(); publish_pending(); }
    bulk_wait<0>();  // every store of this CTA is complete before it exits
}

}  // namespace rfb
