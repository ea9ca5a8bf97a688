// Four-step transform of long strided lines (n = n1*n1, e.g. the 16384-point columns of a 16384 x 8193 half spectrum)
// as ONE persistent kernel whose intermediate never leaves the L2 cache.
//
// The two-launch four-step (engine.cu::run_fourstep) writes the whole intermediate array to HBM after step A and reads
// it back in step B: 4 x the array in DRAM traffic.  Here the neighbouring-lines dim (the columns) is cut into strips of
// CW lines; a strip's intermediate (n x CW points, 8-16 MiB) lives in one slot of a small ring of scratch buffers that
// stays resident in the 126 MB L2 (evict_last policy on its loads/stores, evict_first on the array's), and the
// CTAs of ONE launch work through an ordered list of tiles
//     A(0) A(1) .. A(lag-1) | A(lag) B(0) | A(lag+1) B(1) | ...          (A(s), B(s): the tiles of step A / B of strip s)
// handed out by an atomic ticket.  B(s) waits until all tiles of A(s) are done, A(s) until B(s - ring) has released its
// slot; producers always precede their consumers in ticket order, so a waiting CTA never holds up the CTA it waits for
// (no co-residency assumption, no deadlock) and with lag >= 2 the wait is normally over before it starts.  DRAM sees the
// array once in and once out; there are no wave tails between the steps.
#pragma once
#include "pow2_kernel.cuh"

namespace rfb {

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// MINB: resident CTAs per SM the kernel is compiled for (0: the register kernel's own target).  More CTAs mean fewer
// registers per thread and a few spilled values, but the tile body is latency-bound and the
// fused form has DRAM headroom to use.
template <typename T, int LOGN, int W, int MINB = 0>
__global__ void __launch_bounds__(W *(1 << LOGN) / 16, MINB ? MINB : p2_min_blocks<T, W *(1 << LOGN) / 16>())
    fft_fourstep_fused_kernel(const TileGeom<T> gA, const TileGeom<T> gB, const cx<T> *__restrict__ stw, const Fuse4Ctl c) {
    using Body = Pow2Body<T, LOGN, W, 0>;
    extern __shared__ __align__(16) unsigned char smem_raw_f4[];
    cx<T> *buf = reinterpret_cast<cx<T> *>(smem_raw_f4);
    // Tickets are drawn two tiles ahead and the next tile's dependency is probed while the current tile runs, so that in
    // the steady state a CTA goes from one tile to the next through a single barrier.  (A CTA may hold up to three
    // tickets; it works them off in order, and the smallest unfinished ticket never depends on a larger one.)
    __shared__ uint32_t s_item[3], s_ready[3];
    const uint32_t S = c.nstrips, lag = c.lag;
    // item -> (step, strip, tile); unit order: A(0..lag-1), then pairs A(lag + i), B(i), then the last B's
    auto decode = [&](uint32_t item, bool &stepB, uint32_t &strip, uint32_t &tile) -> bool {
        uint32_t unit;
        fdivmod(item, c.d_tiles, unit, tile);
        return fuse4_decode_unit(unit, S, lag, stepB, strip);
    };
    // the counter a tile waits for (nullptr: none): B(s) needs all of A(s) written, A(s) needs B(s - ring) to have read the slot
    auto dep_flag = [&](bool stepB, uint32_t strip) -> const uint32_t * {
        if (stepB) return c.ctr + 2 + strip;
        if (strip >= c.ring) return c.ctr + 2 + S + (strip - c.ring);
        return nullptr;
    };
    if (threadIdx.x == 0) {
        s_item[0] = atomicAdd(&c.ctr[0], 1u);
        s_item[1] = atomicAdd(&c.ctr[0], 1u);
        s_ready[0] = 0;
        s_ready[1] = 0;
    }
    __syncthreads();
    for (uint32_t slot = 0;; slot = (slot == 2 ? 0 : slot + 1)) {
        const uint32_t slot1 = (slot == 2 ? 0 : slot + 1), slot2 = (slot1 == 2 ? 0 : slot1 + 1);
        bool stepB;
        uint32_t strip, tile;
        if (!decode(s_item[slot], stepB, strip, tile)) return;
        const bool ready = s_ready[slot] != 0;
        uint32_t t2 = 0, nflag = 0;
        if (threadIdx.x == 0) {
            t2 = atomicAdd(&c.ctr[0], 1u);  // ticket of the tile after the next; consumed at the end of this tile
            bool nB;
            uint32_t nstrip, ntile;
            nflag = c.tiles;
            if (decode(s_item[slot1], nB, nstrip, ntile)) {
                const uint32_t *nf = dep_flag(nB, nstrip);
                if (nf) nflag = ld_acquire_u32(nf);  // probe of the next tile's dependency, also consumed at the end
            }
        }
        if (!ready) {
            if (threadIdx.x == 0) {
                const uint32_t *flag = dep_flag(stepB, strip);
                if (flag) {
                    // never hang the GPU: after ~1 s without progress flag an error (ctr[1]) and stop waiting
                    uint32_t spins = 0;
                    while (ld_acquire_u32(flag) < c.tiles) {
                        __nanosleep(200);
                        if (++spins > (1u << 22) || ld_acquire_u32(c.ctr + 1) != 0) { atomicExch(&c.ctr[1], 1u); break; }
                    }
                }
            }
            __syncthreads();
        }
        uint32_t outer, sw, trest, t0, ringq, slot_i;
        fdivmod(strip, c.d_spo, outer, sw);
        fdivmod(tile, gA.d_t0, trest, t0);
        fdivmod(strip, c.d_ring, ringq, slot_i);
        const uint32_t ext0 = min(c.cw, c.cols - sw * c.cw);  // the last strip of an outer item may be narrower
        const int64_t sbytes = (int64_t)slot_i * c.slot_bytes;
        if (t0 * (uint32_t)W < ext0) {
            if (!stepB) Body::template run_tile<1>(gA, stw, buf, tile, (int64_t)outer * c.in_outer + (int64_t)sw * c.in_strip, sbytes, ext0);
            else Body::template run_tile<2>(gB, stw, buf, tile, sbytes, (int64_t)outer * c.out_outer + (int64_t)sw * c.out_strip, ext0);
        }
        if (threadIdx.x == 0) {
            s_item[slot2] = t2;
            s_ready[slot1] = nflag >= c.tiles ? 1u : 0u;
        }
        __syncthreads();  // every thread's stores are issued; shared memory and the control slots are free again
        if (threadIdx.x == 0) {
            // release: the CTA's stores (ordered before this thread by the barrier) become visible GPU-wide first
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            atomicAdd(&c.ctr[2 + (stepB ? S : 0) + strip], 1u);
        }
    }
}

}  // namespace rfb
