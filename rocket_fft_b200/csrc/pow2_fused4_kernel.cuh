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

template <typename T, int LOGN, int W>
__global__ void __launch_bounds__(W *(1 << LOGN) / 16, p2_min_blocks<T, W *(1 << LOGN) / 16>())
    fft_fourstep_fused_kernel(const TileGeom<T> gA, const TileGeom<T> gB, const cx<T> *__restrict__ stw, const Fuse4Ctl c) {
    using Body = Pow2Body<T, LOGN, W, 0>;
    extern __shared__ __align__(16) unsigned char smem_raw_f4[];
    cx<T> *buf = reinterpret_cast<cx<T> *>(smem_raw_f4);
    __shared__ uint32_t s_item;
    const uint32_t S = c.nstrips, lag = c.lag;
    const uint32_t total_units = 2 * S;
    for (;;) {
        __syncthreads();  // the previous tile is out of shared memory; s_item may be overwritten
        if (threadIdx.x == 0) s_item = atomicAdd(&c.ctr[0], 1u);
        __syncthreads();
        uint32_t unit, tile;
        fdivmod(s_item, c.d_tiles, unit, tile);
        if (unit >= total_units) return;
        // unit -> (step, strip): A(0..lag-1), then pairs A(lag + i), B(i), then the last B's
        uint32_t strip;
        bool stepB;
        if (unit < lag) { stepB = false; strip = unit; }
        else if (unit < lag + 2 * (S - lag)) { const uint32_t v = unit - lag; stepB = (v & 1u) != 0; strip = stepB ? (v >> 1) : lag + (v >> 1); }
        else { stepB = true; strip = S - lag + (unit - lag - 2 * (S - lag)); }
        // dependencies
        if (threadIdx.x == 0) {
            const uint32_t *flag = nullptr;
            if (stepB) flag = c.ctr + 2 + strip;                              // all of A(strip) written
            else if (strip >= c.ring) flag = c.ctr + 2 + S + (strip - c.ring);  // B(strip - ring) has read the slot
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
        uint32_t outer, sw, trest, t0;
        fdivmod(strip, c.d_spo, outer, sw);
        fdivmod(tile, gA.d_t0, trest, t0);
        const uint32_t ext0 = min(c.cw, c.cols - sw * c.cw);  // the last strip of an outer item may be narrower
        const int64_t slot = (int64_t)(strip % c.ring) * c.slot_bytes;
        if (t0 * (uint32_t)W < ext0) {
            if (!stepB) Body::template run_tile<1>(gA, stw, buf, tile, (int64_t)outer * c.in_outer + (int64_t)sw * c.in_strip, slot, ext0);
            else Body::template run_tile<2>(gB, stw, buf, tile, slot, (int64_t)outer * c.out_outer + (int64_t)sw * c.out_strip, ext0);
        }
        __threadfence();  // this thread's stores are visible GPU-wide before the tile is counted as done
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(&c.ctr[2 + (stepB ? S : 0) + strip], 1u);
    }
}

}  // namespace rfb
