// Standalone bisect aid for the tensor-map TMA path of fused4v2_kernel.cuh (not part of the product).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

struct Maps { CUtensorMap m[2]; };

__device__ __forceinline__ uint32_t su32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int VARIANT>
__global__ void __launch_bounds__(160) probe(const __grid_constant__ Maps maps, const __grid_constant__ CUtensorMap one, int x0, int y0, int which,
                                            unsigned long long *out_dbg) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + 32768);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    const bool leader = (VARIANT & 1) ? (threadIdx.x == 128) : (threadIdx.x == 0);
    if ((VARIANT & 2) && (threadIdx.x >> 5) == 4 && !leader) return;  // the other lanes of the copy warp exit
    if (leader) {
        const CUtensorMap *mp = (VARIANT & 4) ? &one : &maps.m[which];
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(bar)), "r"(32768u) : "memory");
        if (VARIANT & 8)
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                         ::"r"(su32(sm)), "l"(mp), "r"(x0), "r"(y0), "r"(0), "r"(0), "r"(su32(bar)) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;"
                         ::"r"(su32(sm)), "l"(mp), "r"(x0), "r"(y0), "r"(0), "r"(0), "r"(su32(bar)), "l"(pol) : "memory");
    }
    if ((threadIdx.x >> 5) < 4) {
        uint32_t ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(su32(bar)), "r"(0u) : "memory");
        } while (!ok);
        // checksum of the tile
        const unsigned long long *v = reinterpret_cast<const unsigned long long *>(sm);
        unsigned long long acc = 0;
        for (int i = threadIdx.x; i < 4096; i += 128) acc += v[i] * (unsigned long long)(i + 1);
        atomicAdd(out_dbg, acc);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaFree(0);
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr);
    printf("entry point: err %d qr %d fn %p\n", (int)e, (int)qr, f);
    EncodeTiledFn fn = (EncodeTiledFn)f;
    const uint64_t rows = 16384, cols = 777, pitch = cols * 8;
    unsigned long long *d = nullptr, *dbg = nullptr;
    cudaMalloc(&d, rows * pitch);
    cudaMalloc(&dbg, 8);
    std::vector<unsigned long long> h(rows * cols);
    for (size_t i = 0; i < h.size(); ++i) h[i] = i * 2654435761ull + 12345;
    cudaMemcpy(d, h.data(), rows * pitch, cudaMemcpyHostToDevice);
    alignas(64) Maps maps;
    for (int c = 0; c < 2; ++c) {
        const char *b0 = (const char *)d + c * pitch;
        const uint32_t mis = (uint32_t)((uintptr_t)b0 & 15u);
        const cuuint64_t dim[4] = {cols + mis / 8, 64, 128, 1};
        const cuuint64_t str[3] = {2 * pitch, 128 * pitch, 16384 * pitch};
        const cuuint32_t box[4] = {32, 1, 128, 1};
        const cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = fn(&maps.m[c], CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, (void *)(b0 - mis), dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode map %d: CUresult %d (mis %u)\n", c, (int)r, mis);
    }
    auto expect = [&](int which, int x0, int j0h) {
        unsigned long long acc = 0;
        const int j0 = 2 * j0h + which;
        for (int j1 = 0; j1 < 128; ++j1)
            for (int x = 0; x < 32; ++x) {
                const int col = x0 + x;  // column in the array
                unsigned long long v = (col >= 0 && col < (int)cols) ? h[(size_t)(j1 * 128 + j0) * cols + col] : 0ull;
                acc += v * (unsigned long long)(j1 * 32 + x + 1);
            }
        return acc;
    };
#define RUN(V, which, col0, j0h)                                                                                         \
    do {                                                                                                                 \
        cudaMemset(dbg, 0, 8);                                                                                           \
        cudaFuncSetAttribute(probe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 + 64);                         \
        const int xs = (int)(((uintptr_t)((const char *)d + which * pitch) & 15u) / 8);                                   \
        probe<V><<<1, 160, 32768 + 64>>>(maps, maps.m[which], col0 + xs, j0h, which, dbg);                               \
        cudaError_t er = cudaDeviceSynchronize();                                                                        \
        unsigned long long got = 0;                                                                                      \
        cudaMemcpy(&got, dbg, 8, cudaMemcpyDeviceToHost);                                                                \
        printf("variant %2d map %d col0 %4d j0h %2d: %s  checksum %s\n", V, which, col0, j0h, cudaGetErrorString(er),   \
               got == expect(which, col0, j0h) ? "OK" : "MISMATCH");                                                     \
        if (er != cudaSuccess) return 1;                                                                                 \
    } while (0)
    RUN(12, 0, 0, 0);   // static map param, no cache hint, leader = thread 0
    RUN(4, 0, 0, 0);    // + cache hint
    RUN(0, 0, 0, 0);    // dynamic index into the struct
    RUN(0, 1, 64, 3);   // odd rows (map base 8 bytes down, x one up)
    RUN(0, 1, 768, 5);  // partial strip (clipped at the edge)
    RUN(1, 1, 64, 3);   // leader = lane 0 of warp 4
    RUN(3, 1, 64, 3);   // other lanes of warp 4 exited
    printf("all variants ran\n");
    return 0;
}
