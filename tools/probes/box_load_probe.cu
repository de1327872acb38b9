// How fast can the TMA engine stream a [C][T] uint8 recording as boxes of BW bins x 32 channels (the staging pattern of
// k_encode_rows: every box row is a separate 128/256-byte piece of a different channel row)?  One persistent CTA per SM, W warps,
// every warp walks 32-channel blocks with NST boxes in flight and does nothing with the data.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o box_load_probe box_load_probe.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BW, int NST>
__global__ void __launch_bounds__(1024, 1) k_box(const __grid_constant__ CUtensorMap tmap, int C, int T, int W, unsigned long long* sink) {
    extern __shared__ __align__(1024) uint8_t sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int STAGE = BW * 32;
    uint8_t* st = sm + 1024 + (size_t)warp * NST * STAGE;
    const uint32_t bar0 = smem_u32(sm) + warp * NST * 8;
    if (lane == 0) {
        for (int i = 0; i < NST; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * i));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp >= W) return;
    const int nblk = (C + 31) / 32, nbox = (T + BW - 1) / BW;
    uint32_t phase = 0, acc = 0;
    for (int blk = blockIdx.x + gridDim.x * warp; blk < nblk; blk += gridDim.x * W) {
        auto issue = [&](int b) {
            const uint32_t s = b % NST;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * s), "r"(STAGE) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                             smem_u32(st + s * STAGE)),
                         "l"(&tmap), "r"(b * BW), "r"(blk * 32), "r"(bar0 + 8 * s)
                         : "memory");
        };
        if (lane == 0)
            for (int b = 0; b < NST && b < nbox; ++b) issue(b);
        for (int b = 0; b < nbox; ++b) {
            const uint32_t s = b % NST;
            uint32_t ok;
            do {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(bar0 + 8 * s), "r"((phase >> s) & 1u) : "memory");
            } while (!ok);
            phase ^= 1u << s;
            acc += *reinterpret_cast<volatile uint32_t*>(st + s * STAGE + lane * 4);
            __syncwarp();
            if (lane == 0 && b + NST < nbox) issue(b + NST);
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BW, int NST>
void run(EncFn enc, uint8_t* d, int C, int T, int W, CUtensorMapL2promotion prom, const char* pn) {
    const long long stride = (T + 15) / 16 * 16;
    CUtensorMap m;
    const cuuint64_t gdim[2] = {(cuuint64_t)T, (cuuint64_t)C}, gstr[1] = {(cuuint64_t)stride};
    const cuuint32_t box[2] = {BW, 32}, estr[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     BW == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return; }
    const int smem = 1024 + W * NST * BW * 32;
    if (smem > 227 * 1024) return;
    cudaFuncSetAttribute(k_box<BW, NST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    unsigned long long* sink; cudaMalloc(&sink, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_box<BW, NST><<<148, 1024, smem>>>(m, C, T, W, sink);
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) k_box<BW, NST><<<148, 1024, smem>>>(m, C, T, W, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    printf("{\"box\": %d, \"stages\": %d, \"warps\": %d, \"C\": %d, \"T\": %d, \"l2prom\": \"%s\", \"ms\": %.4f, \"GBs\": %.0f, \"err\": \"%s\"}\n", BW, NST, W, C, T, pn, ms,
           (double)C * T / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    cudaFree(sink);
}

int main() {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncFn enc = (EncFn)p;
    uint8_t* d; cudaMalloc(&d, 1300000000ll); cudaMemset(d, 1, 1300000000ll);
    const int Ts[3] = {2400, 12000, 120000};
    for (int ti = 0; ti < 3; ++ti) {
        const int T = Ts[ti], C = ti == 2 ? 10000 : 100000;
        run<128, 2>(enc, d, C, T, 22, CU_TENSOR_MAP_L2_PROMOTION_NONE, "none");
        run<128, 2>(enc, d, C, T, 22, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "256");
        run<128, 3>(enc, d, C, T, 16, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "256");
        run<128, 4>(enc, d, C, T, 13, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "256");
        run<256, 2>(enc, d, C, T, 13, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "256");
        run<256, 1>(enc, d, C, T, 22, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "256");
        run<64, 4>(enc, d, C, T, 22, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "256");
    }
    return 0;
}
