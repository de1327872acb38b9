// Write-bandwidth of the decoder's output pattern when every LANE stores its own row with a TMA bulk copy
// (cp.async.bulk.global.shared::cta, SASS UBLKCP) instead of the cooperative LDS.128 + ST.128 write-out:
// every warp owns a 32 KB region (32 rows of 1 KB) and writes it in passes of ROWB bytes per row from a
// double-buffered shared-memory tile, with an optional compute-like delay between passes.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int ROWB>
__global__ void k_bulk(uint8_t* out, long long nregions, int delay) {
    extern __shared__ __align__(128) uint8_t sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int RS = ROWB + 16;                         // padded row
    uint8_t* tile = sm + (size_t)warp * 2 * 32 * RS;      // two buffers
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + warp, nw = (long long)gridDim.x * (blockDim.x >> 5);
    int buf = 0;
    for (long long r = w; r < nregions; r += nw) {
        uint8_t* base = out + r * 32768 + lane * 1024;
        for (int pass = 0; pass < 1024 / ROWB; ++pass) {
            uint8_t* row = tile + (size_t)buf * 32 * RS + lane * RS;
            // the buffer used two passes ago must have been read by the TMA engine
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
#pragma unroll
            for (int k = 0; k < ROWB / 16; ++k) reinterpret_cast<uint4*>(row)[k] = make_uint4(pass, k, lane, (uint32_t)r);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + pass * ROWB), "r"(smem_u32(row)), "r"(ROWB) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            buf ^= 1;
            if (delay) {
                const long long t0 = clock64();
                while (clock64() - t0 < delay) {}
            }
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int ROWB>
void run(uint8_t* d, long long bytes, int warps, int delay) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const long long nreg = bytes / 32768;
    const int smem = warps * 2 * 32 * (ROWB + 16);
    cudaFuncSetAttribute(k_bulk<ROWB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_bulk<ROWB><<<148, warps * 32, smem>>>(d, nreg, delay);
    cudaEventRecord(e0);
    for (int i = 0; i < 3; ++i) k_bulk<ROWB><<<148, warps * 32, smem>>>(d, nreg, delay);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
    cudaError_t e = cudaGetLastError();
    printf("{\"bulk_rowb\": %d, \"warps_per_sm\": %d, \"delay\": %d, \"ms\": %.4f, \"GBs\": %.1f, \"err\": \"%s\"}\n", ROWB, warps, delay, ms,
           bytes / ms / 1e6, cudaGetErrorString(e));
}

int main() {
    const long long bytes = 4500000000ll / 32768 * 32768;
    uint8_t* d; cudaMalloc(&d, bytes);
    for (int delay : {0, 2000, 4000}) {
        for (int warps : {8, 14, 20}) {
            run<128>(d, bytes, warps, delay);
            run<256>(d, bytes, warps, delay);
        }
    }
    // spot check: last region written as expected
    uint32_t h[4];
    cudaMemcpy(h, d + (bytes / 32768 - 1) * 32768 + 5 * 1024 + 3 * 256 + 16, 16, cudaMemcpyDeviceToHost);
    printf("{\"check\": [%u, %u, %u, %u]}\n", h[0], h[1], h[2], h[3]);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
