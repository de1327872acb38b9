// Write-bandwidth ceiling of the decoder's output pattern: every warp owns a 32 KB region (32 rows of 1 KB) and
// writes it in passes of ROWB bytes per row (decoder: ROWB = 128), with an optional compute-like delay between passes.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int ROWB>
__global__ void k_write(uint8_t* out, long long nregions, int delay) {
    const int lane = threadIdx.x & 31;
    const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = (long long)gridDim.x * (blockDim.x >> 5);
    constexpr int LPR = ROWB / 16;          // lanes per row
    constexpr int RPI = 32 / LPR;           // rows per instruction
    for (long long r = w; r < nregions; r += nw) {
        uint8_t* base = out + r * 32768;
        for (int pass = 0; pass < 1024 / ROWB; ++pass) {
#pragma unroll
            for (int i = 0; i < 32 / RPI; ++i) {
                const int row = i * RPI + lane / LPR, col = lane % LPR;
                *reinterpret_cast<uint4*>(base + row * 1024 + pass * ROWB + col * 16) = make_uint4(pass, i, lane, (uint32_t)r);
            }
            if (delay) {
                const long long t0 = clock64();
                while (clock64() - t0 < delay) {}
            }
        }
    }
}

template <int ROWB>
void run(uint8_t* d, long long bytes, int warps, int delay) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const long long nreg = bytes / 32768;
    k_write<ROWB><<<148, warps * 32>>>(d, nreg, delay);
    cudaEventRecord(e0);
    for (int i = 0; i < 3; ++i) k_write<ROWB><<<148, warps * 32>>>(d, nreg, delay);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
    printf("{\"rowb\": %d, \"warps_per_sm\": %d, \"delay\": %d, \"ms\": %.4f, \"GBs\": %.1f}\n", ROWB, warps, delay, ms, bytes / ms / 1e6);
}

int main() {
    const long long bytes = 4500000000ll / 32768 * 32768;
    uint8_t* d; cudaMalloc(&d, bytes);
    for (int delay : {0, 2000, 4000}) {
        for (int warps : {8, 14, 20, 32}) {
            run<128>(d, bytes, warps, delay);
            run<256>(d, bytes, warps, delay);
            run<512>(d, bytes, warps, delay);
            run<1024>(d, bytes, warps, delay);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
