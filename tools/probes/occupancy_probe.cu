// Resident CTAs per SM of the encoders / decoders with the shared memory the ABI launches them with.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/occ tools/probes/occupancy_probe.cu && gpurun_out/occ
#include <cstdio>
#include "../../hardware-efficient-mua-compression_b200/csrc/mua_common.cuh"
#include "../../hardware-efficient-mua-compression_b200/csrc/mua_encode.cuh"
#include "../../hardware-efficient-mua-compression_b200/csrc/mua_decode.cuh"
using namespace mua;
template <typename F>
static void show(const char* name, F f, int threads, int smem, bool maxshared) {
    cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (maxshared) cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, f, threads, smem);
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, f);
    printf("%-22s threads %4d smem %6d regs %3d carveout-pref %d -> %d CTAs/SM (%s)\n", name, threads, smem, a.numRegs,
           a.preferredShmemCarveout, n, cudaGetErrorString(e));
}
int main() {
    for (int ms = 0; ms < 2; ++ms) {
        show("k_encode_fast<3>", k_encode_fast<3>, EF_WARPS * 32, EncFastSmem::PER_WARP * EF_WARPS, ms);
        show("k_encode_pair<5>", k_encode_pair<5>, ENC_WARPS * 32, EncPairSmem::PER_WARP * ENC_WARPS, ms);
        show("k_encode_gen", k_encode_gen, ENC_WARPS * 32, EncGenSmem::PER_WARP * ENC_WARPS, ms);
    }
    return 0;
}
