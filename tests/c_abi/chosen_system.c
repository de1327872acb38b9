/* Plain-C client of libmua_b200.so (include/mua_b200.h): the chosen system of the reference
 * (test_chosen_system.py:22-27: S = 3, BP = 50 ms, 2^6-sample histogram, codebook 0/10/11) on a small
 * host-generated recording -- calibrate, encode, decode, verify -- with nothing but the C ABI and the CUDA
 * runtime's C API for memory.  It re-derives every channel's bit count on the host the way the reference
 * counts it (histogram of the saturated, rank-mapped post window times the codeword lengths,
 * get_BR_no_sort.py:171-287 / functions_1.py:75-90) and compares.
 *
 *   gcc -std=c99 -Iinclude -I/usr/local/cuda/include tests/c_abi/chosen_system.c \
 *       -L<pkg> -lmua_b200 -L/usr/local/cuda/lib64 -lcudart -o chosen_system
 */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mua_b200.h"

#define CK(x) do { int rc_ = (x); if (rc_ != MUA_OK) { fprintf(stderr, "%s -> %d: %s\n", #x, rc_, mua_last_error()); return 2; } } while (0)
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 3; } } while (0)

static uint32_t lcg(uint32_t* s) { *s = *s * 1664525u + 1013904223u; return *s >> 8; }

int main(void) {
    enum { C = 300, T = 5000, S = 3, H = 64, K = 1 };
    const int64_t stride = (T + 15) / 16 * 16;
    const int32_t chunk_stride = (T + MUA_CHUNK - 1) / MUA_CHUNK;
    const int64_t slot = ((int64_t)(T / 2 + 16) * 2 + 127) / 128 * 16 + 16;      /* worst case: 2 bits per symbol */
    const uint8_t lens[S] = {1, 2, 2};
    uint16_t codes[S];
    CK(mua_canonical_codebook(lens, K, S, codes));
    if (codes[0] != 0 || codes[1] != 2 || codes[2] != 3) { fprintf(stderr, "codebook is not 0/10/11\n"); return 1; }

    /* synthetic counts: a per-channel mix of 0/1/2/3+ (values above S-1 exercise the saturation) */
    uint8_t* h_sym = (uint8_t*)calloc((size_t)C * stride, 1);
    uint32_t seed = 12345u;
    for (int c = 0; c < C; ++c) {
        const uint32_t p1 = 20 + (uint32_t)(c * 7) % 200, p2 = 5 + (uint32_t)(c * 3) % 60;    /* out of 256 */
        for (int t = 0; t < T; ++t) {
            const uint32_t u = lcg(&seed) & 255u;
            h_sym[c * stride + t] = (uint8_t)(u < p2 ? 2 + (lcg(&seed) & 3u) : (u < p2 + p1 ? 1 : 0));
        }
    }

    uint8_t *d_sym, *d_dec, *d_stream, *d_peak, *d_enc;
    void* d_tab;
    int32_t *d_cut, *d_end, *d_ovf;
    uint32_t* d_co;
    uint32_t* d_so;
    int64_t* d_bits;
    unsigned long long* d_mis;
    CU(cudaMalloc((void**)&d_sym, (size_t)C * stride));
    CU(cudaMalloc((void**)&d_dec, (size_t)C * stride));
    CU(cudaMalloc((void**)&d_stream, (size_t)C * slot));
    CU(cudaMalloc((void**)&d_peak, C));
    CU(cudaMalloc((void**)&d_enc, C));
    CU(cudaMalloc(&d_tab, mua_tables_bytes(S, K)));
    CU(cudaMalloc((void**)&d_cut, C * sizeof(int32_t)));
    CU(cudaMalloc((void**)&d_end, C * sizeof(int32_t)));
    CU(cudaMalloc((void**)&d_ovf, 2 * sizeof(int32_t)));   /* [0] encode overflow flag, [1] decode status */
    CU(cudaMalloc((void**)&d_co, (size_t)C * chunk_stride * sizeof(uint32_t)));
    CU(cudaMalloc((void**)&d_so, (size_t)C * 8 * chunk_stride * sizeof(uint32_t)));
    CU(cudaMalloc((void**)&d_bits, C * sizeof(int64_t)));
    CU(cudaMalloc((void**)&d_mis, sizeof(unsigned long long)));
    CU(cudaMemcpy(d_sym, h_sym, (size_t)C * stride, cudaMemcpyHostToDevice));
    CU(cudaMemset(d_dec, 0xEE, (size_t)C * stride));
    CU(cudaMemset(d_ovf, 0, 2 * sizeof(int32_t)));

    const int32_t hH[1] = {H};
    CK(mua_build_tables(d_tab, lens, codes, S, K, NULL));
    CK(mua_calibrate(d_sym, NULL, NULL, stride, T, C, S, hH, 1, 1, MUA_WINDOW_TRUNCATE, d_tab, 1u, 0u,
                     d_cut, d_end, d_peak, d_enc, NULL, NULL, NULL, NULL, NULL));
    CK(mua_encode(d_sym, NULL, NULL, stride, T, C, S, d_cut, d_end, d_peak, d_enc, d_tab, K, 2, d_stream, slot, d_co,
                  chunk_stride, d_so, 8 * chunk_stride, d_bits, d_ovf, NULL, NULL));
    CK(mua_decode(d_stream, slot, d_co, chunk_stride, d_so, 8 * chunk_stride, NULL, stride, C, S, d_cut, d_end, d_peak, d_enc,
                  d_tab, K, 2, H + T / 2, d_dec, d_ovf + 1, NULL, 0, NULL));
    CK(mua_verify(d_sym, d_dec, NULL, stride, C, S, d_cut, d_end, d_mis, NULL));
    CU(cudaDeviceSynchronize());

    int32_t h_cut[C], h_end[C], ovf[2];
    uint8_t h_peak[C];
    int64_t h_bits[C];
    unsigned long long mis;
    CU(cudaMemcpy(h_cut, d_cut, sizeof h_cut, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h_end, d_end, sizeof h_end, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h_peak, d_peak, sizeof h_peak, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h_bits, d_bits, sizeof h_bits, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(ovf, d_ovf, sizeof ovf, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&mis, d_mis, sizeof mis, cudaMemcpyDeviceToHost));
    if (ovf[0] || ovf[1] || mis) { fprintf(stderr, "encode flag %d, decode status %d, mismatches %llu\n", ovf[0], ovf[1], mis); return 1; }

    long long total = 0;
    for (int c = 0; c < C; ++c) {
        const uint8_t* x = h_sym + c * stride;
        long long a[S] = {0, 0, 0}, post[S] = {0, 0, 0};
        for (int t = 0; t < H; ++t) a[x[t] > S - 1 ? S - 1 : x[t]]++;
        int p = 0;                                             /* first argmax (functions_1.py:77) */
        for (int s = 1; s < S; ++s) if (a[s] > a[p]) p = s;
        for (int t = H; t < H + T / 2; ++t) post[x[t] > S - 1 ? S - 1 : x[t]]++;
        /* approx_sort for S = 3: peak 0 -> [0,1,2], peak 1 -> [1,0,2], peak 2 -> [2,1,0] (rank r holds symbol idx[r]) */
        const int idx[3][3] = {{0, 1, 2}, {1, 0, 2}, {2, 1, 0}};
        long long bits = 0;
        for (int r = 0; r < S; ++r) bits += post[idx[p][r]] * lens[r];
        if (h_cut[c] != H || h_end[c] != H + T / 2 || h_peak[c] != p || h_bits[c] != bits) {
            fprintf(stderr, "channel %d: cutoff %d end %d peak %d bits %lld, expected %d %d %d %lld\n", c, h_cut[c], h_end[c],
                    h_peak[c], (long long)h_bits[c], H, H + T / 2, p, bits);
            return 1;
        }
        total += bits;
    }
    printf("c_abi ok: %d channels x %d bins, %lld bits, lossless, abi %d\n", C, T, total, mua_abi_version());
    return 0;
}
