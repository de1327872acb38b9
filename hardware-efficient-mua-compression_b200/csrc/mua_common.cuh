// Shared device helpers and the device table block of the MUA compression path (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mua_b200.h"

namespace mua {

constexpr unsigned FULL = 0xFFFFFFFFu;
constexpr int TILE = MUA_CHUNK;        // symbols per warp tile == decode chunk

// ---- device table block (built by k_build_tables) ------------------------------------------
// enc1 : uint32 [S][K][16]    raw nibble b -> (len << 16 | code) of rank[p][min(b,S-1)]
// enc2 : uint32 [S][K][256]   raw nibble pair (b0 | b1<<4), b0 first in time -> code | len << 24   (general encoder)
// enc4 : [S][K] x 768 B       four saturated symbols per entry, only when Lmax <= 2 and S <= 3 (fast encoder):
//                             [0,128) codes and [128,256) lengths indexed in base S (q0 + S*q1 + S^2*q2 + S^3*q3, q0 first
//                             in time; S^4 <= 81 entries = one word per bank, conflict-free); [256,512) codes and [512,768)
//                             lengths indexed in base S+1, digit S = "outside the window" without bits (partial tiles)
// encp : [S][K] x 512 B      two saturated symbols (or the null symbol S = "outside the window": no bits), index
//                             i = q1 + (S+1)*q0 (q0 first in time): uint16 code at byte 2i, uint8 length at byte
//                             256 + 2i; only when Lmax <= 8 (pair encoder)
// dec  : uint32 [S][K][1<<W]  W = nsym*Lmax bit window -> nsym symbols, one per byte (first symbol in
//                             byte 0, values < 16) | used_bits << 28
// decv : uint32 [K][1<<Wv]     Wv-bit window -> as many whole symbols as fit, at most 4, as RANKS of codebook row k:
//                             bits [15:0] four PRMT byte selectors into the lane's rank -> symbol map (unused slots
//                             select a zero byte: 0x8 = sign replicate for S <= 8, 0xF = entry 15 of the 16-byte
//                             map for S >= 9), [18:16] symbol count (>= 1: Wv >= Lmax), [23:20] bits used
struct TabHdr {
    int32_t S, K, Lmax, W;
    int32_t nsym;                    // symbols decoded per LUT lookup (4, 2 or 1)
    int32_t enc1_off, enc2_off, enc4_off, dec_off, total_bytes;
    int32_t encp_off;                // 0 when Lmax > 8
    int32_t decv_off, Wv;            // variable-count rank tables of the general decoder: uint32 [K][1 << Wv]
    int32_t pad[3];
    uint8_t lens[MUA_MAX_K][16];     // SCLV rows (Stored_SCLVs_S_<S>.pkl), ascending lengths
    uint16_t codes[MUA_MAX_K][16];   // codeword of rank r
    uint8_t rank[MUA_MAX_S][16];     // rank[p][s]: approx_sort permutation for peak p (functions_1.py:75-90)
    uint8_t idx[MUA_MAX_S][16];      // idx[p][r] = symbol of rank r (what approx_sort returns)
};

__host__ __device__ inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

// symbols per decode lookup (window W = nsym*Lmax bits): 4 when four lookups fit one 32-bit snapshot
// (W <= 8), else 2 (W <= 12), else 1 -- and only while the LUT set of all (peak,row) pairs stays small
__host__ __device__ inline int dec_nsym(int S, int K, int Lmax) {
    const int ns[2] = {4, 2};
    const int wmax[2] = {8, 12};
    for (int i = 0; i < 2; ++i) {
        const int W = ns[i] * Lmax;
        if (W <= wmax[i] && (long long)S * K * (1ll << W) * 4 <= 32 * 1024) return ns[i];
    }
    return 1;
}

// Closed form of approx_sort's permutation (functions_1.py:75-90).
// `p > len(hist)/2` is a true division in the reference (:78)  <=>  2p > S.
__host__ __device__ inline int rank_of(int p, int s, int S) {
    if (2 * p > S) {
        int d = S - 1 - p;
        if (s >= p) return 2 * (s - p);
        if (s >= p - d) return 2 * (p - s) - 1;
        return S - 1 - s;
    }
    if (s < p) return 2 * (p - s) - 1;
    if (s <= 2 * p) return 2 * (s - p);
    return s;
}

// window of the variable-count decode tables: as wide as keeps K tables within 64 KB (72 KB for S = 10), 9..12 bits
#ifndef MUA_DV_WMAX
#define MUA_DV_WMAX 12
#endif
__host__ __device__ inline int decv_window(int K, int Lmax) {
    int W = MUA_DV_WMAX;
    while (W > 9 && (long long)K * (1ll << W) * 4 > 64 * 1024) --W;
    return W < Lmax ? Lmax : W;
}

struct Layout {
    const uint8_t* sym;
    const int64_t* off;
    const int32_t* len;
    int64_t stride;
    int32_t T;
    int32_t C;
};
__device__ __forceinline__ int64_t ch_off(const Layout& L, int c) { return L.off ? L.off[c] : (int64_t)c * L.stride; }
__device__ __forceinline__ int ch_len(const Layout& L, int c) { return L.len ? L.len[c] : L.T; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier + TMA 1-D bulk copy (cp.async.bulk -> SASS UBLKCP) ----------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    uint32_t a = smem_u32(bar);
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!ok);
}

__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

// every byte -> 0xFF if its bit 7 is set, else 0x00 (PRMT sign-replicate mode; __byte_perm masks the
// replicate bit of the selector away, so this needs the PTX form)
__device__ __forceinline__ uint32_t byte_msb_mask(uint32_t v) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(r) : "r"(v));
    return r;
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}

// bytes >= v  ->  bit 7 of that byte (SWAR threshold compare, any byte value 0..255, 1 <= v <= 127)
__device__ __forceinline__ uint32_t ge_mask(uint32_t w, uint32_t lo7, int v) {
    return ((lo7 + (uint32_t)(0x80 - v) * 0x01010101u) | w) & 0x80808080u;
}

// volatile loads: issued where they are written (the compiler would otherwise sink them to their first use)
__device__ __forceinline__ int ldg_s32(const int32_t* p) { int v; asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint32_t ldg_u32(const uint32_t* p) { uint32_t v; asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ int ldg_u8(const uint8_t* p) { uint32_t v; asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p)); return (int)v; }
__device__ __forceinline__ long long ldg_s64(const int64_t* p) { long long v; asm volatile("ld.global.nc.s64 %0, [%1];" : "=l"(v) : "l"(p)); return v; }

// counter-based RNG of the synthetic generator (mirrors oracle/mua_oracle.py:_mix32)
__host__ __device__ inline uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7FEB352Du;
    x ^= x >> 15;
    x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}

}  // namespace mua
