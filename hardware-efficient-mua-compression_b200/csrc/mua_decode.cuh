// Stage 6: table-driven chunk-parallel decode (one lane per 1024-symbol chunk, multi-symbol LUT),
// the device-side round-trip check, the synthetic MUA generator and the binning kernels (stage 1).
#pragma once
#include "mua_common.cuh"

namespace mua {

struct DecParams {
    const uint8_t* stream;
    int64_t slot_bytes;
    const uint32_t* chunk_off;
    int32_t chunk_stride;
    const int64_t* off;
    int64_t stride;
    int32_t C, S;
    const int32_t* start;
    const int32_t* end;
    const uint8_t* peak;
    const uint8_t* enc;
    const uint8_t* tab;
    int32_t K, Lmax;
    uint8_t* dec;
};

constexpr int DEC_THREADS = 128;

__device__ __forceinline__ void dec_store_bytes(uint8_t* dst, unsigned long long lo, unsigned long long hi, int from, int to) {
    for (int k = from; k < to; ++k) dst[k] = (uint8_t)((k < 8 ? lo >> (8 * k) : hi >> (8 * (k - 8))) & 0xFF);
}

template <bool SMEM_LUT>
__global__ void __launch_bounds__(DEC_THREADS) k_decode(const __grid_constant__ DecParams P) {
    extern __shared__ __align__(16) unsigned long long s_lut[];
    const TabHdr* T = reinterpret_cast<const TabHdr*>(P.tab);
    const int K = T->K, W = T->W;
    if (T->S != P.S || K != P.K || T->Lmax != P.Lmax) return;   // host view does not match the table block
    const unsigned long long* g_lut = reinterpret_cast<const unsigned long long*>(P.tab + T->dec_off);
    if (SMEM_LUT) {
        const int nent = (T->S * K) << W;
        for (int i = threadIdx.x; i < nent; i += blockDim.x) s_lut[i] = g_lut[i];
        __syncthreads();
    }
    const long long nitems = (long long)P.C * P.chunk_stride;
    for (long long item = (long long)blockIdx.x * blockDim.x + threadIdx.x; item < nitems;
         item += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(item / P.chunk_stride), j = (int)(item % P.chunk_stride);
        const int start = P.start[c], end = P.end[c];
        if (end <= start || start < 0) continue;
        const int j0 = start / TILE;
        const int nch = (end + TILE - 1) / TILE - j0;
        if (j >= nch) continue;
        const int a = max(start, (j0 + j) * TILE), b = min(end, (j0 + j + 1) * TILE);
        int rem = b - a;
        const unsigned long long* lut = (SMEM_LUT ? s_lut : g_lut) + ((size_t)((int)P.peak[c] * K + (int)P.enc[c]) << W);
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(P.stream + (size_t)c * P.slot_bytes);
        const uint32_t nwords = (uint32_t)(P.slot_bytes >> 2);
        const uint32_t bitpos = P.chunk_off[(size_t)c * P.chunk_stride + j];
        uint32_t widx = bitpos >> 5;
        const int sh = bitpos & 31;
        auto ldw = [&](uint32_t i) -> uint32_t { return i < nwords ? bswap32(__ldg(sw + i)) : 0u; };
        unsigned long long buf = ((unsigned long long)ldw(widx) << 32) | ldw(widx + 1);
        widx += 2;
        buf <<= sh;
        int avail = 64 - sh;

        const int64_t row = P.off ? P.off[c] : (int64_t)c * P.stride;
        const int lead = a & 15;
        uint8_t* dst = P.dec + row + (a - lead);   // 16-byte aligned
        unsigned long long lo = 0, hi = 0, ex = 0;
        int oc = lead;
        bool first = lead > 0;
        while (rem > 0) {
            if (avail < 32) {
                buf |= (unsigned long long)ldw(widx++) << (32 - avail);
                avail += 32;
            }
            const unsigned long long e = lut[buf >> (64 - W)];
            int nsy = (int)((e >> 56) & 15);
            const int used = (int)(e >> 60);
            unsigned long long syms = e & 0x00FFFFFFFFFFFFFFull;
            if (nsy == 0) break;   // corrupt table/stream: never loops forever
            if (nsy > rem) {
                nsy = rem;
                syms &= (1ull << (8 * nsy)) - 1ull;
            }
            buf <<= used;
            avail -= used;
            rem -= nsy;
            if (oc < 8) {
                lo |= syms << (8 * oc);
                if (oc) hi |= syms >> (64 - 8 * oc);
            } else {
                const int o2 = oc - 8;
                hi |= syms << (8 * o2);
                if (o2) ex |= syms >> (64 - 8 * o2);
            }
            oc += nsy;
            if (oc >= 16) {
                if (!first) *reinterpret_cast<uint4*>(dst) = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
                else dec_store_bytes(dst, lo, hi, lead, 16);
                first = false;
                dst += 16;
                lo = ex; hi = 0; ex = 0;
                oc -= 16;
            }
        }
        if (oc > 0) dec_store_bytes(dst, lo, hi, first ? lead : 0, oc);
    }
}

// ---- round-trip check: dec == min(sym, S-1) on [start, end) ------------------------------------
__global__ void __launch_bounds__(256) k_verify(const uint8_t* __restrict__ sym, const uint8_t* __restrict__ dec,
                                                const int64_t* __restrict__ off, int64_t stride, int C, int S,
                                                const int32_t* __restrict__ start, const int32_t* __restrict__ end,
                                                unsigned long long* __restrict__ mismatch) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nw = gridDim.x * (blockDim.x >> 5);
    unsigned long long bad = 0;
    for (int c = blockIdx.x * (blockDim.x >> 5) + warp; c < C; c += nw) {
        const int64_t row = off ? off[c] : (int64_t)c * stride;
        const int a = start[c], b = end[c];
        if (a < 0) continue;
        for (int t = a + lane; t < b; t += 32) {
            const int s = min((int)sym[row + t], S - 1);
            bad += (s != (int)dec[row + t]);
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) bad += __shfl_xor_sync(FULL, bad, d);
    if (lane == 0 && bad) atomicAdd(mismatch, bad);
}

// ---- synthetic MUA (oracle/mua_oracle.py:synth_symbols) ----------------------------------------
__global__ void __launch_bounds__(256) k_synth(uint8_t* __restrict__ sym, int64_t stride, int T, int C, int64_t c0,
                                               uint32_t seed, const uint32_t* __restrict__ thr, int bursty) {
    __shared__ uint32_t s_thr[256 * 24];
    for (int i = threadIdx.x; i < 256 * 24; i += blockDim.x) s_thr[i] = thr[i];
    __syncthreads();
    const int groups = (T + 15) / 16;
    const long long total = (long long)C * groups;
    for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < total; it += (long long)gridDim.x * blockDim.x) {
        const int cl = (int)(it / groups), g = (int)(it % groups);
        const uint32_t ch = (uint32_t)(c0 + cl);
        int cls = (int)(mix32(seed * 0x9E3779B9u + ch) & 255u);
        const uint32_t hc = mix32(ch ^ 0x68E31DA4u);
        if (bursty) {
            const uint32_t hb = mix32(seed ^ hc ^ ((uint32_t)g * 0x85EBCA6Bu) ^ 0xB5297A4Du);
            if (hb < 390451572u) cls = min(cls + 96, 255);
        }
        const uint32_t* th = s_thr + cls * 24;
        uint32_t out[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint32_t t = (uint32_t)(g * 16 + k);
            const uint32_t u = mix32((hc + t * 0x9E3779B1u) ^ seed);
            uint32_t v = 0;
            while (v < 24 && u >= th[v]) ++v;
            if ((int)t < T) out[k >> 2] |= v << (8 * (k & 3));
        }
        *reinterpret_cast<uint4*>(sym + (int64_t)cl * stride + g * 16) = make_uint4(out[0], out[1], out[2], out[3]);
    }
}

// ---- stage 1: binning -----------------------------------------------------------------------
// counts[b][c] = sum of raster[b*r .. min((b+1)*r, T0))[c]  (functions_1.py:11-24), int64 like astype(int)
template <typename TIn, typename TAcc>
__global__ void __launch_bounds__(256) k_bin_counts(const TIn* __restrict__ raster, int64_t T0, int C, int r, int64_t nb,
                                                    int64_t* __restrict__ counts) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    for (int64_t b = blockIdx.y; b < nb; b += gridDim.y) {
        const int64_t r0 = b * r, r1 = min(r0 + (int64_t)r, T0);
        TAcc acc = 0;
        for (int64_t t = r0; t < r1; ++t) acc += (TAcc)raster[t * C + c];
        counts[b * C + c] = (int64_t)acc;
    }
}

// uint8 raster [T0][C] -> channel-major saturated symbols [C][stride]; tile = 128 channels x 64 bins,
// coalesced 4-byte reads along channels, transposed through shared memory, 16-byte writes along bins.
constexpr int BIN_TC = 128, BIN_TB = 64, BIN_LD = 80;
__global__ void __launch_bounds__(256) k_bin_sym(const uint8_t* __restrict__ raster, int64_t T0, int C, int r, int64_t nb,
                                                 uint8_t* __restrict__ sym, int64_t stride, int sat) {
    __shared__ __align__(16) uint8_t tile[BIN_TC][BIN_LD];
    const int c0 = blockIdx.x * BIN_TC;
    const int64_t b0 = (int64_t)blockIdx.y * BIN_TB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = c0 + 4 * lane;
    const bool vec = (C % 4 == 0) && (c + 3 < C);
    for (int i = 0; i < BIN_TB / 8; ++i) {
        const int bl = warp + 8 * i;
        const int64_t b = b0 + bl;
        uint32_t acc[4] = {0, 0, 0, 0};
        if (b < nb) {
            const int64_t r0 = b * r, r1 = min(r0 + (int64_t)r, T0);
            if (vec) {
                for (int64_t t = r0; t < r1; ++t) {
                    const uint32_t v = *reinterpret_cast<const uint32_t*>(raster + t * C + c);
                    acc[0] += v & 0xFF; acc[1] += (v >> 8) & 0xFF; acc[2] += (v >> 16) & 0xFF; acc[3] += v >> 24;
                }
            } else {
                for (int64_t t = r0; t < r1; ++t)
                    for (int k = 0; k < 4; ++k)
                        if (c + k < C) acc[k] += raster[t * C + c + k];
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) tile[4 * lane + k][bl] = (uint8_t)min(acc[k], (uint32_t)sat);
    }
    __syncthreads();
    const int rowi = threadIdx.x >> 1, half = threadIdx.x & 1;
    const int cc = c0 + rowi;
    if (cc < C) {
        const int64_t bs = b0 + half * 32;
        uint8_t* dst = sym + (int64_t)cc * stride + bs;
        const uint8_t* src = &tile[rowi][half * 32];
        if ((stride % 16 == 0) && bs + 32 <= nb) {
            reinterpret_cast<uint4*>(dst)[0] = reinterpret_cast<const uint4*>(src)[0];
            reinterpret_cast<uint4*>(dst)[1] = reinterpret_cast<const uint4*>(src)[1];
        } else {
            for (int k = 0; k < 32 && bs + k < nb; ++k) dst[k] = src[k];
        }
    }
}

}  // namespace mua
