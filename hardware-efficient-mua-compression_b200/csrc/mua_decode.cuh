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
    int32_t item_chunks;        // chunks per channel that can be non-empty (<= chunk_stride)
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

// ---- general decoder (any codebook) ----
// One lane decodes one 1024-symbol chunk; a warp owns 32 consecutive chunks.
//   * the lanes' stream bytes are staged into shared memory by the whole warp with coalesced 16-byte
//     loads (272 B per lane and stage, byte-swapped to MSB-first words on the way in);
//   * every LUT lookup decodes exactly NSYM symbols (window W = NSYM*Lmax bits), so output words are
//     produced at fixed positions -- no variable-length output assembly;
//   * decoded symbols go to a padded shared-memory tile (128 B per lane and period) that the warp
//     writes out with coalesced 16-byte stores.
constexpr int DG_WARPS = 4;
constexpr int DG_STR_W = 68;          // staged stream words per lane: 272 B = 128 bits of alignment slack + 2048 bits
constexpr int DG_OUT_B = 144;         // output tile row: 128 B + 16 B pad
constexpr int DG_PER_WARP = 32 * DG_STR_W * 4 + 32 * DG_OUT_B + 16;   // + mbarrier

template <int NSYM, bool SMEM_LUT>
__global__ void __launch_bounds__(DG_WARPS * 32, 3) k_decode_gen(const __grid_constant__ DecParams P) {
    extern __shared__ __align__(128) uint8_t dsm[];
    const TabHdr* T = reinterpret_cast<const TabHdr*>(P.tab);
    const int K = T->K, W = T->W;
    if (T->S != P.S || K != P.K || T->Lmax != P.Lmax || T->nsym != NSYM) return;   // host view does not match the table block
    const uint32_t* g_lut = reinterpret_cast<const uint32_t*>(P.tab + T->dec_off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* s_str = reinterpret_cast<uint32_t*>(dsm + warp * DG_PER_WARP);
    uint8_t* s_out = dsm + warp * DG_PER_WARP + 32 * DG_STR_W * 4;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_out + 32 * DG_OUT_B);
    const uint32_t* s_lut = reinterpret_cast<const uint32_t*>(dsm + DG_WARPS * DG_PER_WARP);
    if (lane == 0) {
        mbar_init(s_bar, 1);
        fence_barrier_init();
    }
    if (SMEM_LUT) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(dsm + DG_WARPS * DG_PER_WARP);
        const int nent = (T->S * K) << W;
        for (int i = threadIdx.x; i < nent; i += blockDim.x) dst[i] = g_lut[i];
    }
    __syncthreads();
    uint32_t parity = 0;
    const int periods_per_stage = 2048 / (128 * T->Lmax);        // 128-symbol periods one staged row is good for
    const long long nitems = (long long)P.C * P.item_chunks;
    const long long ngroups = (nitems + 31) / 32;
    const uint32_t slot_bytes = (uint32_t)P.slot_bytes;

    for (long long g = (long long)blockIdx.x * DG_WARPS + warp; g < ngroups; g += (long long)gridDim.x * DG_WARPS) {
        // ---- this lane's chunk ----
        const long long item = g * 32 + lane;
        int rem = 0;
        uint32_t bitpos = 0;
        const uint8_t* sbase = P.stream;
        uint8_t* optr = P.dec;
        const uint32_t* lut = SMEM_LUT ? s_lut : g_lut;
        if (item < nitems) {
            const int c = (int)(item / P.item_chunks), j = (int)(item % P.item_chunks);
            const int start = P.start[c], end = P.end[c];
            if (end > start && start >= 0) {
                const int j0 = start / TILE;
                const int nch = (end + TILE - 1) / TILE - j0;
                if (j < nch) {
                    const int a = max(start, (j0 + j) * TILE), b = min(end, (j0 + j + 1) * TILE);
                    rem = b - a;
                    bitpos = P.chunk_off[(size_t)c * P.chunk_stride + j];
                    sbase = P.stream + (size_t)c * P.slot_bytes;
                    optr = P.dec + (P.off ? P.off[c] : (int64_t)c * P.stride) + a;
                    lut += (size_t)((int)P.peak[c] * K + (int)P.enc[c]) << W;
                }
            }
        }
        int done = 0;                                            // symbols already written out
        while (__any_sync(FULL, rem > 0)) {
            // ---- stage 272 stream bytes per lane (one TMA bulk copy each), from the 16-byte unit holding `bitpos` ----
            const uint32_t cur_al = (bitpos >> 7) << 4;
            const uint32_t nbytes = rem > 0 ? min((uint32_t)(DG_STR_W * 4), slot_bytes - cur_al) : 0u;
            const uint32_t total = __reduce_add_sync(FULL, nbytes);
            __syncwarp();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (total) {
                if (lane == 0) mbar_expect_tx(s_bar, total);
                __syncwarp();
                if (nbytes) tma_load_1d(s_str + lane * DG_STR_W, sbase + cur_al, nbytes, s_bar);
                mbar_wait(s_bar, parity);
                parity ^= 1;
            }
            const uint32_t* rowp = s_str + lane * DG_STR_W;
            const uint32_t boff = bitpos - cur_al * 8;           // 0..127
            uint32_t rp = boff >> 5;
            uint32_t hi = bswap32(rowp[rp]), lo = bswap32(rowp[rp + 1]);
            rp += 2;
            uint32_t off = boff & 31;
            uint32_t consumed = 0;                               // bits consumed in this stage

            for (int per = 0; per < periods_per_stage && __any_sync(FULL, rem > 0); ++per) {
                // ---- 128 symbols per lane into the output tile ----
                uint4* orow = reinterpret_cast<uint4*>(s_out + lane * DG_OUT_B);
#pragma unroll 2
                for (int q = 0; q < 8; ++q) {
                    uint32_t ow[4];
                    if (NSYM == 4) {
                        // one 32-bit snapshot feeds 4 lookups of <= 8 bits; refill check once per 16 symbols
                        const uint32_t x = __funnelshift_l(lo, hi, off);
                        uint32_t o = 0;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint32_t e = lut[(x << o) >> (32 - W)];
                            ow[k] = e & 0x0F0F0F0Fu;
                            o += e >> 28;
                        }
                        off += o;
                        consumed += o;
                        if (off >= 32) { hi = lo; lo = bswap32(rowp[min(rp, (uint32_t)(DG_STR_W - 1))]); ++rp; off -= 32; }
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            uint32_t wsym = 0;
#pragma unroll
                            for (int h = 0; h < 4 / NSYM; ++h) {
                                const uint32_t x = __funnelshift_l(lo, hi, off);
                                const uint32_t e = lut[x >> (32 - W)];
                                wsym |= (e & 0x0F0F0F0Fu) << (8 * NSYM * h);
                                const uint32_t used = e >> 28;
                                off += used;
                                consumed += used;
                                if (off >= 32) { hi = lo; lo = bswap32(rowp[min(rp, (uint32_t)(DG_STR_W - 1))]); ++rp; off -= 32; }
                            }
                            ow[k] = wsym;
                        }
                    }
                    orow[q] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                }
                __syncwarp();
                // ---- coalesced write-out: 8 lanes per row, 4 rows per pass ----
                const int vrow_self = min(max(rem, 0), 128);     // valid bytes of my row in this period
#pragma unroll 1
                for (int i = 0; i < 8; ++i) {
                    const int r = i * 4 + (lane >> 3), col = lane & 7;
                    const int vr = __shfl_sync(FULL, vrow_self, r);
                    const unsigned long long dptr = __shfl_sync(FULL, reinterpret_cast<unsigned long long>(optr) + done, r);
                    if (col * 16 < vr) {
                        const uint8_t* sp = s_out + r * DG_OUT_B + col * 16;
                        uint8_t* d = reinterpret_cast<uint8_t*>(dptr) + col * 16;
                        if (col * 16 + 16 <= vr && (dptr & 15) == 0) {
                            *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(sp);
                        } else {   // window edge or unaligned first chunk: byte stores
                            const int nbyte = min(16, vr - col * 16);
                            for (int k = 0; k < nbyte; ++k) d[k] = sp[k];
                        }
                    }
                }
                __syncwarp();
                if (rem > 0) { rem -= 128; done += 128; }
            }
            bitpos += consumed;
            if (rem <= 0) rem = 0;
        }
    }
}

// ---- fast decoder (NSYM = 4: codebooks with Lmax <= 2) ----
// One lane decodes one 1024-symbol chunk; a warp owns 32 consecutive chunks.
//   * The kernel is bound by the LSU data pipe (shared-memory wavefronts), so everything that can stay
//     off that pipe does: every lane stages its whole chunk of the stream (<= 272 B: 2048 bits + 16-byte
//     alignment slack) with ONE TMA bulk copy (cp.async.bulk -> UBLKCP, mbarrier completion): no LSU
//     wavefronts, no registers, one latency exposure per 1024 symbols per lane;
//   * one 32-bit snapshot of the stream feeds 4 LUT lookups (<= 8 bits each, 4 symbols each, fixed
//     output positions); the refill test runs once per 16 symbols and the refill word is read one
//     snapshot ahead of use;
//   * decoded symbols go to a padded shared-memory tile (64 B per lane and period) that the warp
//     writes out with coalesced 16-byte stores (4 lanes per row).
constexpr int DF_WARPS = 4;
constexpr int DF_PER = 128;            // symbols per lane and period
constexpr int DF_ROW_B = 272;          // staged stream bytes per lane
constexpr int DF_STR_B = 32 * DF_ROW_B;
constexpr int DF_OUT_B = 144;          // output tile row: 128 B + 16 B pad
constexpr int DF_PER_WARP = DF_STR_B + 32 * DF_OUT_B + 16;   // + mbarrier

template <bool SMEM_LUT>
__global__ void __launch_bounds__(DF_WARPS * 32, 4) k_decode_fast(const __grid_constant__ DecParams P) {
    extern __shared__ __align__(128) uint8_t dsm[];
    const TabHdr* T = reinterpret_cast<const TabHdr*>(P.tab);
    const int K = T->K, W = T->W;
    if (T->S != P.S || K != P.K || T->Lmax != P.Lmax || T->nsym != 4 || T->Lmax > 2) return;   // host view does not match the table block
    const uint32_t* g_lut = reinterpret_cast<const uint32_t*>(P.tab + T->dec_off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* s_str = dsm + warp * DF_PER_WARP;
    uint8_t* s_out = s_str + DF_STR_B;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_out + 32 * DF_OUT_B);
    const uint32_t* s_lut = reinterpret_cast<const uint32_t*>(dsm + DF_WARPS * DF_PER_WARP);
    if (lane == 0) {
        mbar_init(s_bar, 1);
        fence_barrier_init();
    }
    if (SMEM_LUT) {
        uint32_t* dst = reinterpret_cast<uint32_t*>(dsm + DF_WARPS * DF_PER_WARP);
        const int nent = (T->S * K) << W;
        for (int i = threadIdx.x; i < nent; i += blockDim.x) dst[i] = g_lut[i];
    }
    __syncthreads();
    const long long nitems = (long long)P.C * P.item_chunks;
    const long long ngroups = (nitems + 31) / 32;
    const uint32_t slot_bytes = (uint32_t)P.slot_bytes;
    const int wsh = 32 - W;
    // write-out: 8 lanes per 128-byte row (one full line), 4 rows per pass
    const int wrow = lane >> 3, wcol = lane & 7;
    uint32_t parity = 0;

    for (long long g = (long long)blockIdx.x * DF_WARPS + warp; g < ngroups; g += (long long)gridDim.x * DF_WARPS) {
        // ---- this lane's chunk ----
        const long long item = g * 32 + lane;
        int rem = 0;
        uint32_t bp = 0;                                         // bit position in the channel's stream
        const uint8_t* sbase = P.stream;
        uint8_t* optr = P.dec;
        const uint32_t* lut = SMEM_LUT ? s_lut : g_lut;
        if (item < nitems) {
            const int c = (int)(item / P.item_chunks), j = (int)(item % P.item_chunks);
            const int start = P.start[c], end = P.end[c];
            if (end > start && start >= 0) {
                const int j0 = start / TILE;
                const int nch = (end + TILE - 1) / TILE - j0;
                if (j < nch) {
                    const int a = max(start, (j0 + j) * TILE), b = min(end, (j0 + j + 1) * TILE);
                    rem = b - a;
                    bp = P.chunk_off[(size_t)c * P.chunk_stride + j];
                    sbase = P.stream + (size_t)c * P.slot_bytes;
                    optr = P.dec + (P.off ? P.off[c] : (int64_t)c * P.stride) + a;
                    lut += (size_t)((int)P.peak[c] * K + (int)P.enc[c]) << W;
                }
            }
        }
        // ---- stage the chunk's stream bytes: one TMA bulk copy per lane ----
        const uint32_t al = (bp >> 7) << 4;                      // 16-byte aligned byte offset in the slot
        const uint32_t nbytes = rem > 0 ? min((uint32_t)DF_ROW_B, slot_bytes - al) : 0u;
        const uint32_t total = __reduce_add_sync(FULL, nbytes);
        __syncwarp();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of the rows vs. the async writes
        if (total) {
            if (lane == 0) mbar_expect_tx(s_bar, total);
            __syncwarp();
            if (nbytes) tma_load_1d(s_str + lane * DF_ROW_B, sbase + al, nbytes, s_bar);
            mbar_wait(s_bar, parity);
            parity ^= 1;
        }
        const uint32_t* rowp = reinterpret_cast<const uint32_t*>(s_str + lane * DF_ROW_B);
        const uint32_t boff = bp & 127;
        uint32_t rp = boff >> 5;
        uint32_t hi = bswap32(rowp[rp]), lo = bswap32(rowp[rp + 1]), nx = bswap32(rowp[rp + 2]);
        rp += 3;
        uint32_t off = boff & 31;

        while (__any_sync(FULL, rem > 0)) {
            // ---- DF_PER symbols per lane into the output tile ----
            uint4* orow = reinterpret_cast<uint4*>(s_out + lane * DF_OUT_B);
#pragma unroll
            for (int q = 0; q < DF_PER / 16; ++q) {
                uint32_t ow[4];
                const uint32_t x = __funnelshift_l(lo, hi, off);
                uint32_t o = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t e = lut[(x << o) >> wsh];
                    ow[k] = e & 0x0F0F0F0Fu;
                    o += e >> 28;
                }
                off += o;
                if (off >= 32) { hi = lo; lo = nx; nx = bswap32(rowp[min(rp, (uint32_t)(DF_ROW_B / 4 - 1))]); ++rp; off -= 32; }
                orow[q] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
            }
            __syncwarp();
            // ---- coalesced write-out: 4 lanes per row, 8 rows per pass ----
            const int vrow_self = min(max(rem, 0), DF_PER);      // valid bytes of my row in this period
            const unsigned long long optr_self = reinterpret_cast<unsigned long long>(optr);
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
                const int r = i * 4 + wrow;
                const int vr = __shfl_sync(FULL, vrow_self, r);
                const unsigned long long dptr = __shfl_sync(FULL, optr_self, r);
                if (wcol * 16 < vr) {
                    const uint8_t* sp = s_out + r * DF_OUT_B + wcol * 16;
                    uint8_t* d = reinterpret_cast<uint8_t*>(dptr) + wcol * 16;
                    if (wcol * 16 + 16 <= vr && (dptr & 15) == 0) {
                        *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(sp);
                    } else {   // window edge or unaligned first chunk: byte stores
                        const int nbyte = min(16, vr - wcol * 16);
                        for (int k = 0; k < nbyte; ++k) d[k] = sp[k];
                    }
                }
            }
            __syncwarp();
            rem -= DF_PER;
            optr += DF_PER;
        }
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
