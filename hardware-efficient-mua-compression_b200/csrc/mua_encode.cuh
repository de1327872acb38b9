// Stage 5: Huffman encode.  One warp per channel; rows are staged into shared memory with TMA 1-D
// bulk copies (4-stage ring per warp); every lane codes 32 consecutive symbols through a nibble-pair
// LUT (saturation + approx-sort rank map + codeword folded in), a warp prefix-sum of the bit lengths
// gives each lane its bit offset, lanes funnel-shift their bits into a per-warp staging ring and the
// warp flushes complete 128-bit units with coalesced 16-byte stores.
#pragma once
#include "mua_common.cuh"

namespace mua {

struct EncParams {
    Layout L;
    int32_t S;
    const int32_t* start;
    const int32_t* end;
    const uint8_t* peak;
    const uint8_t* enc;
    const uint8_t* tab;   // table block
    int32_t K, Lmax;      // host-side view of the table block (cross-checked below)
    uint8_t* stream;
    int64_t slot_bytes;
    uint32_t* chunk_off;
    int32_t chunk_stride;
    int64_t* total_bits;
    int32_t* overflow;
};

constexpr int ENC_WARPS = 8;
constexpr int ENC_NST = 4;   // TMA stages per warp

template <int RW>
struct EncSmem {
    static constexpr int IN = 0;                               // ENC_NST * TILE bytes
    static constexpr int LUT2 = IN + ENC_NST * TILE;           // 256 * 8
    static constexpr int LUT1 = LUT2 + 2048;                   // 16 * 4
    static constexpr int RING = LUT1 + 64;                     // RW * 4
    static constexpr int BARS = RING + RW * 4;                 // ENC_NST * 8
    static constexpr int PER_WARP = (BARS + ENC_NST * 8 + 127) / 128 * 128;
};

// inclusive warp scan
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// RW = staging ring words (power of two, > TILE*Lmax/32 + 4); FAST2 = Lmax <= 2 register path
template <int RW, bool FAST2>
__global__ void __launch_bounds__(ENC_WARPS * 32, FAST2 ? 4 : 2) k_encode(const __grid_constant__ EncParams P) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    using SM = EncSmem<RW>;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sm = smem_raw + warp * SM::PER_WARP;
    uint8_t* s_in = sm + SM::IN;
    uint8_t* s_lut2 = sm + SM::LUT2;
    uint32_t* s_lut1 = reinterpret_cast<uint32_t*>(sm + SM::LUT1);
    uint32_t* s_ring = reinterpret_cast<uint32_t*>(sm + SM::RING);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(sm + SM::BARS);

    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < ENC_NST; ++i) mbar_init(&s_bar[i], 1);
        fence_barrier_init();
    }
    __syncwarp();

    const TabHdr* T = reinterpret_cast<const TabHdr*>(P.tab);
    const int K = T->K;
    if (T->S != P.S || K != P.K || T->Lmax != P.Lmax || (FAST2 && T->Lmax > 2) || TILE * T->Lmax > (RW - 8) * 32) {
        if (threadIdx.x == 0) *P.overflow = 2;   // launch configuration does not match the table block
        return;
    }
    const uint32_t* g_enc1 = reinterpret_cast<const uint32_t*>(P.tab + T->enc1_off);
    const uint4* g_enc2 = reinterpret_cast<const uint4*>(P.tab + T->enc2_off);

    const int gwarp = blockIdx.x * ENC_WARPS + warp, nwarps = gridDim.x * ENC_WARPS;
    uint32_t tcount = 0;   // tiles consumed by this warp so far (ring slot / parity)
    int cur_combo = -1;
    constexpr uint32_t RM = RW - 1;

    for (int c = gwarp; c < P.L.C; c += nwarps) {
        const int n = ch_len(P.L, c);
        const int start = P.start[c];
        const int end = min(P.end[c], n);
        uint32_t Pbits = 0;
        if (end > start && start >= 0) {
            const int combo = (int)P.peak[c] * K + (int)P.enc[c];
            if (combo != cur_combo) {   // per-warp copy of this (peak, codebook) pair's LUTs
                __syncwarp();
                const uint4* src = g_enc2 + (size_t)combo * 128;
#pragma unroll
                for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(s_lut2)[lane + 32 * i] = src[lane + 32 * i];
                if (lane < 16) s_lut1[lane] = g_enc1[(size_t)combo * 16 + lane];
                cur_combo = combo;
                __syncwarp();
            }
            const uint8_t* row = P.L.sym + ch_off(P.L, c);
            const int A0 = start & ~(TILE - 1);
            const int nt = (end - A0 + TILE - 1) / TILE;
            const int rd_end = (end + 15) & ~15;
            uint32_t* co = P.chunk_off + (size_t)c * P.chunk_stride;
            uint8_t* out = P.stream + (size_t)c * P.slot_bytes;
            uint32_t carry = 0;

            if (lane == 0) {
                const int npro = nt < ENC_NST ? nt : ENC_NST;
                for (int t = 0; t < npro; ++t) {
                    const uint32_t g = tcount + t;
                    const int ts = A0 + t * TILE;
                    const uint32_t bytes = (uint32_t)min(TILE, rd_end - ts);
                    mbar_expect_tx(&s_bar[g % ENC_NST], bytes);
                    tma_load_1d(s_in + (g % ENC_NST) * TILE, row + ts, bytes, &s_bar[g % ENC_NST]);
                }
            }

            for (int t = 0; t < nt; ++t) {
                const uint32_t g = tcount + t;
                const int slot = g % ENC_NST;
                mbar_wait(&s_bar[slot], (g / ENC_NST) & 1);
                const int ts = A0 + t * TILE;
                const uint4* src = reinterpret_cast<const uint4*>(s_in + slot * TILE + lane * 32);
                const uint4 q0 = src[0], q1 = src[1];
                uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                if (lane == 0) co[t] = Pbits;
                const bool full = (ts >= start) && (ts + TILE <= end);
                uint32_t Ttile;

                if (FAST2 && full) {
                    // ---------- register fast path: 32 symbols, <= 64 bits per lane ----------
                    uint32_t hi = (w[0] | w[1] | w[2]) | (w[3] | w[4] | w[5]) | (w[6] | w[7]);
                    if (__any_sync(FULL, (hi & 0xF0F0F0F0u) != 0)) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {   // bytes >= 16 -> 15 (the LUT saturates the rest)
                            uint32_t h4 = w[j] & 0xF0F0F0F0u;
                            uint32_t t1 = h4 | (h4 >> 1);
                            t1 |= t1 >> 2;
                            uint32_t mk = ((t1 >> 4) & 0x01010101u) * 0xFFu;
                            w[j] = (w[j] | mk) & 0x0F0F0F0Fu;
                        }
                    }
                    uint32_t qc[8], ql[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t y = ((w[j] << 3) | (w[j] >> 1)) & 0x07F807F8u;
                        const uint2 e0 = *reinterpret_cast<const uint2*>(s_lut2 + (y & 0xFFFFu));
                        const uint2 e1 = *reinterpret_cast<const uint2*>(s_lut2 + (y >> 16));
                        qc[j] = (e0.x << e1.y) | e1.x;
                        ql[j] = e0.y + e1.y;
                    }
#pragma unroll
                    for (int j = 0; j < 8; j += 2) { qc[j] = (qc[j] << ql[j + 1]) | qc[j + 1]; ql[j] += ql[j + 1]; }
#pragma unroll
                    for (int j = 0; j < 8; j += 4) { qc[j] = (qc[j] << ql[j + 2]) | qc[j + 2]; ql[j] += ql[j + 2]; }
                    const uint32_t nb = ql[0] + ql[4];                       // 32..64
                    unsigned long long acc = ((unsigned long long)qc[0] << ql[4]) | qc[4];
                    acc <<= (64 - nb);                                       // left-align
                    const uint32_t Ahi = (uint32_t)(acc >> 32), Alo = (uint32_t)acc;
                    const uint32_t incl = warp_incl_scan(nb, lane);
                    Ttile = __shfl_sync(FULL, incl, 31);
                    const uint32_t a = Pbits + incl - nb;
                    const uint32_t sh = a & 31, Wi = a >> 5;
                    uint32_t w0 = Ahi >> sh;
                    const uint32_t w1 = __funnelshift_r(Alo, Ahi, sh);
                    const uint32_t w2 = __funnelshift_r(0u, Alo, sh);
                    const uint32_t e = sh + nb;
                    const uint32_t nfull = e >> 5;                           // 1 or 2
                    const uint32_t tl = (e & 31) ? (nfull == 1 ? w1 : w2) : 0u;
                    uint32_t prev = __shfl_up_sync(FULL, tl, 1);
                    if (lane == 0) prev = carry;
                    w0 |= prev;
                    s_ring[Wi & RM] = w0;
                    if (nfull == 2) s_ring[(Wi + 1) & RM] = w1;
                    carry = __shfl_sync(FULL, tl, 31);
                } else {
                    // ---------- general path: sequential bit writer per lane ----------
                    const int p0 = ts + lane * 32;           // absolute position of this lane's first symbol
                    const uint8_t* sbytes = s_in + slot * TILE + lane * 32;
                    const int vlo = max(start - p0, 0), vhi = min(end - p0, 32);   // valid symbols [vlo, vhi)
                    uint32_t nb = 0;
                    if (full) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            uint32_t wj = w[j];
                            if ((wj & 0xF0F0F0F0u) != 0) {
                                uint32_t h4 = wj & 0xF0F0F0F0u, t1 = h4 | (h4 >> 1);
                                t1 |= t1 >> 2;
                                wj = (wj | (((t1 >> 4) & 0x01010101u) * 0xFFu)) & 0x0F0F0F0Fu;
                                w[j] = wj;
                            }
                            const uint32_t y = ((wj << 3) | (wj >> 1)) & 0x07F807F8u;
                            nb += reinterpret_cast<const uint2*>(s_lut2 + (y & 0xFFFFu))->y;
                            nb += reinterpret_cast<const uint2*>(s_lut2 + (y >> 16))->y;
                        }
                    } else {
                        for (int i = vlo; i < vhi; ++i) nb += s_lut1[min((uint32_t)sbytes[i], 15u)] >> 16;
                    }
                    const uint32_t incl = warp_incl_scan(nb, lane);
                    Ttile = __shfl_sync(FULL, incl, 31);
                    const uint32_t a = Pbits + incl - nb;
                    const uint32_t Wi = a >> 5;
                    unsigned long long pend = 0;
                    int fill = (int)(a & 31);
                    uint32_t first = 0;
                    int nemit = 0;
                    auto append = [&](uint32_t code, int len) {
                        pend |= (unsigned long long)code << (64 - fill - len);
                        fill += len;
                        if (fill >= 32) {
                            const uint32_t word = (uint32_t)(pend >> 32);
                            if (nemit == 0) first = word;
                            else s_ring[(Wi + nemit) & RM] = word;
                            ++nemit;
                            pend <<= 32;
                            fill -= 32;
                        }
                    };
                    if (full) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint32_t y = ((w[j] << 3) | (w[j] >> 1)) & 0x07F807F8u;
                            const uint2 e0 = *reinterpret_cast<const uint2*>(s_lut2 + (y & 0xFFFFu));
                            const uint2 e1 = *reinterpret_cast<const uint2*>(s_lut2 + (y >> 16));
                            append(e0.x, (int)e0.y);
                            append(e1.x, (int)e1.y);
                        }
                    } else {
                        for (int i = vlo; i < vhi; ++i) {
                            const uint32_t e1 = s_lut1[min((uint32_t)sbytes[i], 15u)];
                            append(e1 & 0xFFFFu, (int)(e1 >> 16));
                        }
                    }
                    // tails: a lane that emitted no word passes the incoming partial word through
                    uint32_t v = fill > 0 ? (uint32_t)(pend >> 32) : 0u;
                    int f = nemit > 0;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t tv = __shfl_up_sync(FULL, v, d);
                        const int tf = __shfl_up_sync(FULL, f, d);
                        if (lane >= d && !f) { v |= tv; f |= tf; }
                    }
                    if (!f) v |= carry;
                    uint32_t incoming = __shfl_up_sync(FULL, v, 1);
                    if (lane == 0) incoming = carry;
                    if (nemit > 0) s_ring[Wi & RM] = first | incoming;
                    carry = __shfl_sync(FULL, v, 31);
                }

                // ---------- flush complete 128-bit units ----------
                __syncwarp();
                const uint32_t Pnew = Pbits + Ttile;
                const uint32_t u1 = Pnew >> 7;
                for (uint32_t u = (Pbits >> 7) + lane; u < u1; u += 32) {
                    uint4 v4 = *reinterpret_cast<const uint4*>(&s_ring[(u * 4) & RM]);
                    v4.x = bswap32(v4.x); v4.y = bswap32(v4.y); v4.z = bswap32(v4.z); v4.w = bswap32(v4.w);
                    if ((int64_t)(u + 1) * 16 <= P.slot_bytes) *reinterpret_cast<uint4*>(out + (size_t)u * 16) = v4;
                    else *P.overflow = 1;
                }
                Pbits = Pnew;
                __syncwarp();
                if (lane == 0 && t + ENC_NST < nt) {   // refill the slot just consumed
                    const int ts2 = A0 + (t + ENC_NST) * TILE;
                    const uint32_t bytes = (uint32_t)min(TILE, rd_end - ts2);
                    mbar_expect_tx(&s_bar[slot], bytes);
                    tma_load_1d(s_in + slot * TILE, row + ts2, bytes, &s_bar[slot]);
                }
            }
            tcount += nt;

            // ---------- last partial unit, zero padded to 128 bits ----------
            if (Pbits & 127) {
                const uint32_t u = Pbits >> 7, wfull = Pbits >> 5;
                const uint32_t wi = u * 4 + (lane & 3);
                uint32_t val = wi < wfull ? s_ring[wi & RM] : (wi == wfull ? carry : 0u);
                val = bswap32(val);
                uint4 v4;
                v4.x = __shfl_sync(FULL, val, 0);
                v4.y = __shfl_sync(FULL, val, 1);
                v4.z = __shfl_sync(FULL, val, 2);
                v4.w = __shfl_sync(FULL, val, 3);
                if (lane == 0) {
                    if ((int64_t)(u + 1) * 16 <= P.slot_bytes) *reinterpret_cast<uint4*>(out + (size_t)u * 16) = v4;
                    else *P.overflow = 1;
                }
            }
            __syncwarp();
        }
        if (lane == 0) P.total_bits[c] = Pbits;
    }
}

}  // namespace mua
