// Stage 5: Huffman encode.  One warp per channel; rows are staged into shared memory with TMA 1-D
// bulk copies (4-stage ring per warp, mbarrier completion); every lane codes 32 consecutive symbols,
// a warp prefix-sum of the bit lengths gives each lane its bit offset, lanes funnel-shift their bits
// into a per-warp staging ring (the partial word between two lanes travels by shuffle, no atomics)
// and the warp flushes complete 128-bit units with coalesced 16-byte stores.
//
//   k_encode_fast : codebooks with Lmax <= 2 (S <= 4; the chosen system S=3 '0','10','11').  Bytes are
//                   saturated with 5 SWAR ops per 4 symbols, 4 symbols are gathered into one 8-bit
//                   index with one multiply and coded with ONE 16-bit LUT read; <= 64 bits per lane
//                   stay in registers.
//   k_encode_gen  : any codebook (Lmax <= 9): nibble-pair LUT + sequential per-lane bit writer.
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
    int32_t K, Lmax;      // host-side view of the table block (cross-checked in the kernel)
    uint8_t* stream;
    int64_t slot_bytes;
    uint32_t* chunk_off;
    int32_t chunk_stride;
    int64_t* total_bits;
    int32_t* overflow;
};

constexpr int ENC_WARPS = 8;
constexpr int ENC_NST = 4;   // TMA stages per warp

// Partial words between lanes when some lanes emit no complete word (masked head/tail tiles):
// segmented inclusive OR-scan of the lanes' trailing partial words; a lane that completed a word
// (`emits`) starts a new segment.  Returns the partial word arriving at this lane; updates carry.
__device__ __forceinline__ uint32_t tails_segmented(uint32_t own_tail, bool emits, uint32_t& carry, int lane) {
    uint32_t v = own_tail;
    int f = emits;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t tv = __shfl_up_sync(FULL, v, d);
        const int tf = __shfl_up_sync(FULL, f, d);
        if (lane >= d && !f) { v |= tv; f |= tf; }
    }
    if (!f) v |= carry;
    uint32_t incoming = __shfl_up_sync(FULL, v, 1);
    if (lane == 0) incoming = carry;
    carry = __shfl_sync(FULL, v, 31);
    return incoming;
}

// flush the complete 128-bit units in [Pold, Pnew) from the ring to the channel's slot (<= 32 units)
template <uint32_t RM>
__device__ __forceinline__ void flush_units(const uint32_t* s_ring, uint8_t* out, uint32_t Pold, uint32_t Pnew,
                                            uint32_t slot_units, int32_t* overflow, int lane) {
    const uint32_t u = (Pold >> 7) + lane;
    if (u < (Pnew >> 7)) {
        if (u < slot_units) {
            uint4 v4 = *reinterpret_cast<const uint4*>(&s_ring[(u * 4) & RM]);
            v4.x = bswap32(v4.x); v4.y = bswap32(v4.y); v4.z = bswap32(v4.z); v4.w = bswap32(v4.w);
            *reinterpret_cast<uint4*>(out + (size_t)u * 16) = v4;
        } else {
            *overflow = 1;
        }
    }
}

// last partial unit of a channel, zero padded to 128 bits
template <uint32_t RM>
__device__ __forceinline__ void flush_last(const uint32_t* s_ring, uint8_t* out, uint32_t Pbits, uint32_t carry,
                                           uint32_t slot_units, int32_t* overflow, int lane) {
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
            if (u < slot_units) *reinterpret_cast<uint4*>(out + (size_t)u * 16) = v4;
            else *overflow = 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fast encoder (Lmax <= 2): 2048-symbol tiles, 64 symbols (two 32-symbol halves) per lane
// ---------------------------------------------------------------------------------------------
constexpr int ETILE = 2 * TILE;        // symbols per warp tile of the fast encoder (= 2 decode chunks)
constexpr int EF_NST = 2;              // TMA stages per warp (2 KB each)

struct EncFastSmem {
    static constexpr int RW = 256;                              // staging ring words (4096 bits/tile + slack)
    static constexpr int IN = 0;                                // EF_NST * ETILE bytes
    static constexpr int LUT4 = IN + EF_NST * ETILE;            // 256 * 2, 512-byte aligned
    static constexpr int LUT1 = LUT4 + 512;                     // 16 * 4
    static constexpr int RING = LUT1 + 64;                      // RW * 4
    static constexpr int BARS = RING + RW * 4;                  // EF_NST * 8
    static constexpr int PER_WARP = (BARS + EF_NST * 8 + 511) / 512 * 512;
};

// inclusive warp scan: shfl.up with its in-range predicate feeding a predicated add (2 instructions/step)
__device__ __forceinline__ uint32_t warp_incl_scan_p(uint32_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
        asm volatile("{\n\t.reg .u32 t;\n\t.reg .pred p;\n\tshfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n\t@p add.u32 %0, %0, t;\n\t}"
                     : "+r"(v)
                     : "r"(d));
    return v;
}

__device__ __forceinline__ uint32_t lds_u16(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

// 16 symbols (one uint4) -> code bits (right aligned, <= 32) and bit count (16..32): SWAR saturate to
// S-1 (5 ops per 4 symbols), gather the four 2-bit symbols of a word into an 8-bit index with one
// multiply, one 16-bit LUT read per word (code | len << 8), shift/or tree.
__device__ __forceinline__ void encode16(const uint4 q, uint32_t lut4_saddr, uint32_t satk, uint32_t satv, uint32_t& code,
                                         uint32_t& len) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t qc[4], ql[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t lo7 = w[j] & 0x7F7F7F7Fu;
        const uint32_t g = (lo7 + satk) | w[j];                  // bit 7 of a byte: value > S-1
        const uint32_t m = byte_msb_mask(g);                      // 0xFF where bit 7 is set
        const uint32_t ws = (w[j] & ~m) | (satv & m);             // min(byte, S-1): 2-bit symbols
        const uint32_t e = lds_u16((((ws * 0x01041040u) >> 23) & 0x1FEu) | lut4_saddr);
        qc[j] = e & 0xFFu;
        ql[j] = e >> 8;
    }
    qc[0] = (qc[0] << ql[1]) | qc[1]; ql[0] += ql[1];
    qc[2] = (qc[2] << ql[3]) | qc[3]; ql[2] += ql[3];
    code = (qc[0] << ql[2]) | qc[2];
    len = ql[0] + ql[2];
}

// One tile of the fast encoder.  FULLT: the whole tile lies inside the window (every lane emits >= 64 bits).
template <bool FULLT, uint32_t RM>
__device__ __forceinline__ void enc_fast_tile(const uint8_t* tile, uint32_t lut4_saddr, const uint32_t* s_lut1, uint32_t satk,
                                              uint32_t satv, int ts, int start, int end, int lane, uint32_t* s_ring,
                                              uint32_t& Pbits, uint32_t& carry, uint32_t& a_lane) {
    // The lane's 64 bytes are four 16-byte pieces.  With a 64-byte lane stride, reading piece k in
    // every lane would be a 4-way bank conflict; lane l reads piece (k + rot) & 3, rot = (l >> 1) & 3,
    // which is conflict-free, and a two-level select network puts the four results back in order.
    const uint32_t rot = (lane >> 1) & 3;
    uint32_t pc[4], pl[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        encode16(*reinterpret_cast<const uint4*>(tile + 16 * ((k + rot) & 3)), lut4_saddr, satk, satv, pc[k], pl[k]);
    {
        const bool r1 = rot & 1, r2 = rot & 2;
        uint32_t tc[4], tl4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { tc[j] = r1 ? pc[(j + 3) & 3] : pc[j]; tl4[j] = r1 ? pl[(j + 3) & 3] : pl[j]; }
#pragma unroll
        for (int j = 0; j < 4; ++j) { pc[j] = r2 ? tc[(j + 2) & 3] : tc[j]; pl[j] = r2 ? tl4[(j + 2) & 3] : tl4[j]; }
    }
    unsigned long long acc0 = ((unsigned long long)pc[0] << pl[1]) | pc[1];
    unsigned long long acc1 = ((unsigned long long)pc[2] << pl[3]) | pc[3];
    uint32_t nb0 = pl[0] + pl[1], nb1 = pl[2] + pl[3];           // 32..64 bits each
    if (!FULLT) {
        // head/tail tile: a half outside the window emits nothing; a half cut by the window boundary is
        // recoded symbol by symbol (at most two such halves per channel)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int p0 = ts + lane * 64 + h * 32;
            const int vlo = max(start - p0, 0), vhi = min(end - p0, 32);
            if (vlo > 0 || vhi < 32) {
                unsigned long long acc = 0;
                uint32_t nb = 0;
                for (int i = vlo; i < vhi; ++i) {
                    const uint32_t e1 = s_lut1[min((uint32_t)tile[h * 32 + i], 15u)];
                    acc = (acc << (e1 >> 16)) | (e1 & 0xFFFFu);
                    nb += e1 >> 16;
                }
                if (h == 0) { acc0 = acc; nb0 = nb; } else { acc1 = acc; nb1 = nb; }
            }
        }
    }
    // ---- bit offsets ----
    const uint32_t nb = nb0 + nb1;
    const uint32_t incl = warp_incl_scan_p(nb);
    const uint32_t Pnew = Pbits + __shfl_sync(FULL, incl, 31);
    const uint32_t a = Pbits + incl - nb;
    a_lane = a;
    const uint32_t Wi = a >> 5;
    // ---- place the two halves back to back: A at bit a, B at bit a + nb0 ----
    const uint32_t sh0 = a & 31;
    if (!FULLT) { acc0 = nb0 ? acc0 << (64 - nb0) : 0ull; acc1 = nb1 ? acc1 << (64 - nb1) : 0ull; }
    else { acc0 <<= (64 - nb0); acc1 <<= (64 - nb1); }
    const uint32_t Ahi = (uint32_t)(acc0 >> 32), Alo = (uint32_t)acc0;
    uint32_t A0 = Ahi >> sh0;
    const uint32_t A1 = __funnelshift_r(Alo, Ahi, sh0);
    const uint32_t A2 = __funnelshift_r(0u, Alo, sh0);
    const uint32_t e0 = sh0 + nb0;
    const uint32_t nf0 = e0 >> 5;                                 // complete words of A: 0..2 (FULLT: 1..2)
    const uint32_t sh1 = e0 & 31;
    const uint32_t Bhi = (uint32_t)(acc1 >> 32), Blo = (uint32_t)acc1;
    uint32_t B0 = Bhi >> sh1;
    const uint32_t B1 = __funnelshift_r(Blo, Bhi, sh1);
    const uint32_t B2 = __funnelshift_r(0u, Blo, sh1);
    const uint32_t e1 = sh1 + nb1;
    const uint32_t nf1 = e1 >> 5;                                 // complete words of B: 0..2 (FULLT: 1..2)
    if (FULLT) {
        B0 |= sh1 ? (nf0 == 1 ? A1 : A2) : 0u;                    // A's trailing partial word shares B's first word
        const uint32_t tl = (e1 & 31) ? (nf1 == 1 ? B1 : B2) : 0u;
        uint32_t incoming = __shfl_up_sync(FULL, tl, 1);
        if (lane == 0) incoming = carry;
        carry = __shfl_sync(FULL, tl, 31);
        A0 |= incoming;
        s_ring[Wi & RM] = A0;
        if (nf0 == 2) s_ring[(Wi + 1) & RM] = A1;
        s_ring[(Wi + nf0) & RM] = B0;
        if (nf1 == 2) s_ring[(Wi + nf0 + 1) & RM] = B1;
    } else {
        B0 |= sh1 ? (nf0 == 0 ? A0 : (nf0 == 1 ? A1 : A2)) : 0u;
        const uint32_t tl = (e1 & 31) ? (nf1 == 0 ? B0 : (nf1 == 1 ? B1 : B2)) : 0u;
        const uint32_t incoming = tails_segmented(tl, (nf0 + nf1) > 0, carry, lane);
        if (nf0 >= 1) A0 |= incoming; else B0 |= incoming;        // the word at Wi
        if (nf0 >= 1) s_ring[Wi & RM] = A0;
        if (nf0 == 2) s_ring[(Wi + 1) & RM] = A1;
        if (nf1 >= 1) s_ring[(Wi + nf0) & RM] = B0;
        if (nf1 == 2) s_ring[(Wi + nf0 + 1) & RM] = B1;
    }
    Pbits = Pnew;
}

__global__ void __launch_bounds__(ENC_WARPS * 32, 4) k_encode_fast(const __grid_constant__ EncParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    using SM = EncFastSmem;
    constexpr uint32_t RM = SM::RW - 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sm = smem_raw + warp * SM::PER_WARP;
    uint8_t* s_in = sm + SM::IN;
    const uint32_t lut4_saddr = smem_u32(sm + SM::LUT4);          // 512-byte aligned: OR-able with idx*2
    uint32_t* s_lut1 = reinterpret_cast<uint32_t*>(sm + SM::LUT1);
    uint32_t* s_ring = reinterpret_cast<uint32_t*>(sm + SM::RING);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(sm + SM::BARS);

    const TabHdr* T = reinterpret_cast<const TabHdr*>(P.tab);
    const int K = T->K, S = T->S;
    if (S != P.S || K != P.K || T->Lmax != P.Lmax || T->Lmax > 2 || T->enc4_off == 0 || (lut4_saddr & 511u)) {
        if (threadIdx.x == 0) *P.overflow = 2;   // launch configuration does not match the table block
        return;
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < EF_NST; ++i) mbar_init(&s_bar[i], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const uint32_t* g_enc1 = reinterpret_cast<const uint32_t*>(P.tab + T->enc1_off);
    const uint4* g_enc4 = reinterpret_cast<const uint4*>(P.tab + T->enc4_off);
    const uint32_t satk = (uint32_t)(0x7F - (S - 1)) * 0x01010101u;   // SWAR saturation constants
    const uint32_t satv = (uint32_t)(S - 1) * 0x01010101u;
    const uint32_t slot_units = (uint32_t)min((long long)(P.slot_bytes >> 4), 0x7FFFFFFFll);

    const int gwarp = blockIdx.x * ENC_WARPS + warp, nwarps = gridDim.x * ENC_WARPS;
    uint32_t slot = 0, parity = 0;   // ring position of the next tile to consume
    int cur_combo = -1;

    for (int c = gwarp; c < P.L.C; c += nwarps) {
        const int n = ch_len(P.L, c);
        const int start = P.start[c];
        const int end = min(P.end[c], n);
        uint32_t Pbits = 0;
        if (end > start && start >= 0) {
            const int combo = (int)P.peak[c] * K + (int)P.enc[c];
            if (combo != cur_combo) {   // this (peak, codebook row) pair's LUTs: 512 B + 64 B
                __syncwarp();
                reinterpret_cast<uint4*>(sm + SM::LUT4)[lane] = g_enc4[(size_t)combo * 32 + lane];
                if (lane < 16) s_lut1[lane] = g_enc1[(size_t)combo * 16 + lane];
                cur_combo = combo;
                __syncwarp();
            }
            const uint8_t* row = P.L.sym + ch_off(P.L, c);
            const int A0 = start & ~(ETILE - 1);
            const int nt = (end - A0 + ETILE - 1) / ETILE;
            const int rd_end = (end + 15) & ~15;
            // chunk (1024-symbol) side info: tile t holds chunks 2t+dj and 2t+dj+1, numbered from start/1024
            uint32_t* co = P.chunk_off + (size_t)c * P.chunk_stride + (A0 / TILE - start / TILE);
            uint8_t* out = P.stream + (size_t)c * P.slot_bytes;
            uint32_t carry = 0;

            if (lane == 0) {   // prologue: fill the ring
                uint32_t s2 = slot;
                const int npro = nt < EF_NST ? nt : EF_NST;
                for (int t = 0; t < npro; ++t) {
                    const int ts = A0 + t * ETILE;
                    const uint32_t bytes = (uint32_t)min(ETILE, rd_end - ts);
                    mbar_expect_tx(&s_bar[s2], bytes);
                    tma_load_1d(s_in + s2 * ETILE, row + ts, bytes, &s_bar[s2]);
                    s2 = (s2 + 1) & (EF_NST - 1);
                }
            }

            int ts = A0;
            for (int t = 0; t < nt; ++t, ts += ETILE) {
                mbar_wait(&s_bar[slot], parity);
                const uint8_t* tile = s_in + slot * ETILE + lane * 64;
                const uint32_t Pold = Pbits;
                uint32_t a_lane;
                const bool full = (ts >= start) && (ts + ETILE <= end);       // warp-uniform
                if (full) enc_fast_tile<true, RM>(tile, lut4_saddr, s_lut1, satk, satv, ts, start, end, lane, s_ring, Pbits, carry, a_lane);
                else enc_fast_tile<false, RM>(tile, lut4_saddr, s_lut1, satk, satv, ts, start, end, lane, s_ring, Pbits, carry, a_lane);
                // chunk offsets: lane 0 starts the tile's first chunk, lane 16 its second
                if ((lane & 15) == 0) {
                    const int cs = ts + (lane >> 4) * TILE;                   // absolute start of that chunk
                    if (cs + TILE > start && cs < end) co[2 * t + (lane >> 4)] = a_lane;
                }
                // ---- flush complete 128-bit units; refill the TMA slot ----
                __syncwarp();
                for (uint32_t b = Pold; (b >> 7) < (Pbits >> 7); b += 32 * 128)
                    flush_units<RM>(s_ring, out, b, Pbits, slot_units, P.overflow, lane);
                __syncwarp();
                if (lane == 0 && t + EF_NST < nt) {
                    const int ts2 = ts + EF_NST * ETILE;
                    const uint32_t bytes = (uint32_t)min(ETILE, rd_end - ts2);
                    mbar_expect_tx(&s_bar[slot], bytes);
                    tma_load_1d(s_in + slot * ETILE, row + ts2, bytes, &s_bar[slot]);
                }
                slot = (slot + 1) & (EF_NST - 1);
                parity ^= (slot == 0);
            }
            flush_last<RM>(s_ring, out, Pbits, carry, slot_units, P.overflow, lane);
            __syncwarp();
        }
        if (lane == 0) P.total_bits[c] = Pbits;
    }
}

// ---------------------------------------------------------------------------------------------
// general encoder (any Lmax <= 9)
// ---------------------------------------------------------------------------------------------
struct EncGenSmem {
    static constexpr int RW = 512;                              // 1024 symbols * 9 bits = 288 words + slack
    static constexpr int IN = 0;
    static constexpr int LUT2 = IN + ENC_NST * TILE;            // 256 * 4
    static constexpr int LUT1 = LUT2 + 1024;                    // 16 * 4
    static constexpr int RING = LUT1 + 64;
    static constexpr int BARS = RING + RW * 4;
    static constexpr int PER_WARP = (BARS + ENC_NST * 8 + 127) / 128 * 128;
};

__global__ void __launch_bounds__(ENC_WARPS * 32, 3) k_encode_gen(const __grid_constant__ EncParams P) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    using SM = EncGenSmem;
    constexpr uint32_t RM = SM::RW - 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sm = smem_raw + warp * SM::PER_WARP;
    uint8_t* s_in = sm + SM::IN;
    uint8_t* s_lut2 = sm + SM::LUT2;
    uint32_t* s_lut1 = reinterpret_cast<uint32_t*>(sm + SM::LUT1);
    uint32_t* s_ring = reinterpret_cast<uint32_t*>(sm + SM::RING);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(sm + SM::BARS);

    const TabHdr* T = reinterpret_cast<const TabHdr*>(P.tab);
    const int K = T->K;
    if (T->S != P.S || K != P.K || T->Lmax != P.Lmax || TILE * T->Lmax > (SM::RW - 8) * 32) {
        if (threadIdx.x == 0) *P.overflow = 2;
        return;
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < ENC_NST; ++i) mbar_init(&s_bar[i], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const uint32_t* g_enc1 = reinterpret_cast<const uint32_t*>(P.tab + T->enc1_off);
    const uint4* g_enc2 = reinterpret_cast<const uint4*>(P.tab + T->enc2_off);
    const uint32_t slot_units = (uint32_t)min((long long)(P.slot_bytes >> 4), 0x7FFFFFFFll);

    const int gwarp = blockIdx.x * ENC_WARPS + warp, nwarps = gridDim.x * ENC_WARPS;
    uint32_t slot = 0, parity = 0;
    int cur_combo = -1;

    for (int c = gwarp; c < P.L.C; c += nwarps) {
        const int n = ch_len(P.L, c);
        const int start = P.start[c];
        const int end = min(P.end[c], n);
        uint32_t Pbits = 0;
        if (end > start && start >= 0) {
            const int combo = (int)P.peak[c] * K + (int)P.enc[c];
            if (combo != cur_combo) {
                __syncwarp();
                const uint4* src = g_enc2 + (size_t)combo * 64;
#pragma unroll
                for (int i = 0; i < 2; ++i) reinterpret_cast<uint4*>(s_lut2)[lane + 32 * i] = src[lane + 32 * i];
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
                uint32_t s2 = slot;
                const int npro = nt < ENC_NST ? nt : ENC_NST;
                for (int t = 0; t < npro; ++t) {
                    const int ts = A0 + t * TILE;
                    const uint32_t bytes = (uint32_t)min(TILE, rd_end - ts);
                    mbar_expect_tx(&s_bar[s2], bytes);
                    tma_load_1d(s_in + s2 * TILE, row + ts, bytes, &s_bar[s2]);
                    s2 = (s2 + 1) & (ENC_NST - 1);
                }
            }

            int ts = A0;
            for (int t = 0; t < nt; ++t, ts += TILE) {
                mbar_wait(&s_bar[slot], parity);
                const uint8_t* tile = s_in + slot * TILE + lane * 32;
                const uint4 q0 = *reinterpret_cast<const uint4*>(tile);
                const uint4 q1 = *reinterpret_cast<const uint4*>(tile + 16);
                uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                if (lane == 0) co[t] = Pbits;
                const bool full = (ts >= start) && (ts + TILE <= end);
                const int p0 = ts + lane * 32;
                const int vlo = max(start - p0, 0), vhi = min(end - p0, 32);   // valid symbols [vlo, vhi)

                // pass 1: look every symbol pair up once (code | len << 24) and count the bits this lane will emit
                uint32_t nb = 0;
                uint32_t pe[16];
                if (full) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        uint32_t wj = w[j];
                        if ((wj & 0xF0F0F0F0u) != 0) {   // bytes >= 16 -> 15 (the LUT saturates the rest)
                            uint32_t h4 = wj & 0xF0F0F0F0u, t1 = h4 | (h4 >> 1);
                            t1 |= t1 >> 2;
                            wj = (wj | (((t1 >> 4) & 0x01010101u) * 0xFFu)) & 0x0F0F0F0Fu;
                        }
                        const uint32_t y = ((wj << 2) | (wj >> 2)) & 0x03FC03FCu;   // two byte offsets of 4-byte entries
                        pe[2 * j] = *reinterpret_cast<const uint32_t*>(s_lut2 + (y & 0xFFFFu));
                        pe[2 * j + 1] = *reinterpret_cast<const uint32_t*>(s_lut2 + (y >> 16));
                        nb += (pe[2 * j] >> 24) + (pe[2 * j + 1] >> 24);
                    }
                } else {
                    for (int i = vlo; i < vhi; ++i) nb += s_lut1[min((uint32_t)tile[i], 15u)] >> 16;
                }
                const uint32_t incl = warp_incl_scan_p(nb);
                const uint32_t Pnew = Pbits + __shfl_sync(FULL, incl, 31);
                const uint32_t a = Pbits + incl - nb;
                const uint32_t Wi = a >> 5;

                // pass 2: sequential bit writer; the first completed word waits for the incoming partial word
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
                    for (int j = 0; j < 16; ++j) append(pe[j] & 0xFFFFFFu, (int)(pe[j] >> 24));
                } else {
                    for (int i = vlo; i < vhi; ++i) {
                        const uint32_t e1 = s_lut1[min((uint32_t)tile[i], 15u)];
                        append(e1 & 0xFFFFu, (int)(e1 >> 16));
                    }
                }
                const uint32_t tl = fill > 0 ? (uint32_t)(pend >> 32) : 0u;
                uint32_t incoming;
                if (full) {   // 32 symbols of >= 1 bit: every lane completes a word, the partial word comes from the previous lane
                    incoming = __shfl_up_sync(FULL, tl, 1);
                    if (lane == 0) incoming = carry;
                    carry = __shfl_sync(FULL, tl, 31);
                } else {
                    incoming = tails_segmented(tl, nemit > 0, carry, lane);
                }
                if (nemit > 0) s_ring[Wi & RM] = first | incoming;

                __syncwarp();
                for (uint32_t b = Pbits; (b >> 7) < (Pnew >> 7); b += 32 * 128)   // up to 72 units per tile
                    flush_units<RM>(s_ring, out, b, Pnew, slot_units, P.overflow, lane);
                Pbits = Pnew;
                __syncwarp();
                if (lane == 0 && t + ENC_NST < nt) {
                    const int ts2 = ts + ENC_NST * TILE;
                    const uint32_t bytes = (uint32_t)min(TILE, rd_end - ts2);
                    mbar_expect_tx(&s_bar[slot], bytes);
                    tma_load_1d(s_in + slot * TILE, row + ts2, bytes, &s_bar[slot]);
                }
                slot = (slot + 1) & (ENC_NST - 1);
                parity ^= (slot == 0);
            }
            flush_last<RM>(s_ring, out, Pbits, carry, slot_units, P.overflow, lane);
            __syncwarp();
        }
        if (lane == 0) P.total_bits[c] = Pbits;
    }
}

}  // namespace mua
