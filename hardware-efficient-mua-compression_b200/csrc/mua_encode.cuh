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
//   k_encode_pair : codebooks with Lmax <= 8 (S <= 9): pair LUT in base S+1 with a null digit for symbols outside
//                   the window, 8-symbol pieces OR-ed into a zeroed ring with shared-memory atomics.
//   k_encode_gen  : any codebook (Lmax <= 9): nibble-pair LUT + sequential per-lane bit writer.
#pragma once
#include <cuda.h>

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
    uint32_t* sub_off;    // NULL, or uint32 [C][sub_stride]: bit offset of every 128-symbol sub-chunk (fast encoder)
    int32_t sub_stride;
    int64_t* total_bits;
    int32_t* overflow;
    // multi-GPU report sink (mua_report_sink): row (row0 + c) of every peer's int32 [C_total][4] report buffer
    int32_t n_peers;
    int64_t row0;
    int32_t* rep[MUA_MAX_PEERS];
    // signal_step > 0: the last block to retire tells every peer that this rank's rows of that step are written
    int32_t signal_step, rank;
    int32_t* flags[MUA_MAX_PEERS];
};

// Channel epilogue of every encoder: the bit count for the local caller and -- when a report sink is attached -- the channel's
// report row {bits, window symbols, SCLV row, peak} (get_BR_no_sort.py:282-287: what the BR report needs) stored straight
// into every peer's report buffer, one 16-byte store per peer over NVLink (lane p serves peer p): the "gather" of the
// multi-GPU step costs no kernel and no collective.
__device__ __forceinline__ void publish_channel(const EncParams& P, int c, uint32_t Pbits, int start, int end, int pk, int en, int lane) {
    if (lane == 0) P.total_bits[c] = Pbits;
    if (lane < P.n_peers) {
        const int4 row = make_int4((int)Pbits, max(end - start, 0), en, pk);
        *reinterpret_cast<int4*>(P.rep[lane] + 4 * (P.row0 + c)) = row;
    }
}

// End of an encoder block when the sink asks for the signal: every warp makes its peer stores visible system-wide, the block's
// last thread through the barrier counts the block in; the block that completes the count publishes the step to every peer
// (what k_report_signal does as a separate launch) and re-arms the counter.
__device__ __forceinline__ void signal_when_last(const EncParams& P) {
    if (P.signal_step <= 0) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < 32) {
        int32_t* counter = P.flags[P.rank] + MUA_MAX_PEERS + 1;
        int last = 0;
        if (threadIdx.x == 0) {
            __threadfence_system();
            last = atomicAdd(counter, 1) == (int)gridDim.x - 1;
        }
        last = __shfl_sync(FULL, last, 0);
        if (last) {
            __threadfence_system();
            if (threadIdx.x == 0) *counter = 0;
            if ((int)threadIdx.x < P.n_peers)
                asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(P.flags[threadIdx.x] + P.rank), "r"(P.signal_step) : "memory");
        }
    }
}

// k_encode_fast with the recording described to the TMA engine (see there)
struct EncFastParams {
    EncParams E;
    alignas(64) CUtensorMap tmap;   // TENSOR: the recording's bytes as uint8 [bytes / 64][64], box 64 x 32 (one 2048-symbol tile), 64-byte swizzle
};

__device__ __forceinline__ void tma_load_2d_p(void* dst_smem, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar))
                 : "memory");
}

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
            *overflow = MUA_ENC_OVERFLOW;
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
            else *overflow = MUA_ENC_OVERFLOW;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fast encoder (Lmax <= 2, S <= 3): 64 symbols per lane and 2048-symbol tile
// ---------------------------------------------------------------------------------------------
// Tiles start at the window (origin = start rounded down to 64 symbols, one lane's share), not at absolute multiples of the
// tile size: every tile but the last is a full tile (and the first one when the window starts off a 64-symbol boundary).
// The last tile is coded with as few 16-symbol pieces per lane as cover what is left of the window (1..4, warp-uniform):
// a 1 200-symbol window costs 3 pieces per lane, not a whole 2048-symbol tile.  Partial tiles use a second table whose
// base-(S+1) index has a null digit without bits for symbols outside the window (SWAR range mask), so they run the code
// of a full tile -- no per-symbol loops.  The 1024-symbol chunks of the side info stay aligned to ABSOLUTE bin indices (the
// stream format does not change): in a full tile a chunk starts with the first symbol of a lane; in a partial tile it falls on
// a 16-symbol piece boundary inside some lane, whose bit offset is the lane's plus the lengths of its pieces before it.
static_assert(TILE == 1024, "chunk arithmetic below assumes 1024-symbol chunks");
static_assert(TILE == 1024, "chunk arithmetic below assumes 1024-symbol chunks");
constexpr int EF_NG = 2;               // 32-symbol groups per lane
constexpr int ETILE = EF_NG * TILE;    // symbols per warp tile of the fast encoder (two decode chunks)
constexpr int EF_NST = 2;              // TMA stages per warp
constexpr int EF_WARPS = 8;            // warps per CTA
constexpr int EF_LUT_B = 768;          // per (peak, row) pair: base-S codes [0,128) + lengths [128,256), null-digit codes [256,512) + lengths [512,768)

struct EncFastSmem {
    static constexpr int RW = 256;                              // staging ring words (2*ETILE bits per tile + slack)
    static constexpr int IN = 0;                                // EF_NST * ETILE bytes
    static constexpr int LUT = IN + EF_NST * ETILE;             // EF_LUT_B bytes, 256-byte aligned
    static constexpr int RING = LUT + EF_LUT_B;                 // RW * 4
    static constexpr int BARS = RING + RW * 4;                  // EF_NST * 8
    static constexpr int PER_WARP = (BARS + EF_NST * 8 + 255) / 256 * 256;
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

__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
template <int OFF>
__device__ __forceinline__ uint32_t lds_u8_at(uint32_t saddr) {   // same index, second table (+OFF bytes)
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1+%2];" : "=r"(v) : "r"(saddr), "n"(OFF));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8_256(uint32_t saddr) { return lds_u8_at<256>(saddr); }

// 16 symbols (one uint4) -> code bits (right aligned, <= 32) and bit count (16..32).  The kernel is bound by the
// ALU pipe, so the per-word work is cut to 3 ALU ops (+2 on the FMA pipe, 2 byte loads):
//   g   = w + satk            (FMA pipe)  bit 7 of a byte set  <=>  byte > S-1   (bytes < 129: checked per tile)
//   m   = prmt.msb(g)         (ALU)       0xFF in those bytes
//   ws  = lop3(w, m, satv)    (ALU)       min(byte, S-1): four 2-bit symbols
//   p   = ws * gmul           (FMA pipe)  gathers them into the top byte in base S: q0 + S*q1 + S^2*q2 + S^3*q3
//                                         (gmul = S^3 | S^2<<8 | S<<16 | 1<<24; no carries: every partial sum < 256)
//   adr = prmt(p, lut)        (ALU)       LUT base (256-byte aligned) with its low byte replaced by the index
//   code = lut[adr], len = lut[adr + 128] (two byte loads: no unpacking; S^4 <= 81 entries: one word per bank)
template <int SV>
__device__ __forceinline__ void encode16(const uint4 q, uint32_t lut_saddr, uint32_t& code, uint32_t& len) {
    constexpr uint32_t satk = (uint32_t)(0x7F - (SV - 1)) * 0x01010101u;   // SWAR saturation constants
    constexpr uint32_t satv = (uint32_t)(SV - 1) * 0x01010101u;
    constexpr uint32_t gmul = (uint32_t)(SV * SV * SV) | ((uint32_t)(SV * SV) << 8) | ((uint32_t)SV << 16) | (1u << 24);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t qc[4], ql[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t g = w[j] + satk;
        const uint32_t m = byte_msb_mask(g);
        const uint32_t ws = (w[j] & ~m) | (satv & m);
        const uint32_t adr = __byte_perm(ws * gmul, lut_saddr, 0x7653);
        qc[j] = lds_u8(adr);
        ql[j] = lds_u8_at<128>(adr);
    }
    qc[0] = (qc[0] << ql[1]) | qc[1]; ql[0] += ql[1];
    qc[2] = (qc[2] << ql[3]) | qc[3]; ql[2] += ql[3];
    code = (qc[0] << ql[2]) | qc[2];
    len = ql[0] + ql[2];
}

// Same for a piece of a partial tile: symbols whose index in the lane (i0 .. i0+15) lies outside [vlo, vhi) become the
// null digit S of the base-(S+1) table at lutn_saddr (= the pair's table block + 256) and code no bits: 0..32 bits.
template <int SV>
__device__ __forceinline__ void encode16n(const uint4 q, uint32_t lutn_saddr, uint32_t i0, uint32_t vlo4, uint32_t vhi4, uint32_t& code,
                                          uint32_t& len) {
    constexpr uint32_t satk = (uint32_t)(0x7F - (SV - 1)) * 0x01010101u;
    constexpr uint32_t satv = (uint32_t)(SV - 1) * 0x01010101u;
    constexpr uint32_t nullv = (uint32_t)SV * 0x01010101u;
    constexpr int B = SV + 1;
    constexpr uint32_t gmul = (uint32_t)(B * B * B) | ((uint32_t)(B * B) << 8) | ((uint32_t)B << 16) | (1u << 24);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t qc[4], ql[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t g = w[j] + satk;
        const uint32_t m = byte_msb_mask(g);
        uint32_t ws = (w[j] & ~m) | (satv & m);
        const uint32_t iv = (0x83828180u + 0x04040404u * (uint32_t)j) + i0 * 0x01010101u;   // symbol indices of this word, bit 7 set (< 64)
        const uint32_t ok = (iv - vlo4) & ~(iv - vhi4);                                      // bit 7: vlo <= index < vhi
        const uint32_t vm = byte_msb_mask(ok);
        ws = (ws & vm) | (nullv & ~vm);
        const uint32_t adr = __byte_perm(ws * gmul, lutn_saddr, 0x7653);
        qc[j] = lds_u8(adr);
        ql[j] = lds_u8_at<256>(adr);
    }
    qc[0] = (qc[0] << ql[1]) | qc[1]; ql[0] += ql[1];
    qc[2] = (qc[2] << ql[3]) | qc[3]; ql[2] += ql[3];
    code = (qc[0] << ql[2]) | qc[2];
    len = ql[0] + ql[2];
}

// bytes >= 128 would carry in `w + satk`: clamp them to 127 first (they saturate to S-1 anyway)
__device__ __forceinline__ uint4 clamp127(uint4 q) {
    uint32_t* w = reinterpret_cast<uint32_t*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = (w[j] | (byte_msb_mask(w[j]) & 0x7F7F7F7Fu)) & 0x7F7F7F7Fu;
    return q;
}

// The four 16-symbol pieces of a lane in a full tile.  The lane's 64 bytes are four 16-byte pieces; with a lane stride
// of 64 bytes, reading piece k in every lane would be a bank conflict.
//   TENSOR: the tile was written by a TMA tensor box with the 64-byte swizzle (piece k of lane l at k ^ ((l >> 1) & 3)): the lane
//           reads its pieces in order, conflict-free, and codes them in order;
//   else  : lane l reads piece (k + rot) mod 4 (rot from the lane index: conflict-free) and a two-level select network puts the
//           results back in order (16 SEL + the rotation arithmetic per tile, on the ALU pipe that bounds the kernel).
template <int SV, bool TENSOR>
__device__ __forceinline__ void enc_fast_pieces(const uint8_t* tile, uint32_t lut_saddr, int lane, uint32_t (&pc)[4], uint32_t (&pl)[4]) {
    constexpr int NP = 4;
    const uint32_t rot = (lane >> 1) & 3;
    uint4 qv[NP];
    uint32_t any_hi = 0;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        qv[k] = TENSOR ? *reinterpret_cast<const uint4*>(tile + 16 * (k ^ rot)) : *reinterpret_cast<const uint4*>(tile + 16 * ((k + rot) & (NP - 1)));
        any_hi |= (qv[k].x | qv[k].y) | (qv[k].z | qv[k].w);
    }
    if (__any_sync(FULL, (any_hi & 0x80808080u) != 0)) {   // rare: a count >= 128 somewhere in the tile
#pragma unroll
        for (int k = 0; k < NP; ++k) qv[k] = clamp127(qv[k]);
    }
#pragma unroll
    for (int k = 0; k < NP; ++k) encode16<SV>(qv[k], lut_saddr, pc[k], pl[k]);
    if (!TENSOR) {
#pragma unroll
        for (int lev = 1; lev < NP; lev <<= 1) {
            const bool r = rot & lev;
            uint32_t tc[NP], tl4[NP];
#pragma unroll
            for (int j = 0; j < NP; ++j) { tc[j] = r ? pc[(j - lev) & (NP - 1)] : pc[j]; tl4[j] = r ? pl[(j - lev) & (NP - 1)] : pl[j]; }
#pragma unroll
            for (int j = 0; j < NP; ++j) { pc[j] = tc[j]; pl[j] = tl4[j]; }
        }
    }
}

// The pieces of a lane in a partial tile: lane l holds symbols [16 k l, 16 k (l + 1)) of the tile, k = pieces per lane (1..4,
// warp-uniform); symbols outside the window code no bits; pieces >= k are empty.
template <int SV, bool TENSOR>
__device__ __forceinline__ void enc_fast_pieces_partial(const uint8_t* tile_w, uint32_t lutn_saddr, int ts, int start, int end, int k, int lane,
                                                        uint32_t (&pc)[4], uint32_t (&pl)[4]) {
    const uint8_t* lp = tile_w + lane * 16 * k;
    const int p0 = ts + lane * 16 * k;                                  // absolute index of the lane's first symbol
    const uint32_t vlo4 = (uint32_t)min(max(start - p0, 0), 64) * 0x01010101u;   // valid symbols of this lane: [vlo, vhi)
    const uint32_t vhi4 = (uint32_t)min(max(end - p0, 0), 64) * 0x01010101u;
    uint4 qv[4];
    uint32_t any_hi = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        qv[i] = make_uint4(0, 0, 0, 0);
        if (i < k) {
            if (TENSOR) {   // swizzled tile: the piece at byte offset o sits in row o / 64 at chunk (o / 16 % 4) ^ (row / 2 % 4)
                const uint32_t o = (uint32_t)(lane * 16 * k + 16 * i);
                qv[i] = *reinterpret_cast<const uint4*>(tile_w + ((o & ~63u) | ((((o >> 4) ^ (o >> 7)) & 3u) << 4)));
            } else {
                qv[i] = *reinterpret_cast<const uint4*>(lp + 16 * i);
            }
            any_hi |= (qv[i].x | qv[i].y) | (qv[i].z | qv[i].w);
        }
    }
    if (__any_sync(FULL, (any_hi & 0x80808080u) != 0)) {
#pragma unroll
        for (int i = 0; i < 4; ++i) qv[i] = clamp127(qv[i]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        pc[i] = 0; pl[i] = 0;
        if (i < k) encode16n<SV>(qv[i], lutn_saddr, 16u * i, vlo4, vhi4, pc[i], pl[i]);
    }
}

// Bit offsets and placement of a tile's pieces (piece lengths in stream order).  FULLT: every 32-symbol group of every lane
// emits 32..64 bits.  Complete words go to the staging ring; the lane's first complete word waits for the partial word of the
// lanes before it (one shuffle in a full tile, a segmented OR-scan when some lanes complete no word).
template <bool FULLT, uint32_t RM>
__device__ __forceinline__ void enc_fast_place(const uint32_t (&pc)[4], const uint32_t (&pl)[4], int lane, uint32_t* s_ring, uint32_t& Pbits,
                                               uint32_t& carry, uint32_t& a_lane) {
    constexpr int NG = 2;
    unsigned long long acc[NG];
    uint32_t nb[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        acc[g] = ((unsigned long long)pc[2 * g] << pl[2 * g + 1]) | pc[2 * g + 1];
        nb[g] = pl[2 * g] + pl[2 * g + 1];                          // full tile: 32..64 bits
    }
    // ---- bit offsets ----
    const uint32_t nbt = nb[0] + nb[1];
    const uint32_t incl = warp_incl_scan_p(nbt);
    const uint32_t Pnew = Pbits + __shfl_sync(FULL, incl, 31);
    const uint32_t a = Pbits + incl - nbt;
    a_lane = a;
    // ---- place the groups back to back; the lane's first complete word waits for the previous lane's partial word ----
    uint32_t pos = a, cur = 0, firstW = 0, firstIdx = a >> 5;
    bool have_first = false;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const uint32_t sh = pos & 31, wi = pos >> 5;
        const unsigned long long A = FULLT ? (acc[g] << (64 - nb[g])) : (nb[g] ? acc[g] << (64 - nb[g]) : 0ull);
        const uint32_t Ahi = (uint32_t)(A >> 32), Alo = (uint32_t)A;
        const uint32_t W0 = (Ahi >> sh) | cur;
        const uint32_t W1 = __funnelshift_r(Alo, Ahi, sh);
        const uint32_t W2 = __funnelshift_r(0u, Alo, sh);
        const uint32_t e = sh + nb[g];
        const uint32_t nf = e >> 5;                               // complete words: 0..2 (FULLT: 1..2)
        if (FULLT) {
            if (g == 0) firstW = W0; else s_ring[wi & RM] = W0;
            if (nf == 2) s_ring[(wi + 1) & RM] = W1;
            cur = (e & 31) ? (nf == 1 ? W1 : W2) : 0u;
        } else {
            if (nf >= 1) {
                if (!have_first) { firstW = W0; firstIdx = wi; have_first = true; }
                else s_ring[wi & RM] = W0;
            }
            if (nf == 2) s_ring[(wi + 1) & RM] = W1;
            cur = (e & 31) ? (nf == 0 ? W0 : (nf == 1 ? W1 : W2)) : 0u;
        }
        pos += nb[g];
    }
    if (FULLT) {
        uint32_t incoming = __shfl_up_sync(FULL, cur, 1);
        if (lane == 0) incoming = carry;
        carry = __shfl_sync(FULL, cur, 31);
        s_ring[firstIdx & RM] = firstW | incoming;
    } else {
        const uint32_t incoming = tails_segmented(cur, have_first, carry, lane);
        if (have_first) s_ring[firstIdx & RM] = firstW | incoming;
    }
    Pbits = Pnew;
}

// TENSOR (fixed row stride, a multiple of 64): a tile is ONE TMA tensor box -- the recording's bytes as a 2-D tensor of 64-byte
// rows, box 64 x 32 = the tile's 2048 bytes, 64-byte swizzle -- instead of a 1-D bulk copy: the hardware swizzle makes the lanes'
// reads of their 64 bytes conflict-free IN ORDER, which removes the rotated reads and the select network that undoes them
// (~23 of the tile's 232 ALU-pipe instructions; the kernel is bound by that pipe).  Bytes past the end of the buffer are zero-filled
// by the engine, bytes past the window are masked by the null digit as before.
template <int SV, bool TENSOR>   // alphabet size as a compile-time constant: the SWAR constants and the gather multipliers are immediates
__global__ void __launch_bounds__(EF_WARPS * 32, 4) k_encode_fast(const __grid_constant__ EncFastParams PF) {
    const EncParams& P = PF.E;
    static_assert(SV == 2 || SV == 3, "the null-digit table needs (S+1)^4 <= 256 entries");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    using SM = EncFastSmem;
    constexpr uint32_t RM = SM::RW - 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sm = smem_raw + warp * SM::PER_WARP;
    uint8_t* s_in = sm + SM::IN;
    const uint32_t lut_saddr = smem_u32(sm + SM::LUT);            // 256-byte aligned: the index replaces its low byte
    uint32_t* s_ring = reinterpret_cast<uint32_t*>(sm + SM::RING);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(sm + SM::BARS);

    const TabHdr* T = reinterpret_cast<const TabHdr*>(P.tab);
    const int K = T->K;
    constexpr int S = SV;
    if (T->S != SV || S != P.S || K != P.K || T->Lmax != P.Lmax || T->Lmax > 2 || T->enc4_off == 0 || (lut_saddr & 255u)) {
        if (threadIdx.x == 0) *P.overflow = MUA_ENC_BAD_TABLE;   // launch configuration does not match the table block
        return;
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < EF_NST; ++i) mbar_init(&s_bar[i], 1);
        fence_barrier_init();
    }
    __syncwarp();
    const uint4* g_enc4 = reinterpret_cast<const uint4*>(P.tab + T->enc4_off);
    const uint32_t slot_units = (uint32_t)min((long long)(P.slot_bytes >> 4), 0x7FFFFFFFll);

    const int gwarp = blockIdx.x * EF_WARPS + warp, nwarps = gridDim.x * EF_WARPS;
    uint32_t slot = 0, parity = 0;   // ring position of the next tile to consume
    int cur_combo = -1;

    for (int c = gwarp; c < P.L.C; c += nwarps) {
        const int n = ch_len(P.L, c);
        const int start = P.start[c];
        const int end = min(P.end[c], n);
        uint32_t Pbits = 0;
        const int pk_c = P.peak[c], en_c = P.enc[c];
        if (pk_c >= P.S || en_c >= K) {                   // not a channel state this table block can code: flag, encode nothing
            if (lane == 0) *P.overflow = MUA_ENC_BAD_TABLE;
        } else if (end > start && start >= 0) {
            const int combo = pk_c * K + en_c;
            if (combo != cur_combo) {   // this (peak, codebook row) pair's tables: 768 B
                __syncwarp();
                const uint4* src = g_enc4 + (size_t)combo * (EF_LUT_B / 16);
                for (int i = lane; i < EF_LUT_B / 16; i += 32) reinterpret_cast<uint4*>(sm + SM::LUT)[i] = src[i];
                cur_combo = combo;
                __syncwarp();
            }
            const int64_t row_off = ch_off(P.L, c);
            const uint8_t* row = P.L.sym + row_off;
            const int A0 = start & ~63;                   // tiles start with the window (rounded down to a lane's 64 symbols)
            const int nt = (end - A0 + ETILE - 1) / ETILE;
            const int rd_end = (end + 15) & ~15;
            // chunk (1024-symbol) side info, numbered from start/1024; chunk 0 starts with the window
            uint32_t* co = P.chunk_off + (size_t)c * P.chunk_stride - start / TILE;
            if (lane == 0) co[start / TILE] = 0;
            // full tiles: the two absolute multiples of 1024 inside a tile are the first symbols of the same two lanes in every
            // tile of the channel (the tile origin is a multiple of 64); every second lane starts a 128-symbol sub-chunk.  The
            // lane's side-info slots are kept as 32-bit BYTE offsets into the arrays (the host checks that they fit; the kernel
            // sits at its register limit and would otherwise rebuild two 64-bit addresses per tile): slot in tile 0, +8 / +64
            // bytes per tile; bit 31 set = this lane writes nothing.
            const int l1 = ((-A0) & (TILE - 1)) >> 6;      // 0..15
            const uint32_t co_b = (lane & 15) == l1
                                      ? 4u * (uint32_t)((size_t)c * P.chunk_stride + ((A0 + 64 * l1 + (lane >> 4) * TILE) >> 10) - start / TILE)
                                      : 0x80000000u;
            // 128-symbol sub-chunk side info (for the sub-chunk decoder): entry 8 j + i = bit offset of sub-chunk i of chunk j
            uint32_t* so = P.sub_off ? P.sub_off + (size_t)c * P.sub_stride - 8 * (start / TILE) : nullptr;
            if (so && lane == 0) so[start >> 7] = 0;                             // the sub-chunk that holds the window start
            const uint32_t so_b = (so && (((A0 >> 6) + lane) & 1) == 0)
                                      ? 4u * (uint32_t)((size_t)c * P.sub_stride + ((A0 + 64 * lane) >> 7) - 8 * (start / TILE))
                                      : 0x80000000u;
            uint8_t* out = P.stream + (size_t)c * P.slot_bytes;
            uint32_t carry = 0;

            if (lane == 0) {   // prologue: fill the ring
                uint32_t s2 = slot;
                const int npro = nt < EF_NST ? nt : EF_NST;
                for (int t = 0; t < npro; ++t) {
                    const int ts = A0 + t * ETILE;
                    if (TENSOR) {
                        mbar_expect_tx(&s_bar[s2], ETILE);
                        tma_load_2d_p(s_in + s2 * ETILE, &PF.tmap, 0, (int)((row_off + ts) >> 6), &s_bar[s2]);
                    } else {
                        const uint32_t bytes = (uint32_t)min(ETILE, rd_end - ts);
                        mbar_expect_tx(&s_bar[s2], bytes);
                        tma_load_1d(s_in + s2 * ETILE, row + ts, bytes, &s_bar[s2]);
                    }
                    s2 = (s2 + 1) & (EF_NST - 1);
                }
            }

            int ts = A0;
            for (int t = 0; t < nt; ++t, ts += ETILE) {
                mbar_wait(&s_bar[slot], parity);
                const uint8_t* tile_w = s_in + slot * ETILE;
                const uint32_t Pold = Pbits;
                uint32_t a_lane, pc[4], pl[4];
                const bool full = (ts >= start) && (ts + ETILE <= end);       // warp-uniform
                if (full) {
                    enc_fast_pieces<SV, TENSOR>(tile_w + lane * 64, lut_saddr, lane, pc, pl);
                    enc_fast_place<true, RM>(pc, pl, lane, s_ring, Pbits, carry, a_lane);
                    if ((int)co_b >= 0) *reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(P.chunk_off) + (co_b + 8u * (uint32_t)t)) = a_lane;
                    if ((int)so_b >= 0) *reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(P.sub_off) + (so_b + 64u * (uint32_t)t)) = a_lane;
                } else {
                    // partial tile: k 16-symbol pieces per lane cover what is left of the window
                    const int k = min(4, (end - ts + 511) >> 9);
                    enc_fast_pieces_partial<SV, TENSOR>(tile_w, lut_saddr + 256, ts, start, end, k, lane, pc, pl);
                    enc_fast_place<false, RM>(pc, pl, lane, s_ring, Pbits, carry, a_lane);
                    // chunk offsets: the (at most two) absolute multiples of 1024 inside the tile and strictly inside the window
                    const uint32_t kinv = k == 4 ? 64u : (k == 3 ? 86u : (k == 2 ? 128u : 256u));   // floor(pi / k) = pi * kinv >> 8 for pi < 128
                    int cs = (ts + TILE - 1) & ~(TILE - 1);
                    const int cs_end = min(end, ts + ETILE);
                    if (cs <= start) cs += TILE;
#pragma unroll
                    for (int b = 0; b < 2; ++b, cs += TILE) {
                        if (cs < cs_end) {
                            const int pi = (cs - ts) >> 4;                    // piece of the tile the chunk starts with
                            const int ol = (int)(((uint32_t)pi * kinv) >> 8), sub = pi - ol * k;
                            if (lane == ol) {
                                uint32_t o = a_lane;
                                if (sub > 0) o += pl[0];
                                if (sub > 1) o += pl[1];
                                if (sub > 2) o += pl[2];
                                co[cs >> 10] = o;
                            }
                        }
                    }
                    if (so) {
                        // sub-chunk starts among this lane's pieces: at most one (8 pieces apart, k <= 4 pieces per lane)
                        uint32_t o = a_lane;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int p = ts + 16 * (k * lane + i);           // absolute bin of the piece's first symbol
                            if (i < k && (p & 127) == 0 && p > start && p < end) so[p >> 7] = o;
                            o += pl[i];
                        }
                    }
                }
                // ---- flush complete 128-bit units; refill the TMA slot ----
                __syncwarp();
                for (uint32_t b = Pold; (b >> 7) < (Pbits >> 7); b += 32 * 128)
                    flush_units<RM>(s_ring, out, b, Pbits, slot_units, P.overflow, lane);
                __syncwarp();
                if (lane == 0 && t + EF_NST < nt) {
                    const int ts2 = ts + EF_NST * ETILE;
                    if (TENSOR) {
                        mbar_expect_tx(&s_bar[slot], ETILE);
                        tma_load_2d_p(s_in + slot * ETILE, &PF.tmap, 0, (int)((row_off + ts2) >> 6), &s_bar[slot]);
                    } else {
                        const uint32_t bytes = (uint32_t)min(ETILE, rd_end - ts2);
                        mbar_expect_tx(&s_bar[slot], bytes);
                        tma_load_1d(s_in + slot * ETILE, row + ts2, bytes, &s_bar[slot]);
                    }
                }
                slot = (slot + 1) & (EF_NST - 1);
                parity ^= (slot == 0);
            }
            flush_last<RM>(s_ring, out, Pbits, carry, slot_units, P.overflow, lane);
            __syncwarp();
        }
        publish_channel(P, c, Pbits, start, end, pk_c, en_c, lane);
    }
    signal_when_last(P);
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
        if (threadIdx.x == 0) *P.overflow = MUA_ENC_BAD_TABLE;
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
        const int pk_c = P.peak[c], en_c = P.enc[c];
        if (pk_c >= P.S || en_c >= K) {                   // not a channel state this table block can code: flag, encode nothing
            if (lane == 0) *P.overflow = MUA_ENC_BAD_TABLE;
        } else if (end > start && start >= 0) {
            const int combo = pk_c * K + en_c;
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
        publish_channel(P, c, Pbits, start, end, pk_c, en_c, lane);
    }
    signal_when_last(P);
}

// ---------------------------------------------------------------------------------------------
// pair encoder (Lmax <= 8: every SCLV table up to S = 9)
// ---------------------------------------------------------------------------------------------
// Warp per channel, 2048-symbol tiles (two chunks; lane = 32 consecutive symbols of each), TMA ring,
// but the per-symbol work follows the fast encoder:
//   * bytes are saturated with the carry-free SWAR sequence (add on the FMA pipe, PRMT sign-replicate mask, LOP3);
//   * one multiply by 2 | 2(S+1) << 8 turns the four symbols of a word into two byte offsets (bytes 1 and 3) of
//     pair-table entries in base S+1, spliced into the table's shared address by one PRMT each; the pair's code
//     (uint16) and length (uint8) are two loads from that one address;
//   * digit S of the base-(S+1) index is a null symbol without bits: symbols outside the window are replaced by
//     it (SWAR range mask), so head and tail tiles run the same code as full tiles;
//   * pairs are merged to 8-symbol pieces (<= 64 bits) in registers; a prefix sum of the lanes' bit counts gives the
//     bit offsets; every piece is OR-ed into a zeroed staging ring with shared-memory atomics -- words that are
//     all zero (the common codeword of the most frequent symbol is '0') are not touched at all; the flush writes
//     complete 128-bit units with coalesced 16-byte stores and re-zeroes what it has read.
#ifndef MUA_EP_DENSE
#define MUA_EP_DENSE 1                   // register placement for tiles whose lanes all code <= 64 bits per chunk (else: atomics)
#endif
constexpr int EP_NG = 2;                 // 1024-symbol chunks per warp tile
constexpr int EP_TILE = EP_NG * TILE;
constexpr int EP_NST = 2;                // TMA stages per warp
struct EncPairSmem {
    static constexpr int RW = 1024;                             // staging ring words: 2048 symbols * 8 bits = 512 words + the open unit
    static constexpr int IN = 0;                                // EP_NST * EP_TILE bytes
    static constexpr int LUTP = IN + EP_NST * EP_TILE;          // 512 B, 256-byte aligned: codes, then lengths
    static constexpr int RING = LUTP + 512;
    static constexpr int BARS = RING + RW * 4;
    static constexpr int PER_WARP = (BARS + EP_NST * 8 + 511) / 512 * 512;
};

__device__ __forceinline__ uint32_t lds_u16(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

// flush the complete 128-bit units in [Pold, Pnew) and zero them in the ring (<= 32 units per call)
template <uint32_t RM>
__device__ __forceinline__ void flush_units_z(uint32_t* s_ring, uint8_t* out, uint32_t Pold, uint32_t Pnew, uint32_t slot_units,
                                              int32_t* overflow, int lane) {
    const uint32_t u = (Pold >> 7) + lane;
    if (u < (Pnew >> 7)) {
        uint4* rp = reinterpret_cast<uint4*>(&s_ring[(u * 4) & RM]);
        uint4 v4 = *rp;
        *rp = make_uint4(0, 0, 0, 0);
        if (u < slot_units) {
            v4.x = bswap32(v4.x); v4.y = bswap32(v4.y); v4.z = bswap32(v4.z); v4.w = bswap32(v4.w);
            *reinterpret_cast<uint4*>(out + (size_t)u * 16) = v4;
        } else {
            *overflow = MUA_ENC_OVERFLOW;
        }
    }
}

// One 2048-symbol tile: lane l codes symbols [32l, 32l+32) of each of the tile's two chunks.  Returns the bit
// count of the first chunk (the second chunk's offset).
// TENSOR: `tile` is the stage base and the tile was written by a TMA tensor box with the 64-byte swizzle (see k_encode_fast): the 16-byte
// piece at byte offset o sits in row o / 64 at chunk (o / 16 % 4) ^ (row / 2 % 4), which makes the lanes' reads at a 32-byte lane stride
// conflict-free (linear layout: two-way conflicts); else `tile` is the lane's first byte of a linear tile.
template <int SV, bool FULLT, uint32_t RM, bool TENSOR>
__device__ __forceinline__ uint32_t enc_pair_tile(const uint8_t* tile, uint32_t lutp_saddr, int ts, int start, int end, int lane,
                                                  uint32_t* s_ring, uint32_t& Pbits) {
    constexpr uint32_t satk = (uint32_t)(0x7F - (SV - 1)) * 0x01010101u;
    constexpr uint32_t satv = (uint32_t)(SV - 1) * 0x01010101u;
    constexpr uint32_t nullv = (uint32_t)SV * 0x01010101u;
    constexpr uint32_t mult = 2u | ((uint32_t)(2 * (SV + 1)) << 8);
    uint4 q[2 * EP_NG];
    uint32_t any_hi = 0;
#pragma unroll
    for (int g = 0; g < EP_NG; ++g) {
        if (TENSOR) {   // row = 16 g + lane / 2 (row / 2 % 4 == lane / 4 % 4), chunks 2 (lane & 1) and 2 (lane & 1) + 1
            const uint8_t* rowp = tile + g * TILE + (lane >> 1) * 64;
            const int sw = (lane >> 2) & 3, c0 = 2 * (lane & 1);
            q[2 * g] = *reinterpret_cast<const uint4*>(rowp + ((c0 ^ sw) << 4));
            q[2 * g + 1] = *reinterpret_cast<const uint4*>(rowp + (((c0 + 1) ^ sw) << 4));
        } else {
            q[2 * g] = *reinterpret_cast<const uint4*>(tile + g * TILE);
            q[2 * g + 1] = *reinterpret_cast<const uint4*>(tile + g * TILE + 16);
        }
        any_hi |= (q[2 * g].x | q[2 * g].y) | (q[2 * g].z | q[2 * g].w) | (q[2 * g + 1].x | q[2 * g + 1].y) | (q[2 * g + 1].z | q[2 * g + 1].w);
    }
    if (__any_sync(FULL, (any_hi & 0x80808080u) != 0)) {   // rare: a count >= 128 somewhere in the tile
#pragma unroll
        for (int i = 0; i < 2 * EP_NG; ++i) q[i] = clamp127(q[i]);
    }
    unsigned long long oc[4 * EP_NG];
    uint32_t ol[4 * EP_NG], nbp = 0;                      // nbp: bit counts of the two chunks, 16 bits each
#pragma unroll
    for (int g = 0; g < EP_NG; ++g) {
        const uint32_t w[8] = {q[2 * g].x, q[2 * g].y, q[2 * g].z, q[2 * g].w, q[2 * g + 1].x, q[2 * g + 1].y, q[2 * g + 1].z, q[2 * g + 1].w};
        uint32_t vlo4 = 0, vhi4 = 0;
        if (!FULLT) {
            const int p0 = ts + g * TILE + lane * 32;
            vlo4 = (uint32_t)min(max(start - p0, 0), 32) * 0x01010101u;   // valid symbols of this lane: [vlo, vhi)
            vhi4 = (uint32_t)min(max(end - p0, 0), 32) * 0x01010101u;
        }
        uint32_t qc[8], ql[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t gg = w[j] + satk;
            const uint32_t m = byte_msb_mask(gg);
            uint32_t ws = (w[j] & ~m) | (satv & m);
            if (!FULLT) {
                const uint32_t iv = (0x03020100u + 0x04040404u * (uint32_t)j) | 0x80808080u;   // symbol indices of this word, bit 7 set
                const uint32_t ok = (iv - vlo4) & ~(iv - vhi4);                                  // bit 7: vlo <= index < vhi
                const uint32_t vm = byte_msb_mask(ok);
                ws = (ws & vm) | (nullv & ~vm);
            }
            const uint32_t prod = ws * mult;
            const uint32_t a0 = __byte_perm(prod, lutp_saddr, 0x7651), a1 = __byte_perm(prod, lutp_saddr, 0x7653);
            const uint32_t c0 = lds_u16(a0), l0 = lds_u8_256(a0), c1 = lds_u16(a1), l1 = lds_u8_256(a1);
            qc[j] = (c0 << l1) | c1;
            ql[j] = l0 + l1;
        }
        uint32_t nb = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            oc[4 * g + j] = ((unsigned long long)qc[2 * j] << ql[2 * j + 1]) | qc[2 * j + 1];
            ol[4 * g + j] = ql[2 * j] + ql[2 * j + 1];
            nb += ol[4 * g + j];
        }
        nbp |= nb << (16 * g);
    }
    const uint32_t incl = warp_incl_scan_p(nbp);          // both chunks at once: every sum < 2^16
    const uint32_t tot = __shfl_sync(FULL, incl, 31);
    const uint32_t excl = incl - nbp;
    const uint32_t tot0 = tot & 0xFFFFu;
    uint32_t pos[EP_NG] = {Pbits + (excl & 0xFFFFu), Pbits + tot0 + (excl >> 16)};
#if MUA_EP_DENSE
    // Register placement (the fast encoder's): in a full tile every symbol codes >= 1 bit, so a lane's 32 symbols of a chunk
    // are >= 32 bits and complete at least one word; when they are also <= 64 bits in EVERY lane (warp vote -- true for nearly
    // every tile of MUA counts, whose frequent symbols code in 1..2 bits) the four pieces of a chunk are merged into one 64-bit
    // register, shifted to their bit offset with funnel shifts and stored as whole words: the partial word between two lanes
    // travels by one shuffle, between the two chunks of the tile by another.  No atomics, no data-dependent branches.  The ring
    // keeps the protocol of the atomic path (zero ahead of the stream, the open word lives in the ring), so the two paths mix.
    if (FULLT && __all_sync(FULL, ((nbp & 0xFFFFu) <= 64u) & ((nbp >> 16) <= 64u))) {
        uint32_t incoming = (Pbits & 31u) ? s_ring[(Pbits >> 5) & RM] : 0u;     // the stream's open word (broadcast read)
#pragma unroll
        for (int g = 0; g < EP_NG; ++g) {
            unsigned long long acc = oc[4 * g];
#pragma unroll
            for (int j = 1; j < 4; ++j) acc = (acc << ol[4 * g + j]) | oc[4 * g + j];
            const uint32_t nb = (nbp >> (16 * g)) & 0xFFFFu;                     // 32..64
            const uint32_t sh = pos[g] & 31u, wi = pos[g] >> 5;
            const unsigned long long A = acc << (64u - nb);
            const uint32_t Ahi = (uint32_t)(A >> 32), Alo = (uint32_t)A;
            const uint32_t W0 = Ahi >> sh;
            const uint32_t W1 = __funnelshift_r(Alo, Ahi, sh);
            const uint32_t W2 = __funnelshift_r(0u, Alo, sh);
            const uint32_t e = sh + nb;                                          // 32..95: one or two complete words
            const bool two = e >= 64u;
            const uint32_t tail = (e & 31u) ? (two ? W2 : W1) : 0u;
            uint32_t inc = __shfl_up_sync(FULL, tail, 1);
            if (lane == 0) inc = incoming;
            incoming = __shfl_sync(FULL, tail, 31);
            s_ring[wi & RM] = W0 | inc;
            if (two) s_ring[(wi + 1) & RM] = W1;
        }
        Pbits += tot0 + (tot >> 16);
        if (lane == 0 && (Pbits & 31u)) s_ring[(Pbits >> 5) & RM] = incoming;    // the new open word
        return tot0;
    }
#endif
#pragma unroll
    for (int i = 0; i < 4 * EP_NG; ++i) {
        // zero bits need no write (the ring is zero); an empty piece has oc == 0, so its shift amount does not matter
        const unsigned long long A = oc[i] << ((64 - ol[i]) & 63);
        const uint32_t Ahi = (uint32_t)(A >> 32), Alo = (uint32_t)A;
        const uint32_t p = pos[i >> 2];
        const uint32_t sh = p & 31, wi = p >> 5;
        const uint32_t W0 = Ahi >> sh;
        const uint32_t W1 = __funnelshift_r(Alo, Ahi, sh);
        const uint32_t W2 = __funnelshift_r(0u, Alo, sh);
        if (W0) atomicOr(&s_ring[wi & RM], W0);
        if (W1) atomicOr(&s_ring[(wi + 1) & RM], W1);
        if (W2) atomicOr(&s_ring[(wi + 2) & RM], W2);
        pos[i >> 2] = p + ol[i];
    }
    Pbits += tot0 + (tot >> 16);
    return tot0;
}

template <int SV, bool TENSOR>   // TENSOR: tiles as TMA tensor boxes with the 64-byte swizzle (row stride a multiple of 64), as in k_encode_fast
__global__ void __launch_bounds__(ENC_WARPS * 32, 3) k_encode_pair(const __grid_constant__ EncFastParams PF) {
    const EncParams& P = PF.E;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    using SM = EncPairSmem;
    constexpr uint32_t RM = SM::RW - 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sm = smem_raw + warp * SM::PER_WARP;
    uint8_t* s_in = sm + SM::IN;
    const uint32_t lutp_saddr = smem_u32(sm + SM::LUTP);
    uint32_t* s_ring = reinterpret_cast<uint32_t*>(sm + SM::RING);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(sm + SM::BARS);

    const TabHdr* T = reinterpret_cast<const TabHdr*>(P.tab);
    const int K = T->K;
    if (T->S != SV || SV != P.S || K != P.K || T->Lmax != P.Lmax || T->Lmax > 8 || T->encp_off == 0 || (lutp_saddr & 255u)) {
        if (threadIdx.x == 0) *P.overflow = MUA_ENC_BAD_TABLE;   // launch configuration does not match the table block
        return;
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < EP_NST; ++i) mbar_init(&s_bar[i], 1);
        fence_barrier_init();
    }
    for (int i = lane; i < SM::RW; i += 32) s_ring[i] = 0;
    __syncwarp();
    const uint4* g_encp = reinterpret_cast<const uint4*>(P.tab + T->encp_off);
    const uint32_t slot_units = (uint32_t)min((long long)(P.slot_bytes >> 4), 0x7FFFFFFFll);

    const int gwarp = blockIdx.x * ENC_WARPS + warp, nwarps = gridDim.x * ENC_WARPS;
    uint32_t slot = 0, parity = 0;
    int cur_combo = -1;

    for (int c = gwarp; c < P.L.C; c += nwarps) {
        const int n = ch_len(P.L, c);
        const int start = P.start[c];
        const int end = min(P.end[c], n);
        uint32_t Pbits = 0;
        const int pk_c = P.peak[c], en_c = P.enc[c];
        if (pk_c >= P.S || en_c >= K) {                   // not a channel state this table block can code: flag, encode nothing
            if (lane == 0) *P.overflow = MUA_ENC_BAD_TABLE;
        } else if (end > start && start >= 0) {
            const int combo = pk_c * K + en_c;
            if (combo != cur_combo) {   // this (peak, codebook row) pair's table: 512 B
                __syncwarp();
                reinterpret_cast<uint4*>(sm + SM::LUTP)[lane] = g_encp[(size_t)combo * 32 + lane];
                cur_combo = combo;
                __syncwarp();
            }
            const int64_t row_off = ch_off(P.L, c);
            const uint8_t* row = P.L.sym + row_off;
            const int A0 = start & ~(EP_TILE - 1);
            const int nt = (end - A0 + EP_TILE - 1) / EP_TILE;
            const int rd_end = (end + 15) & ~15;
            // chunk (1024-symbol) side info: tile t holds chunks 2t and 2t+1 counted from A0, numbered from start/1024
            uint32_t* co = P.chunk_off + (size_t)c * P.chunk_stride + (A0 / TILE - start / TILE);
            uint8_t* out = P.stream + (size_t)c * P.slot_bytes;

            if (lane == 0) {
                uint32_t s2 = slot;
                const int npro = nt < EP_NST ? nt : EP_NST;
                for (int t = 0; t < npro; ++t) {
                    const int ts = A0 + t * EP_TILE;
                    if (TENSOR) {
                        mbar_expect_tx(&s_bar[s2], EP_TILE);
                        tma_load_2d_p(s_in + s2 * EP_TILE, &PF.tmap, 0, (int)((row_off + ts) >> 6), &s_bar[s2]);
                    } else {
                        const uint32_t bytes = (uint32_t)min(EP_TILE, rd_end - ts);
                        mbar_expect_tx(&s_bar[s2], bytes);
                        tma_load_1d(s_in + s2 * EP_TILE, row + ts, bytes, &s_bar[s2]);
                    }
                    s2 = (s2 + 1) & (EP_NST - 1);
                }
            }

            int ts = A0;
            for (int t = 0; t < nt; ++t, ts += EP_TILE) {
                mbar_wait(&s_bar[slot], parity);
                const uint8_t* tile = s_in + slot * EP_TILE + (TENSOR ? 0 : lane * 32);
                const uint32_t Pold = Pbits;
                const bool full = (ts >= start) && (ts + EP_TILE <= end);    // warp-uniform
                uint32_t tot0;
                if (full) tot0 = enc_pair_tile<SV, true, RM, TENSOR>(tile, lutp_saddr, ts, start, end, lane, s_ring, Pbits);
                else tot0 = enc_pair_tile<SV, false, RM, TENSOR>(tile, lutp_saddr, ts, start, end, lane, s_ring, Pbits);
                if (lane < EP_NG) {
                    const int cs = ts + lane * TILE;                         // absolute start of that chunk
                    if (cs + TILE > start && cs < end) co[EP_NG * t + lane] = Pold + (lane ? tot0 : 0u);
                }
                __syncwarp();
                for (uint32_t b = Pold; (b >> 7) < (Pbits >> 7); b += 32 * 128)   // up to 128 units per tile
                    flush_units_z<RM>(s_ring, out, b, Pbits, slot_units, P.overflow, lane);
                __syncwarp();
                if (lane == 0 && t + EP_NST < nt) {
                    const int ts2 = ts + EP_NST * EP_TILE;
                    if (TENSOR) {
                        mbar_expect_tx(&s_bar[slot], EP_TILE);
                        tma_load_2d_p(s_in + slot * EP_TILE, &PF.tmap, 0, (int)((row_off + ts2) >> 6), &s_bar[slot]);
                    } else {
                        const uint32_t bytes = (uint32_t)min(EP_TILE, rd_end - ts2);
                        mbar_expect_tx(&s_bar[slot], bytes);
                        tma_load_1d(s_in + slot * EP_TILE, row + ts2, bytes, &s_bar[slot]);
                    }
                }
                slot = (slot + 1) & (EP_NST - 1);
                parity ^= (slot == 0);
            }
            // last partial unit (zero padded: the ring holds zeros past the last bit), then leave the ring clean
            if ((Pbits & 127) && lane == 0) {
                const uint32_t u = Pbits >> 7;
                uint4* rp = reinterpret_cast<uint4*>(&s_ring[(u * 4) & RM]);
                uint4 v4 = *rp;
                *rp = make_uint4(0, 0, 0, 0);
                if (u < slot_units) {
                    v4.x = bswap32(v4.x); v4.y = bswap32(v4.y); v4.z = bswap32(v4.z); v4.w = bswap32(v4.w);
                    *reinterpret_cast<uint4*>(out + (size_t)u * 16) = v4;
                } else {
                    *P.overflow = MUA_ENC_OVERFLOW;
                }
            }
            __syncwarp();
        }
        publish_channel(P, c, Pbits, start, end, pk_c, en_c, lane);
    }
    signal_when_last(P);
}


// ---------------------------------------------------------------------------------------------
// report sink flags (mua_report_signal / mua_report_wait)
// ---------------------------------------------------------------------------------------------
struct PeerFlags {
    int32_t n, rank;
    int32_t* flags[MUA_MAX_PEERS];   // int32 [MUA_MAX_PEERS + 1] on every peer
};

// Enqueued after the encoder on the same stream: the encoder's peer stores are complete (kernel boundary); the fence +
// system-scope release store publish them to the peer that polls the flag.
__global__ void k_report_signal(const __grid_constant__ PeerFlags F, int32_t step) {
    const int p = threadIdx.x;
    if (p < F.n) {
        __threadfence_system();
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(F.flags[p] + F.rank), "r"(step) : "memory");
    }
}

__device__ __forceinline__ long long global_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// lane p polls the flag of source rank p in the OWN flag block; gives up after `max_ns` (sticky marker in slot MUA_MAX_PEERS)
__global__ void k_report_wait(const __grid_constant__ PeerFlags F, int32_t step, long long max_ns) {
    const int p = threadIdx.x;
    if (p < F.n) {
        const int32_t* f = F.flags[F.rank] + p;
        const long long t0 = global_ns();
        int32_t v;
        for (;;) {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if (v >= step) break;
            if (global_ns() - t0 > max_ns) {
                F.flags[F.rank][MUA_MAX_PEERS] = 1;
                break;
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// dense packing of the per-channel slots (mua_pack_streams): what a caller ships off the device
// ---------------------------------------------------------------------------------------------
// unit_off[c] = sum_{i<c} ceil(total_bits[i] / 128) (16-byte units), unit_off[C] = total: one CTA, chunked block scan.
__global__ void __launch_bounds__(1024) k_pack_offsets(const int64_t* __restrict__ total_bits, int C, int64_t slot_units,
                                                       int64_t* __restrict__ unit_off) {
    __shared__ long long s_warp[32];
    __shared__ long long s_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < C; c0 += 1024) {
        const int c = c0 + threadIdx.x;
        long long u = 0;
        if (c < C) {
            u = (total_bits[c] + 127) >> 7;
            if (u > slot_units) u = slot_units;              // an overflowed slot holds no more than this
        }
        long long incl = u;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long t = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            long long w = s_warp[lane], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const long long t = __shfl_up_sync(FULL, wi, d);
                if (lane >= d) wi += t;
            }
            s_warp[lane] = wi - w;                           // exclusive
        }
        __syncthreads();
        const long long base = s_base;
        if (c < C) unit_off[c] = base + s_warp[warp] + incl - u;
        __syncthreads();
        if (threadIdx.x == 1023) s_base = base + s_warp[warp] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) unit_off[C] = s_base;
}

// warp per channel: copy the used 16-byte units of the slot to the dense buffer (coalesced both ways)
__global__ void __launch_bounds__(256) k_pack_copy(const uint8_t* __restrict__ stream, int64_t slot_bytes, int C,
                                                   const int64_t* __restrict__ unit_off, uint8_t* __restrict__ dense,
                                                   int64_t dense_units) {
    const int lane = threadIdx.x & 31;
    const int nw = gridDim.x * (blockDim.x >> 5);
    for (int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < C; c += nw) {
        const long long o = unit_off[c], n = unit_off[c + 1] - o;
        const uint4* src = reinterpret_cast<const uint4*>(stream + (size_t)c * slot_bytes);
        uint4* dst = reinterpret_cast<uint4*>(dense) + o;
        for (long long i = lane; i < n; i += 32)
            if (o + i < dense_units) dst[i] = src[i];
    }
}

}  // namespace mua
