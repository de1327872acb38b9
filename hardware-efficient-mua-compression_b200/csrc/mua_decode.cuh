// Stage 6: table-driven chunk-parallel decode (one lane per 1024-symbol chunk, multi-symbol LUT),
// the device-side round-trip check, the synthetic MUA generator and the binning kernels (stage 1).
#pragma once
#include <cuda.h>

#include "mua_common.cuh"

namespace mua {

struct DecParams {
    const uint8_t* stream;
    int64_t slot_bytes;
    const uint32_t* chunk_off;
    int32_t chunk_stride;
    const uint32_t* sub_off;    // NULL, or the encoder's 128-symbol sub-chunk offsets (k_decode_sub)
    int32_t sub_stride;
    int32_t item_chunks;        // chunks per channel that can be non-empty (<= chunk_stride)
    int32_t by_chunk;           // item order, see dec_item
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
    int32_t* status;            // see dec_flag
    // fused report wait (mua_report_sink passed to mua_decode): block 0 ends by polling the own flag block until every rank
    // has signalled wait_step -- the decode that follows an encode then also guarantees the gathered report, with no extra launch
    const int32_t* wait_flags;
    int32_t wait_n, wait_step;
    int32_t var_str_w, var_pps;   // k_decode_var: staged stream words per lane and stage, 128-symbol periods per stage
};

constexpr int DG_OUT_B = 144;         // output tile row of the general decoder: 128 B + 16 B pad

// Work item (one lane's chunk) -> (channel, chunk).  Long rows: item = c * item_chunks + j, a warp's 32 lanes hold consecutive
// chunks of a channel (their stream bytes are neighbours).  Short rows (by_chunk: a few chunks per channel of very different
// lengths -- a 1 200-symbol window that starts at 64 is a 960- and a 240-symbol chunk): blocks of 32 channels, chunk-major inside
// a block, so that the 32 lanes of a group hold the SAME chunk index of 32 channels and finish together instead of idling
// behind the longest chunk (100k x 2 400: 37 % fewer period passes).  Items: ceil(C / 32) * 32 * item_chunks then.
__device__ __forceinline__ long long dec_nitems(const DecParams& P) {
    return P.by_chunk ? (long long)((P.C + 31) / 32) * 32 * P.item_chunks : (long long)P.C * P.item_chunks;
}
__device__ __forceinline__ bool dec_item(const DecParams& P, long long item, long long nitems, int& c, int& j) {
    c = 0; j = 0;
    if (item >= nitems) return false;
    if (P.by_chunk) {
        const long long per = 32ll * P.item_chunks;
        const long long blk = item / per;
        const int r = (int)(item - blk * per);
        j = r >> 5;
        c = (int)(blk * 32) + (r & 31);
        if (c >= P.C) { c = 0; j = 0; return false; }
        return true;
    }
    c = (int)(item / P.item_chunks);
    j = (int)(item - (long long)c * P.item_chunks);
    return true;
}

// Status word of mua_decode (int32 [1], zeroed by the caller): 0 = ok; MUA_DEC_BAD_OFFSET = some chunk's side-info bit
// offset lies past its slot (an encode that overflowed, or corrupt side info): the chunk is skipped, nothing outside the
// stream buffer is read; MUA_DEC_BAD_TABLE = the table block does not match the S/K/Lmax the host passed, or a channel's
// peak >= S / SCLV row >= K: nothing (or not that channel) is decoded.
// little-endian word of the stream -> stream bit j at register bit j
__device__ __forceinline__ uint32_t stream_rev(uint32_t w) { return __byte_perm(__brev(w), 0, 0x0123); }

__device__ __forceinline__ void dec_wait_report(const DecParams& P) {
    if (P.wait_n > 0 && blockIdx.x == 0 && (int)threadIdx.x < P.wait_n) {
        const int32_t* f = P.wait_flags + threadIdx.x;
        long long t0 = 0;
        for (unsigned it = 0;; ++it) {
            int32_t v;
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if (v >= P.wait_step) break;
            long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (it == 0) t0 = t;
            if (t - t0 > 2000000000ll) {                       // 2 s: give up, leave the sticky marker k_report_wait leaves
                const_cast<int32_t*>(P.wait_flags)[MUA_MAX_PEERS] = 1;
                break;
            }
        }
    }
}

__device__ __forceinline__ void dec_flag(int32_t* status, int code) {
    if (status) atomicMax(status, code);
}

// ---- general decoder with variable-count lookups (every codebook the lane / fast decoders do not take) ----
// (Its round-1 predecessor decoded a FIXED number of symbols per lookup -- 1 for S >= 7 -- from per-(peak,row) tables
// that outgrow shared memory: 17 instructions and one global load per symbol; removed.)
// A lookup decodes AS MANY whole symbols as its Wv-bit window holds (at most 4; the SCLV codes of skewed MUA
// counts are 1..2 bits for the frequent symbols, so mostly 4) from rank tables per codebook ROW (K x 2^Wv entries,
// <= 72 KB: always in shared memory); the lane's peak is applied by the PRMT that unpacks an entry (S <= 8: one PRMT
// into the 8-byte rank -> symbol map; S >= 9: two PRMTs into the halves of the 16-byte map and a select).  Decoded
// symbols are appended to a 64-bit register queue and leave for the output tile one 4-byte word at a time, so the
// tile and the write-out are those of the fixed-count decoders.  Same chunk bookkeeping and stream staging (one TMA
// bulk copy per lane and stage, sized for MUA_DV_PPS periods).  One persistent CTA per SM with as many warps as fit beside
// the tables.
// entry (or a bare rank in the low nibble with the other slots set to "none") -> symbols, one per byte: the selectors index the
// lane's rank -> symbol map (TabHdr::idx[peak]); S <= 8: one PRMT into the 8-byte map (selector 0x8 = sign replicate of byte 0 = a
// zero byte); S >= 9: two PRMTs into the halves of the 16-byte map and a select by the selectors' bit 3 (0xF = entry 15 = zero)
template <bool WIDE>
__device__ __forceinline__ uint32_t dv_syms(uint32_t e, uint32_t m0, uint32_t m1, uint32_t m2, uint32_t m3) {
    uint32_t syms;
    if (!WIDE) {
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(syms) : "r"(m0), "r"(m1), "r"(e));
    } else {
        uint32_t slo, shi, msk;
        const uint32_t e7 = e & 0x7777u;
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(slo) : "r"(m0), "r"(m1), "r"(e7));
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(shi) : "r"(m2), "r"(m3), "r"(e7));
        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(msk) : "r"(0x0000FF00u), "r"(0u), "r"((e >> 3) & 0x1111u));
        syms = (slo & ~msk) | (shi & msk);
    }
    return syms;
}

// table entry of the window in the low bits of x; windows that start with the bits of the all-zero window's entry (x & zm == 0)
// are that entry (ez) and do not load
#ifndef MUA_DV_ZSKIP
#define MUA_DV_ZSKIP 1
#endif
__device__ __forceinline__ uint32_t dv_lookup(uint32_t tab_sa, uint32_t x, uint32_t wmask, uint32_t zm, uint32_t ez) {
    uint32_t e;
#if MUA_DV_ZSKIP
    asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.u32 %0, %3;\n\t@p ld.shared.u32 %0, [%1];\n\t}"
        : "=r"(e)
        : "r"(tab_sa + ((x & wmask) << 2)), "r"(x & zm), "r"(ez));
#else
    asm("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(tab_sa + ((x & wmask) << 2)));
#endif
    return e;
}

#ifndef MUA_DV_WARPS
#define MUA_DV_WARPS 20
#endif
constexpr int DV_WARPS = MUA_DV_WARPS;            // at most; the launch takes as many as fit beside the tables
constexpr int DV_LENS_B = (MUA_MAX_K * 16 + 127) / 128 * 128;
constexpr int DV_ROW_SKEW = 11;         // words between the tables of two codebook rows beyond 2^Wv (bank skew)

template <bool WIDE>
__global__ void __launch_bounds__(DV_WARPS * 32, 1) k_decode_var(const __grid_constant__ DecParams P) {
    extern __shared__ __align__(1024) uint8_t dsm[];
    const TabHdr* T = reinterpret_cast<const TabHdr*>(P.tab);
    const int K = T->K, Wv = T->Wv;
    if (T->S != P.S || K != P.K || T->Lmax != P.Lmax || T->decv_off == 0 || (WIDE != (T->S > 8))) {
        if (threadIdx.x == 0) dec_flag(P.status, MUA_DEC_BAD_TABLE);   // host view does not match the table block
        return;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int STR_W = P.var_str_w;                                                   // staged stream words per lane
    const int per_warp = 32 * STR_W * 4 + 32 * DG_OUT_B + 16;
    uint32_t* s_str = reinterpret_cast<uint32_t*>(dsm + warp * per_warp);
    uint8_t* s_out = dsm + warp * per_warp + 32 * STR_W * 4;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_out + 32 * DG_OUT_B);
    uint32_t* s_map = reinterpret_cast<uint32_t*>(dsm + nwarps * per_warp);          // uint4 [MUA_MAX_S]: idx[p][0..15]
    uint32_t* s_ticket = s_map + 4 * MUA_MAX_S;                                      // next group of this CTA (4 words reserved)
    uint8_t* s_lens = reinterpret_cast<uint8_t*>(s_ticket + 4);                      // SCLV rows: uint8 [MUA_MAX_K][16]
    uint32_t* s_tab = reinterpret_cast<uint32_t*>(s_lens + DV_LENS_B);
    if (lane == 0) {
        mbar_init(s_bar, 1);
        fence_barrier_init();
    }
    if (threadIdx.x == 0) *s_ticket = 0;
    {
        const uint32_t* g = reinterpret_cast<const uint32_t*>(P.tab + T->decv_off);
        // entry of window v of row k at word k * (2^Wv + DV_ROW_SKEW) + bitreverse(v): the index is the window with its FIRST
        // stream bit in bit 0, so the bank of a lookup is decided by the first five bits of the window (the symbols being
        // decoded) and rows are skewed against each other -- the all-zero window of every row and the windows with one early
        // non-zero symbol, which make up most lookups of MUA counts, fall into different banks.  (Folding the upper index bits
        // onto the low five as well -- idx ^ idx >> 5 ^ idx >> 10, so that windows whose first five bits are zero leave bank 0
        // too -- cuts the remaining conflicts but its three instructions sit on the lookup chain: 2.08 -> 2.19 ms, rejected.)
        for (int i = threadIdx.x; i < (K << Wv); i += blockDim.x) {
            const int k = i >> Wv, v = i & ((1 << Wv) - 1);
            s_tab[k * ((1 << Wv) + DV_ROW_SKEW) + (int)(__brev((uint32_t)v) >> (32 - Wv))] = g[i];
        }
        const uint32_t* gi = reinterpret_cast<const uint32_t*>(&T->idx[0][0]);
        for (int i = threadIdx.x; i < 4 * MUA_MAX_S; i += blockDim.x) s_map[i] = gi[i];
        for (int i = threadIdx.x; i < K * 16; i += blockDim.x) s_lens[i] = T->lens[i >> 4][i & 15];
    }
    __syncthreads();
    uint32_t parity = 0;
    // the host sized the staged row: 127 bits of alignment slack + one worst-case period (128 * Lmax bits) + 96 bits of look-ahead
    // + var_extra bits; a stage is decoded period by period for as long as every lane still holds a worst-case period
    const uint32_t need_bits = 128u * (uint32_t)T->Lmax + 96u;
    const long long nitems = dec_nitems(P);
    const long long ngroups = (nitems + 31) / 32;
    const uint32_t slot_bytes = (uint32_t)P.slot_bytes;
    const uint32_t wmask = (1u << Wv) - 1u;
    const int row_words = (1 << Wv) + DV_ROW_SKEW;

    // groups by ticket (the warps of a CTA do not sit evenly on the four schedulers): ticket t = group blockIdx + t * grid
    for (;;) {
        uint32_t tk = 0;
        if (lane == 0) tk = atomicAdd(s_ticket, 1u);
        tk = __shfl_sync(FULL, tk, 0);
        const long long g = (long long)blockIdx.x + (long long)tk * gridDim.x;
        if (g >= ngroups) break;
        const long long item = g * 32 + lane;
        int rem = 0;
        uint32_t bitpos = 0;
        const uint8_t* sbase = P.stream;
        uint8_t* optr = P.dec;
        const uint32_t* tab = s_tab;
        const uint8_t* lens_row = s_lens;
        uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
        int c, j;
        if (dec_item(P, item, nitems, c, j)) {
            const int start = P.start[c], end = P.end[c];
            if (end > start && start >= 0) {
                const int j0 = start / TILE;
                const int nch = (end + TILE - 1) / TILE - j0;
                if (j < nch) {
                    const int a = max(start, (j0 + j) * TILE), b = min(end, (j0 + j + 1) * TILE);
                    const uint32_t bp0 = P.chunk_off[(size_t)c * P.chunk_stride + j];
                    const int pk = P.peak[c], en = P.enc[c];
                    if (pk >= T->S || en >= K) {
                        dec_flag(P.status, MUA_DEC_BAD_TABLE);
                    } else if (((bp0 >> 7) << 4) >= slot_bytes) {
                        dec_flag(P.status, MUA_DEC_BAD_OFFSET);
                    } else {
                        rem = b - a;
                        bitpos = bp0;
                        sbase = P.stream + (size_t)c * P.slot_bytes;
                        optr = P.dec + (P.off ? P.off[c] : (int64_t)c * P.stride) + a;
                        tab += en * row_words;
                        lens_row += en * 16;
                        const uint4 mp = reinterpret_cast<const uint4*>(s_map)[pk];
                        m0 = mp.x; m1 = mp.y; m2 = mp.z; m3 = mp.w;
                    }
                }
            }
        }
        // The entry of the all-zero window (four symbols of rank 0 whenever their codewords fit) is the entry of EVERY window that
        // starts with the bits it uses -- most windows of MUA counts.  Those lookups are answered from a register and their lanes
        // take no part in the table load: fewer distinct addresses per load, fewer bank conflicts (the loads were 60 % of the kernel's
        // shared-memory wavefronts, half of them conflicts).
        const uint32_t tab_sa = smem_u32(tab);
        const uint32_t ez = tab[0];
        const uint32_t zm = (ez & 0x40000u) ? ((1u << ((ez >> 20) & 0xFu)) - 1u) : wmask;
        int done = 0;                                            // symbols already written out
        while (__any_sync(FULL, rem > 0)) {
            // ---- stage one period's stream bytes per lane (one TMA bulk copy each), from the 16-byte unit holding `bitpos` ----
            uint32_t cur_al = (bitpos >> 7) << 4;
            if (rem > 0 && cur_al >= slot_bytes) {               // ran past the slot (corrupt stream): drop the rest of the chunk
                dec_flag(P.status, MUA_DEC_BAD_OFFSET);
                rem = 0;
                cur_al = 0;
            }
            const uint32_t nbytes = rem > 0 ? min((uint32_t)(STR_W * 4), slot_bytes - cur_al) : 0u;
            const uint32_t total = __reduce_add_sync(FULL, nbytes);
            __syncwarp();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (total) {
                if (lane == 0) mbar_expect_tx(s_bar, total);
                __syncwarp();
                if (nbytes) tma_load_1d(s_str + lane * STR_W, sbase + cur_al, nbytes, s_bar);
                mbar_wait(s_bar, parity);
                parity ^= 1;
            }
            const uint32_t boff = bitpos - cur_al * 8;           // 0..127
            const uint32_t* rp = s_str + lane * STR_W + (boff >> 5);
            // the stream is kept bit-reversed: stream bit j of a word at register bit j, w0 = the word holding the position
            uint32_t w0 = stream_rev(rp[0]), w1 = stream_rev(rp[1]);
            rp += 2;
            uint32_t off = boff & 31;
            uint32_t consumed = 0;                               // bits consumed in this stage
            // this lane can decode another period from the staged row while a worst-case period (+ look-ahead) is still inside
            // it -- or the row already reaches the end of the slot
            const uint32_t staged_bits = (cur_al + nbytes >= slot_bytes) ? 0xFFFFFFFFu : nbytes * 8u;

            for (bool more = true; more;) {
                // ---- 128 symbols per lane into the output tile: 16 steps of 8 symbols = two table lookups ----
                // A lookup returns as many whole symbols as its Wv-bit window holds, at most 4.  The common case -- both
                // lookups of a step return 4 (the frequent MUA counts code in 1..2 bits) -- has fixed output positions and no
                // per-symbol work: the second lookup is issued speculatively at the bit offset the first one reports.  When
                // any lane of the warp meets a window with fewer than 4 symbols, the step is redone quad by quad: the entry's
                // symbols are kept and the rest of the quad is decoded one symbol at a time through the same table (first
                // symbol of the entry at the running offset, length from the row's SCLV).
                uint4* orow = reinterpret_cast<uint4*>(s_out + lane * DG_OUT_B);
                uint32_t sA0 = 0, sB0 = 0;                       // symbols of the even step: the tile takes 16 bytes at a time
#pragma unroll 2
                for (int q = 0; q < 16; ++q) {
                    const uint32_t x = __funnelshift_r(w0, w1, off);             // next 32 stream bits, first one at bit 0
                    const uint32_t eA = dv_lookup(tab_sa, x, wmask, zm, ez);
                    const uint32_t uA = (eA >> 20) & 0xFu;
                    const uint32_t eB = dv_lookup(tab_sa, x >> uA, wmask, zm, ez);
                    uint32_t sA = dv_syms<WIDE>(eA, m0, m1, m2, m3), sB = dv_syms<WIDE>(eB, m0, m1, m2, m3);
                    uint32_t used = uA + ((eB >> 20) & 0xFu);
                    const bool esc = ((eA & eB) & 0x40000u) == 0u && q * 8 < rem;   // count field [18:16] == 4 <=> bit 18
                    if (__any_sync(FULL, esc)) {
                        uint32_t e = eA;
#pragma unroll 1
                        for (int h = 0; h < 2; ++h) {
                            uint32_t syms = dv_syms<WIDE>(e, m0, m1, m2, m3);
                            uint32_t u = (e >> 20) & 0xFu;                         // the entry's symbols: <= Wv bits
                            uint32_t j = (e >> 16) & 7u;
                            for (;;) {                                             // per-lane trip count 1..4
                                off += u;
                                consumed += u;
                                if (off >= 32u) { w0 = w1; w1 = stream_rev(*rp); ++rp; off -= 32u; }
                                if (j >= 4u) break;
                                const uint32_t r = tab[__funnelshift_r(w0, w1, off) & wmask] & 0xFu;   // first symbol at the running offset
                                syms |= dv_syms<WIDE>(r | (WIDE ? 0xFFF0u : 0x8880u), m0, m1, m2, m3) << (8u * j);
                                u = lens_row[r];
                                ++j;
                            }
                            if (h == 0) sA = syms; else sB = syms;
                            e = tab[__funnelshift_r(w0, w1, off) & wmask];  // entry of the second quad at its true offset
                        }
                    } else {
                        off += used;
                        consumed += used;
                        if (off >= 32u) { w0 = w1; w1 = stream_rev(*rp); ++rp; off -= 32u; }
                    }
                    // 16-byte stores: with the 144-byte row stride a quarter warp covers all 32 banks (8-byte stores conflict 2-way)
                    if (q & 1) orow[q >> 1] = make_uint4(sA0, sB0, sA, sB);
                    else { sA0 = sA; sB0 = sB; }
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
                            for (int kk = 0; kk < nbyte; ++kk) d[kk] = sp[kk];
                        }
                    }
                }
                __syncwarp();
                if (rem > 0) { rem -= 128; done += 128; }
                more = __any_sync(FULL, rem > 0) && __all_sync(FULL, rem <= 0 || boff + consumed + need_bits <= staged_bits);
            }
            bitpos += consumed;
            if (rem <= 0) rem = 0;
        }
    }
    dec_wait_report(P);
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
    extern __shared__ __align__(1024) uint8_t dsm[];
    const TabHdr* T = reinterpret_cast<const TabHdr*>(P.tab);
    const int K = T->K, W = T->W;
    if (T->S != P.S || K != P.K || T->Lmax != P.Lmax || T->nsym != 4 || T->Lmax > 2) {
        if (threadIdx.x == 0) dec_flag(P.status, MUA_DEC_BAD_TABLE);   // host view does not match the table block
        return;
    }
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
    const long long nitems = dec_nitems(P);
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
        int c, j;
        if (dec_item(P, item, nitems, c, j)) {
            const int start = P.start[c], end = P.end[c];
            if (end > start && start >= 0) {
                const int j0 = start / TILE;
                const int nch = (end + TILE - 1) / TILE - j0;
                if (j < nch) {
                    const int a = max(start, (j0 + j) * TILE), b = min(end, (j0 + j + 1) * TILE);
                    const uint32_t bp0 = P.chunk_off[(size_t)c * P.chunk_stride + j];
                    const int pk = P.peak[c], en = P.enc[c];
                    if (pk >= T->S || en >= K) {
                        dec_flag(P.status, MUA_DEC_BAD_TABLE);
                    } else if (((bp0 >> 7) << 4) >= slot_bytes) {
                        dec_flag(P.status, MUA_DEC_BAD_OFFSET);
                    } else {
                        rem = b - a;
                        bp = bp0;
                        sbase = P.stream + (size_t)c * P.slot_bytes;
                        optr = P.dec + (P.off ? P.off[c] : (int64_t)c * P.stride) + a;
                        lut += (size_t)(pk * K + en) << W;
                    }
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
    dec_wait_report(P);
}

// ---- fast decoder with lane-private LUT banks (NSYM = 4, W = 8, K <= 3, S <= 8: the chosen system) ----
// k_decode_fast is bound by the LSU data pipe, and ~45 % of its shared-memory wavefronts are bank conflicts of
// the LUT gathers (32 lanes index a 256-word table with skewed, data-dependent indices: ~3.7 wavefronts per
// lookup).  Here every lane owns a bank: entry i of codebook row k for lane l lives at byte k*32K + i*128 + l*4,
// so a lookup is ONE wavefront whatever the 32 indices are.  A replicated table takes 32 KB, so there is one per
// codebook ROW only: the tables hold RANKS (the table of peak 0, whose rank map is the identity) and the lane's
// peak is applied by the PRMT that unpacks an entry -- the entry's low 16 bits are four byte selectors into the
// lane's rank -> symbol map (TabHdr::idx[p], 8 bytes in two registers), which costs the same one instruction as
// the mask it replaces.  What is left of the SM's shared memory feeds 20 warps (K = 1) only because the per-lane
// stream staging is a 128-byte ring instead of the whole chunk.  The kernel is latency bound (one dependent
// chain per lane), so the chain of a lookup and everything around it is kept short:
//   * the stream is kept bit-reversed in registers (stream bit j at register bit j) and the tables are indexed
//     by the bit-reversed window, so the window of lookup k lands on address bits [14:7] with ONE funnel shift
//     by the running bit count; the tables are 32 KB aligned in the shared window so that one LOP3 masks the
//     index and ORs the lane's table address in: chain = SHF -> LOP3 -> LDS -> LEA.HI;
//   * the stream reaches each lane through a 128-byte ring that is topped up with cp.async (LDGSTS) one
//     128-symbol period ahead of its use; the next group's first bytes are requested before the last period of
//     the current group is written out;
//   * the write-out (8 lanes per 128-byte row) uses per-group precomputed row pointers and store counts (for how many
//     periods a pass stores a complete aligned 16 bytes); periods with window edges or unaligned rows are flagged
//     warp-wide and take a general byte-store path.
// (Round 2 tried the other way out of the tile: the decoded buffer described to the TMA engine as a 3-D tensor (byte in chunk, chunk,
// channel), one 128 x 8 box store per eight lanes and period from a 128-byte-swizzled dense tile -- commit d9cb5c0 of this repository's
// history, "Experiment: lane decoder with TMA tensor stores".  Bit-exact, LSU data pipe 60 -> 41 %, 90 registers instead of 120, and
// SLOWER: 1.18 vs 1.13 ms.  The write-out is not what bounds this kernel; see profiles/r02_summary.md.)
#ifndef MUA_DL_TICKETS
#define MUA_DL_TICKETS 1
#endif
#ifndef MUA_DL_WARPS
#define MUA_DL_WARPS 14        // 8: 1.67 ms, 10: 1.45, 12: 1.32, 14: 1.20, 16: 1.22, 18: 1.27 (96 registers, spills), 20: 1.35
#endif
constexpr int DL_WARPS = MUA_DL_WARPS;           // launched warps (more do not help: see profiles/r01_summary.md); those with a buffer in the runtime layout work
constexpr int DL_ROW_B = 144;          // stream ring row: 128 B + 16 B pad
constexpr int DL_OUT_B = 144;          // output tile row: 128 B + 16 B pad (row starts 4 banks apart)
constexpr int DL_MAX_ROWS = 3;         // codebook rows (K) whose lane-replicated tables fit
constexpr int DL_TAB_B = 256 * 32 * 4; // one lane-replicated table: 32 KB
constexpr int DL_PER_WARP = 32 * DL_ROW_B + 32 * DL_OUT_B;
constexpr int DL_SMEM = 227 * 1024;    // everything the SM has: tables sit at 32 KB aligned addresses, warps around them

struct DecRaw {              // the loads of one lane's chunk bookkeeping, nothing derived yet
    int c, j, valid;
    int start, end, pk, en;
    uint32_t bp;
    long long off;
};
struct DecItem {
    int col, ch;             // absolute bin index of the chunk's first symbol, channel
    int rem;                 // symbols of this lane's chunk (0: nothing to do)
    uint32_t bp;             // bit position of the chunk in the channel's stream
    const uint8_t* sbase;    // the channel's slot
    uint8_t* optr;           // where the chunk's first symbol goes
    int pk, en;              // peak (rank -> symbol map) and codebook row (table)
};

__device__ __forceinline__ DecRaw dec_load_cj(const DecParams& P, int c, int j, bool valid) {
    DecRaw r;
    r.valid = valid;
    r.c = valid ? c : 0;
    r.j = valid ? j : 0;
    r.start = ldg_s32(P.start + r.c);
    r.end = ldg_s32(P.end + r.c);
    r.pk = ldg_u8(P.peak + r.c);
    r.en = ldg_u8(P.enc + r.c);
    r.bp = ldg_u32(P.chunk_off + (size_t)r.c * P.chunk_stride + r.j);
    r.off = P.off ? ldg_s64(P.off + r.c) : (long long)r.c * P.stride;
    return r;
}

__device__ __forceinline__ DecRaw dec_load(const DecParams& P, long long item, long long nitems) {
    int c, j;
    const bool valid = dec_item(P, item, nitems, c, j);
    return dec_load_cj(P, c, j, valid);
}

__device__ __forceinline__ DecItem dec_finish(const DecParams& P, DecRaw r, int K, int S) {
    // nothing derived from the loaded values may be scheduled before this point
    asm volatile("" : "+r"(r.start), "+r"(r.end), "+r"(r.pk), "+r"(r.en), "+r"(r.bp), "+l"(r.off));
    DecItem it;
    it.col = 0; it.ch = 0; it.rem = 0; it.bp = 0; it.sbase = P.stream; it.optr = P.dec; it.pk = 0; it.en = 0;
    if (r.valid && r.end > r.start && r.start >= 0) {
        const int j0 = r.start / TILE;
        const int nch = (r.end + TILE - 1) / TILE - j0;
        if (r.j < nch && (r.pk >= S || r.en >= K)) {
            dec_flag(P.status, MUA_DEC_BAD_TABLE);
        } else if (r.j < nch && (long long)(r.bp >> 3) >= P.slot_bytes) {
            dec_flag(P.status, MUA_DEC_BAD_OFFSET);
        } else if (r.j < nch) {
            const int a = max(r.start, (j0 + r.j) * TILE), b = min(r.end, (j0 + r.j + 1) * TILE);
            it.col = a;
            it.ch = r.c;
            it.rem = b - a;
            it.bp = r.bp;
            it.sbase = P.stream + (size_t)r.c * P.slot_bytes;
            it.optr = P.dec + r.off + a;
            it.pk = r.pk;
            it.en = r.en;
        }
    }
    return it;
}

__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

// One lane's stream ring: 128 bytes, byte b of the stream (counted from the ring origin, a 32-byte aligned
// address at or before the chunk's first bit) lives at ring offset b mod 128.  The ring is topped up with two
// 16-byte cp.async copies per 128-symbol period (<= 32 bytes are consumed per period), issued one period ahead
// of their use, so no staging latency is exposed while a chunk is being decoded.
struct DecRing {
    unsigned long long org;   // global address of the ring origin
    uint32_t wp;              // bytes issued so far, from the origin (multiple of 32)
};

// decoded-symbol store (st.global.cs / .wt instead of the default write-back operator: no difference, 1.204 / 1.207 / 1.210 ms)
__device__ __forceinline__ void st_out16(uint8_t* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }
__device__ __forceinline__ void cp_async16(uint32_t dst, unsigned long long src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// NC = chunks decoded concurrently by one lane (independent dependency chains interleaved in one instruction
// stream).  The shared-memory budget fixes the number of chains per SM (14 * 32), so NC = 2 runs 7 warps.
template <int NC>
__global__ void __launch_bounds__(DL_WARPS / NC * 32, 1) k_decode_lane(const __grid_constant__ DecParams P) {
    extern __shared__ __align__(1024) uint8_t dsm[];
    const TabHdr* T = reinterpret_cast<const TabHdr*>(P.tab);
    const int K = T->K;
    if (T->S != P.S || K != P.K || T->Lmax != P.Lmax || T->nsym != 4 || T->W != 8 || K > DL_MAX_ROWS || T->S > 8) {
        if (threadIdx.x == 0) dec_flag(P.status, MUA_DEC_BAD_TABLE);   // host view does not match the table block
        return;
    }
    const uint32_t* g_lut = reinterpret_cast<const uint32_t*>(P.tab + T->dec_off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // ---- shared-memory layout: rank -> symbol maps, then per-(warp, chain) buffers below and above the tables,
    //      which sit at the first 32 KB boundary ----
    const uint32_t map_a = smem_u32(dsm);                         // uint2 [8]: idx[p][0..7]
    const uint32_t base_a = map_a + 128;
    const uint32_t tab_a = (base_a + (DL_TAB_B - 1)) & ~(uint32_t)(DL_TAB_B - 1);
    const uint32_t hi_a = tab_a + K * DL_TAB_B;
    const uint32_t end_a = map_a + DL_SMEM;
    const int n_low = (int)((tab_a - base_a) / DL_PER_WARP);
    const int n_high = hi_a <= end_a ? (int)((end_a - hi_a) / DL_PER_WARP) : 0;
    const int nw = min(DL_WARPS, n_low + n_high) / NC;            // warps that have their buffers
    {
        // table of row k = the dec table of (peak 0, row k), whose symbols are the ranks; entry = four rank nibbles
        // (PRMT byte selectors) in the low 16 bits | used bits << 28; indexed by the bit-reversed window
        uint32_t* s_lut = reinterpret_cast<uint32_t*>(dsm + (tab_a - map_a));
        const int nword = K * 256 * 32;
        for (int i = threadIdx.x; i < nword; i += blockDim.x) {
            const int e = i >> 5, k = e >> 8, idx = e & 255;
            const uint32_t v = g_lut[k * 256 + idx];
            const uint32_t sel = (v & 0xFu) | ((v >> 4) & 0xF0u) | ((v >> 8) & 0xF00u) | ((v >> 12) & 0xF000u);
            s_lut[(k * 256 + (int)(__brev((uint32_t)idx) >> 24)) * 32 + (i & 31)] = sel | (v & 0xF0000000u);
        }
        if (threadIdx.x < 16) {
            const int p = threadIdx.x >> 1, h = threadIdx.x & 1;
            uint32_t m = 0;
            if (p < T->S)
                for (int r = 0; r < 4; ++r) m |= (uint32_t)T->idx[p][4 * h + r] << (8 * r);
            reinterpret_cast<uint32_t*>(dsm)[threadIdx.x] = m;
        }
        if (threadIdx.x == 16) reinterpret_cast<uint32_t*>(dsm)[16] = (uint32_t)nw;     // next group ticket of this CTA
    }
    __syncthreads();
    if (warp >= nw) return;
    uint32_t buf_a[NC], ring_a[NC];
    const uint32_t* rowp[NC];
    uint8_t* s_out[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int u = warp * NC + c;
        buf_a[c] = u < n_low ? base_a + u * DL_PER_WARP : hi_a + (u - n_low) * DL_PER_WARP;
        uint8_t* s_str = dsm + (buf_a[c] - map_a);
        s_out[c] = s_str + 32 * DL_ROW_B;
        rowp[c] = reinterpret_cast<const uint32_t*>(s_str + lane * DL_ROW_B);
        ring_a[c] = buf_a[c] + lane * DL_ROW_B;
    }
    const long long nitems = dec_nitems(P);
    const long long ngroups = (nitems + 32 * NC - 1) / (32 * NC);
    const int wrow = lane >> 3, wcol = lane & 7;
    const unsigned long long lo_addr = reinterpret_cast<unsigned long long>(P.stream);
    const unsigned long long hi_addr = lo_addr + (unsigned long long)P.C * (unsigned long long)P.slot_bytes;

    // first 96 bytes of a chunk (>= 65 bytes past its first bit); pieces outside the stream buffer are skipped
    auto ring_start = [&](const DecItem& it, DecRing& R, uint32_t ra) {
        R.org = (reinterpret_cast<unsigned long long>(it.sbase) + (it.bp >> 3)) & ~31ull;
        R.wp = 96;
        if (it.rem > 0) {
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const unsigned long long a = R.org + 16 * k;
                if (a >= lo_addr && a + 16 <= hi_addr) cp_async16(ra + 16 * k, a);
            }
        }
    };

    // Groups are handed out by ticket: 14 warps are 4 + 4 + 3 + 3 per scheduler, so warps do not run at the same speed;
    // ticket t of a CTA is the group the static schedule would have given warp t % nw in round t / nw.
#if MUA_DL_TICKETS
    uint32_t* s_ticket = reinterpret_cast<uint32_t*>(dsm) + 16;
#endif
    const long long gstride = (long long)gridDim.x * nw;
    long long g = (long long)blockIdx.x * nw + warp;
    DecItem cur[NC];
    DecRing R[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        cur[c] = dec_finish(P, dec_load(P, (g * NC + c) * 32 + lane, g < ngroups ? nitems : 0), K, T->S);
        ring_start(cur[c], R[c], ring_a[c]);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    while (g < ngroups) {
        long long gn = g + gstride;
        uint32_t lbase[NC], mlo[NC], mhi[NC], rp[NC], off[NC], w0[NC], w1[NC], wn[NC];
        int rem[NC];
        // write-out roles: in pass i this lane stores 16 bytes of row 4i + wrow of every chain's tile.  Per group and pass
        // the lane knows for how many periods that store is a complete, aligned 16-byte store (nibble i of wnf); the
        // periods in which any lane of the warp has something else to store (a window edge inside its 16 bytes, an
        // unaligned first chunk) are flagged warp-wide in wslow and take the general path.
        unsigned long long wptr[NC][8];
        uint32_t wnf[NC], wslow[NC];
        int maxrem = 0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            lbase[c] = tab_a + (uint32_t)cur[c].en * DL_TAB_B + lane * 4;             // this lane's bank of its row's table
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(mlo[c]), "=r"(mhi[c]) : "r"(map_a + cur[c].pk * 8));
            rem[c] = cur[c].rem;
            maxrem = max(maxrem, rem[c]);
            wnf[c] = 0;
            uint32_t slow = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = i * 4 + wrow;
                wptr[c][i] = __shfl_sync(FULL, reinterpret_cast<unsigned long long>(cur[c].optr), r) + wcol * 16;
                const int R = __shfl_sync(FULL, rem[c], r) - wcol * 16;              // valid bytes from this lane's column on
                if ((wptr[c][i] & 15) == 0) {
                    const int nf = R >= 16 ? ((R - 16) >> 7) + 1 : 0;                 // periods p with R - 128 p >= 16
                    wnf[c] |= (uint32_t)nf << (4 * i);
                    if (R - 128 * nf > 0) slow |= 1u << nf;                           // 1..15 bytes left in the period after
                } else if (R > 0) {
                    slow |= (1u << ((R + 127) >> 7)) - 1u;                            // unaligned row: byte stores throughout
                }
            }
            wslow[c] = __reduce_or_sync(FULL, slow);
            // position of the chunk's first bit in the ring frame
            const uint32_t boff =
                (uint32_t)(reinterpret_cast<unsigned long long>(cur[c].sbase) + (cur[c].bp >> 3) - R[c].org) * 8 + (cur[c].bp & 7);
            rp[c] = boff >> 5;
            off[c] = boff & 31;
            w0[c] = w1[c] = wn[c] = 0;
        }
        int done = 0;
        const int nper = max((__reduce_max_sync(FULL, maxrem) + 127) >> 7, 1);        // periods of this group
        for (int per = 0; per < nper; ++per) {
            // ---- top the rings up for the NEXT period, then wait for everything issued before that ----
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                if (rem[c] > done + 128) {
                    const uint32_t rd32 = (done == 0 ? rp[c] : rp[c] - 3) * 4 & ~31u;  // read position, rounded down to 32 bytes
                    if (R[c].wp + 32 - rd32 <= 128) {
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const unsigned long long a = R[c].org + R[c].wp + 16 * k;
                            if (a + 16 <= hi_addr) cp_async16(ring_a[c] + ((R[c].wp + 16 * k) & 127), a);
                        }
                        R[c].wp += 32;
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            if (done == 0) {
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    w0[c] = stream_rev(rowp[c][rp[c]]); w1[c] = stream_rev(rowp[c][rp[c] + 1]);
                    wn[c] = rowp[c][rp[c] + 2];                                          // kept raw, see the refill below
                    rp[c] += 3;
                }
            }
            // ---- 128 symbols per lane and chain into the output tiles ----
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                uint32_t ow[NC][4], xl[NC], xh[NC], o[NC];
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const uint32_t xr = __funnelshift_r(w0[c], w1[c], off[c]);       // next 32 stream bits, first one at bit 0
                    xl[c] = xr << 7; xh[c] = xr >> 25;
                    o[c] = 0;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const uint32_t t = __funnelshift_r(xl[c], xh[c], o[c]);      // window of this lookup on bits [14:7]
                        uint32_t adr;                                                    // (t & 0x7F80) | lbase in one LOP3
                        asm("lop3.b32 %0, %1, 0x7F80, %2, 0xEA;" : "=r"(adr) : "r"(t), "r"(lbase[c]));
                        const uint32_t e = lds_u32(adr);
                        asm("prmt.b32 %0, %1, %2, %3;" : "=r"(ow[c][k]) : "r"(mlo[c]), "r"(mhi[c]), "r"(e));   // ranks -> symbols
                        o[c] += e >> 28;
                    }
                }
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    off[c] += o[c];
                    {   // branch-free refill: the word after next is read every time and kept RAW (it is bit-reversed when it
                        // moves into w1, a refill later, so nothing ever waits for this load)
                        const bool rf = off[c] >= 32;
                        const uint32_t nxt = rowp[c][rp[c] & 31];
                        const uint32_t w1n = stream_rev(wn[c]);
                        w0[c] = rf ? w1[c] : w0[c];
                        w1[c] = rf ? w1n : w1[c];
                        wn[c] = rf ? nxt : wn[c];
                        rp[c] += rf ? 1u : 0u;
                        off[c] -= rf ? 32u : 0u;
                    }
                    reinterpret_cast<uint4*>(s_out[c] + lane * DL_OUT_B)[q] = make_uint4(ow[c][0], ow[c][1], ow[c][2], ow[c][3]);
                }
            }
            if (per == nper - 1) {   // every ring of the warp is free: start the next group's chunks before writing this period out
#if MUA_DL_TICKETS
                uint32_t tk = 0;
                if (lane == 0) tk = atomicAdd(s_ticket, 1u);
                tk = __shfl_sync(FULL, tk, 0);
                gn = (long long)blockIdx.x * nw + (tk % (uint32_t)nw) + (long long)(tk / (uint32_t)nw) * gstride;
#endif
                const long long nlim = gn < ngroups ? nitems : 0;
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    cur[c] = dec_finish(P, dec_load(P, (gn * NC + c) * 32 + lane, nlim), K, T->S);
                    ring_start(cur[c], R[c], ring_a[c]);
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            // ---- coalesced write-out: 8 lanes per 128-byte row, 4 rows per pass ----
            __syncwarp();
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                uint4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const uint4*>(s_out[c] + (i * 4 + wrow) * DL_OUT_B + wcol * 16);
                if (!((wslow[c] >> per) & 1u)) {   // warp-uniform: every store of this period is a complete aligned 16-byte store or nothing
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if ((uint32_t)per < ((wnf[c] >> (4 * i)) & 15u)) st_out16(reinterpret_cast<uint8_t*>(wptr[c][i]) + done, v[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = i * 4 + wrow;
                        const int vr = __shfl_sync(FULL, rem[c], r) - wcol * 16 - done;   // valid bytes from this lane's column on
                        uint8_t* d = reinterpret_cast<uint8_t*>(wptr[c][i]) + done;
                        if (vr >= 16 && (wptr[c][i] & 15) == 0) {
                            *reinterpret_cast<uint4*>(d) = v[i];
                        } else if (vr >= 16 && (wptr[c][i] & 7) == 0) {   // first chunk of a window that starts on an 8-byte boundary
                            reinterpret_cast<uint2*>(d)[0] = make_uint2(v[i].x, v[i].y);
                            reinterpret_cast<uint2*>(d)[1] = make_uint2(v[i].z, v[i].w);
                        } else if (vr >= 16 && (wptr[c][i] & 3) == 0) {   // ... on a 4-byte boundary
                            reinterpret_cast<uint32_t*>(d)[0] = v[i].x; reinterpret_cast<uint32_t*>(d)[1] = v[i].y;
                            reinterpret_cast<uint32_t*>(d)[2] = v[i].z; reinterpret_cast<uint32_t*>(d)[3] = v[i].w;
                        } else if (vr > 0) {   // window edge or odd start: byte stores
                            const uint8_t* sp = s_out[c] + r * DL_OUT_B + wcol * 16;
                            const int nbyte = min(16, vr);
                            for (int k = 0; k < nbyte; ++k) d[k] = sp[k];
                        }
                    }
                }
            }
            __syncwarp();
            done += 128;
        }
        g = gn;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    dec_wait_report(P);
}

// ---- sub-chunk decoder (the chosen system's codebook class, fixed row stride, the encoder's 128-symbol side info) ----
// k_decode_lane gives every lane a 1024-symbol chunk, so a warp writes 32 rows of 128 bytes that lie 1 KB apart in every pass, and
// its per-lane stream ring and output row limit an SM to 14 warps of dependent lookup chains.  With a bit offset for every
// 128-symbol sub-chunk (uint32: +3.1 % of the symbol bytes) a lane decodes ONE sub-chunk and a warp 32
// consecutive ones of a channel:
//   * the warp's output of a pass is 4 KB of consecutive symbols, written with fully coalesced 16-byte stores from a dense,
//     128-byte-swizzled tile (piece q of row l at position q ^ (l & 7): conflict-free for the lanes' stores and for the write-out);
//   * the 32 lanes' stream bytes are consecutive too (<= 1 KB + slack): the warp copies them with cp.async (16 bytes per lane and
//     round) into one of two staging buffers, one group ahead; no per-lane ring, no top-ups, no refill bookkeeping beyond a pointer;
//   * per-warp shared memory drops to 6.1 KB and the state of a lane to a handful of registers: 28 warps (896 chains) per SM (30 or 31 fit and are no faster).
// Same lane-private-bank tables, rank -> symbol PRMT and lookup chain as k_decode_lane.  Groups never span channels (a channel's
// sub-chunks are padded to a multiple of 32); the side info of the next two groups and the stream of the next group are requested
// while the current group is decoded.  Partial sub-chunks (window start / end inside) and rows that are not 16-byte aligned take
// a per-row path with predicated, alignment-dependent stores.
#ifndef MUA_DS_WARPS
#define MUA_DS_WARPS 28
#endif
constexpr int DS_WARPS = MUA_DS_WARPS;
constexpr int DS_STR_B = 1088;                     // staged stream bytes per group: 31 x 32 B between the first and the last lane's
                                                   // first byte + 15 (alignment) + 48 + 15 (last lane's words, rounded up) <= 1072
constexpr int DS_TILE_B = 32 * 128;

// Shared memory of k_decode_sub: [rank maps + ticket: 128 B] ... [K tables at the first 32 KB boundary] ...; the n warps' tiles
// (4 KB, 1024-byte aligned: the swizzle) and then their pairs of stream buffers are bump-allocated below the tables first, then
// above them.  Returns false when warp w's buffers do not fit.
__device__ __forceinline__ bool ds_layout(int n, int w, uint32_t base_a, uint32_t tab_a, uint32_t hi_a, uint32_t end_a, uint32_t& tile,
                                          uint32_t& str) {
    uint32_t lo = (base_a + 1023u) & ~1023u, hi = hi_a;
    tile = 0; str = 0;
    for (int i = 0; i < n; ++i) {
        uint32_t a;
        if (lo + DS_TILE_B <= tab_a) { a = lo; lo += DS_TILE_B; }
        else if (hi + DS_TILE_B <= end_a) { a = hi; hi += DS_TILE_B; }
        else return false;
        if (i == w) tile = a;
    }
    for (int i = 0; i < n; ++i) {
        uint32_t a;
        if (lo + 2 * DS_STR_B <= tab_a) { a = lo; lo += 2 * DS_STR_B; }
        else if (hi + 2 * DS_STR_B <= end_a) { a = hi; hi += 2 * DS_STR_B; }
        else return false;
        if (i == w) str = a;
    }
    return true;
}

struct DsRaw {               // the loads of one lane's sub-chunk bookkeeping, nothing derived yet
    int c, mrel, valid;
    int start, end, pk, en;
    uint32_t sub;
};
struct DsItem {
    int rem;                 // symbols of this lane's sub-chunk (0: idle)
    uint32_t bitoff;         // bit offset of its first symbol in the channel's stream
    uint8_t* optr;           // where its first symbol goes
    const uint8_t* sbase;    // the channel's slot
    uint32_t info;           // peak | SCLV row << 4 | (window start / 1024) << 12
};

// the decoded buffer as a tensor for the TMA engine: uint8 [C][stride / 128][128], box 128 x 32 x 1, 128-byte swizzle
struct DecSubParams {
    DecParams D;
    alignas(64) CUtensorMap tmap;
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src_smem, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src_smem), "r"(x),
                 "r"(y), "r"(z)
                 : "memory");
}

__device__ __forceinline__ DsRaw ds_load(const DecParams& P, uint32_t g, uint32_t ngroups, uint32_t gpc, int lane) {
    DsRaw r;
    r.valid = g < ngroups;
    const uint32_t gg = r.valid ? g : 0u;
    r.c = (int)(gg / gpc);
    r.mrel = (int)(gg - (uint32_t)r.c * gpc) * 32 + lane;
    r.start = ldg_s32(P.start + r.c);
    r.end = ldg_s32(P.end + r.c);
    r.pk = ldg_u8(P.peak + r.c);
    r.en = ldg_u8(P.enc + r.c);
    r.sub = ldg_u32(P.sub_off + (size_t)r.c * P.sub_stride + r.mrel);
    return r;
}

__device__ __forceinline__ DsItem ds_finish(const DecParams& P, DsRaw r, int K, int S) {
    asm volatile("" : "+r"(r.start), "+r"(r.end), "+r"(r.pk), "+r"(r.en), "+r"(r.sub));
    DsItem it;
    it.rem = 0; it.bitoff = 0; it.optr = P.dec; it.sbase = P.stream; it.info = 0;
    if (r.valid && r.end > r.start && r.start >= 0) {
        const int mabs = 8 * (r.start >> 10) + r.mrel;                 // absolute sub-chunk index
        const int a = max(r.start, mabs << 7), b = min(r.end, (mabs << 7) + 128);
        if (b > a) {
            const uint32_t bo = r.sub;
            if (r.pk >= S || r.en >= K) {
                dec_flag(P.status, MUA_DEC_BAD_TABLE);
            } else if ((long long)(bo >> 3) >= P.slot_bytes) {
                dec_flag(P.status, MUA_DEC_BAD_OFFSET);
            } else {
                it.rem = b - a;
                it.bitoff = bo;
                it.sbase = P.stream + (size_t)r.c * P.slot_bytes;
                it.optr = P.dec + (long long)r.c * P.stride + a;
                it.info = (uint32_t)r.pk | ((uint32_t)r.en << 4) | ((uint32_t)(r.start >> 10) << 12);
            }
        }
    }
    return it;
}

#ifndef MUA_DS_TMA
#define MUA_DS_TMA 1        // complete groups leave through one TMA tensor store (0: 16-byte stores from the tile)
#endif
__global__ void __launch_bounds__(DS_WARPS * 32, 1) k_decode_sub(const __grid_constant__ DecSubParams PS) {
    extern __shared__ __align__(1024) uint8_t dsm[];
    const DecParams& P = PS.D;
    const TabHdr* T = reinterpret_cast<const TabHdr*>(P.tab);
    const int K = T->K;
    if (T->S != P.S || K != P.K || T->Lmax != P.Lmax || T->nsym != 4 || T->W != 8 || K > DL_MAX_ROWS || T->S > 8 || P.off != nullptr ||
        P.sub_off == nullptr) {
        if (threadIdx.x == 0) dec_flag(P.status, MUA_DEC_BAD_TABLE);   // host view does not match the table block
        return;
    }
    const uint32_t* g_lut = reinterpret_cast<const uint32_t*>(P.tab + T->dec_off);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // ---- shared-memory layout: rank -> symbol maps + ticket, per-warp units below and above the tables (32 KB aligned) ----
    const uint32_t map_a = smem_u32(dsm);                         // uint2 [8]: idx[p][0..7]; word 16: ticket
    const uint32_t base_a = map_a + 128;
    const uint32_t tab_a = (map_a + 128 + (DL_TAB_B - 1)) & ~(uint32_t)(DL_TAB_B - 1);
    const uint32_t hi_a = tab_a + K * DL_TAB_B;
    const uint32_t end_a = map_a + DL_SMEM;
    uint32_t tile_a = 0, str_a = 0;
    int nw = (int)(blockDim.x >> 5);                              // warps that have their buffers
    while (nw > 0 && !ds_layout(nw, nw - 1, base_a, tab_a, hi_a, end_a, tile_a, str_a)) --nw;
    if (nw == 0 || hi_a > end_a) {
        if (threadIdx.x == 0) dec_flag(P.status, MUA_DEC_BAD_TABLE);
        return;
    }
    {
        uint32_t* s_lut = reinterpret_cast<uint32_t*>(dsm + (tab_a - map_a));
        const int nword = K * 256 * 32;
        for (int i = threadIdx.x; i < nword; i += blockDim.x) {
            const int e = i >> 5, k = e >> 8, idx = e & 255;
            const uint32_t v = g_lut[k * 256 + idx];
            const uint32_t sel = (v & 0xFu) | ((v >> 4) & 0xF0u) | ((v >> 8) & 0xF00u) | ((v >> 12) & 0xF000u);
            s_lut[(k * 256 + (int)(__brev((uint32_t)idx) >> 24)) * 32 + (i & 31)] = sel | (v & 0xF0000000u);
        }
        if (threadIdx.x < 16) {
            const int p = threadIdx.x >> 1, h = threadIdx.x & 1;
            uint32_t m = 0;
            if (p < T->S)
                for (int r = 0; r < 4; ++r) m |= (uint32_t)T->idx[p][4 * h + r] << (8 * r);
            reinterpret_cast<uint32_t*>(dsm)[threadIdx.x] = m;
        }
        if (threadIdx.x == 16) reinterpret_cast<uint32_t*>(dsm)[16] = (uint32_t)(3 * nw);   // next group ticket of this CTA
    }
    __syncthreads();
    if (warp >= nw) return;
    ds_layout(nw, warp, base_a, tab_a, hi_a, end_a, tile_a, str_a);   // this warp's tile and its two stream buffers of DS_STR_B bytes
    const uint32_t* s_str = reinterpret_cast<const uint32_t*>(dsm + (str_a - map_a));
    const int col8 = lane & 7, row4 = lane >> 3;
    const uint32_t gpc = (uint32_t)((8 * P.item_chunks + 31) / 32);   // groups per channel
    const uint32_t ngroups = (uint32_t)P.C * gpc;
    const uint32_t slot_bytes = (uint32_t)P.slot_bytes;
    uint32_t* s_ticket = reinterpret_cast<uint32_t*>(dsm) + 16;
    const uint32_t gstride = gridDim.x * (uint32_t)nw;

    // group order: the static schedule for the first three rounds (g0 + k * gstride), tickets after that
    auto ticket_group = [&]() -> uint32_t {
        uint32_t tk = 0;
        if (lane == 0) tk = atomicAdd(s_ticket, 1u);
        tk = __shfl_sync(FULL, tk, 0);
        return blockIdx.x * (uint32_t)nw + (tk % (uint32_t)nw) + (tk / (uint32_t)nw) * gstride;
    };
    // request a group's stream bytes (the 16-byte units from the first lane's first bit to the last lane's last possible word)
    auto stage = [&](const DsItem& it, int buf, uint32_t& lo_out) {
        const uint32_t by = it.bitoff >> 3;
        const uint32_t lo = __reduce_min_sync(FULL, it.rem > 0 ? (by & ~15u) : 0xFFFFFFFFu);
        const uint32_t hi = __reduce_max_sync(FULL, it.rem > 0 ? ((by + 48u + 15u) & ~15u) : 0u);
        lo_out = lo;
        if (hi > lo) {                                            // some lane is active (all of one channel: same slot)
            const uint8_t* sb = reinterpret_cast<const uint8_t*>(__shfl_sync(FULL, reinterpret_cast<unsigned long long>(it.sbase),
                                                                             __ffs(__ballot_sync(FULL, it.rem > 0)) - 1));
            const uint32_t span = min(hi - lo, (uint32_t)DS_STR_B);
            for (uint32_t o = lane * 16u; o < span; o += 512u)
                if (lo + o + 16u <= slot_bytes) cp_async16(str_a + buf * DS_STR_B + o, reinterpret_cast<unsigned long long>(sb) + lo + o);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    uint32_t g0 = blockIdx.x * (uint32_t)nw + warp, g1 = g0 + gstride, g2 = g1 + gstride;
    DsItem cur = ds_finish(P, ds_load(P, g0, ngroups, gpc, lane), K, T->S);
    DsItem nxt = ds_finish(P, ds_load(P, g1, ngroups, gpc, lane), K, T->S);
    DsRaw raw = ds_load(P, g2, ngroups, gpc, lane);
    uint32_t lo_cur, lo_nxt;
    int buf = 0;
    stage(cur, 0, lo_cur);
    stage(nxt, 1, lo_nxt);
    while (g0 < ngroups) {
        // ---- this group's lane state ----
        const uint32_t lbase = tab_a + ((cur.info >> 4) & 0xFFu) * DL_TAB_B + lane * 4;   // this lane's bank of its row's table
        uint32_t mlo, mhi;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(mlo), "=r"(mhi) : "r"(map_a + (cur.info & 0xFu) * 8));
        const int rem = cur.rem;
        const unsigned long long optr_self = reinterpret_cast<unsigned long long>(cur.optr);
        const uint32_t* sp = s_str + buf * (DS_STR_B / 4);
        uint32_t rp = rem > 0 ? (cur.bitoff >> 5) - (lo_cur >> 2) : 0u;
        uint32_t off = rem > 0 ? (cur.bitoff & 31u) : 0u;
        asm volatile("cp.async.wait_group 1;" ::: "memory");     // this group's stream has landed (the next group's may be in flight)
        __syncwarp();
        uint32_t w0 = stream_rev(sp[rp]), w1 = stream_rev(sp[rp + 1]), wn = sp[rp + 2];
        rp += 3;
        const uint32_t trow_a = tile_a + lane * 128;
#if MUA_DS_TMA
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the TMA engine has read the previous group's tile
        __syncwarp();
#endif
        // ---- 128 symbols per lane into the tile ----
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            uint32_t ow[4];
            const uint32_t xr = __funnelshift_r(w0, w1, off);                        // next 32 stream bits, first one at bit 0
            const uint32_t xl = xr << 7, xh = xr >> 25;
            uint32_t o = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t t = __funnelshift_r(xl, xh, o);                       // window of this lookup on bits [14:7]
                uint32_t adr;                                                        // (t & 0x7F80) | lbase in one LOP3
                asm("lop3.b32 %0, %1, 0x7F80, %2, 0xEA;" : "=r"(adr) : "r"(t), "r"(lbase));
                const uint32_t e = lds_u32(adr);
                asm("prmt.b32 %0, %1, %2, %3;" : "=r"(ow[k]) : "r"(mlo), "r"(mhi), "r"(e));   // ranks -> symbols
                o += e >> 28;
            }
            off += o;
            {   // branch-free refill: the word after next is read every time and kept raw
                const bool rf = off >= 32;
                const uint32_t nx = sp[rp];
                const uint32_t w1n = stream_rev(wn);
                w0 = rf ? w1 : w0;
                w1 = rf ? w1n : w1;
                wn = rf ? nx : wn;
                rp += rf ? 1u : 0u;
                off -= rf ? 32u : 0u;
            }
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(trow_a + (uint32_t)((q ^ col8) << 4)), "r"(ow[0]), "r"(ow[1]),
                         "r"(ow[2]), "r"(ow[3])
                         : "memory");
        }
        // ---- the groups after: finish the bookkeeping loaded a pass ago, request its stream, load the next bookkeeping ----
        const uint32_t g3 = ticket_group();
        DsItem nn = ds_finish(P, raw, K, T->S);
        raw = ds_load(P, g3, ngroups, gpc, lane);
        __syncwarp();                                             // every lane is done with this group's stream buffer and has stored its row
        uint32_t lo_nn;
        stage(nn, buf, lo_nn);                                    // (into the buffer just read)
        // ---- write-out ----
        if (__all_sync(FULL, rem == 128)) {
            // 32 complete sub-chunks of one channel: 4 KB of consecutive symbols
#if MUA_DS_TMA
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the lanes' tile stores, before the TMA engine reads them
            __syncwarp();
            if (lane == 0) {
                const uint32_t ch = g0 / gpc;
                tma_store_3d(&PS.tmap, tile_a, 0, (int)(8u * (cur.info >> 12) + (g0 - ch * gpc) * 32u), (int)ch);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else if (false) {
#endif
            uint8_t* d0 = reinterpret_cast<uint8_t*>(__shfl_sync(FULL, optr_self, 0));
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = i * 4 + row4;
                uint4 v;
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(tile_a + r * 128 + (uint32_t)((col8 ^ (r & 7)) << 4)));
                *reinterpret_cast<uint4*>(d0 + r * 128 + col8 * 16) = v;
            }
        } else {
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
                const int r = i * 4 + row4;
                const int vr = __shfl_sync(FULL, rem, r) - col8 * 16;                  // valid bytes from this lane's column on
                const unsigned long long dptr = __shfl_sync(FULL, optr_self, r) + (unsigned long long)(col8 * 16);
                if (vr > 0) {
                    uint4 v;
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(tile_a + r * 128 + (uint32_t)((col8 ^ (r & 7)) << 4)));
                    uint8_t* d = reinterpret_cast<uint8_t*>(dptr);
                    if (vr >= 16 && (dptr & 15) == 0) {
                        *reinterpret_cast<uint4*>(d) = v;
                    } else if (vr >= 16 && (dptr & 7) == 0) {
                        reinterpret_cast<uint2*>(d)[0] = make_uint2(v.x, v.y);
                        reinterpret_cast<uint2*>(d)[1] = make_uint2(v.z, v.w);
                    } else if (vr >= 16 && (dptr & 3) == 0) {
                        reinterpret_cast<uint32_t*>(d)[0] = v.x; reinterpret_cast<uint32_t*>(d)[1] = v.y;
                        reinterpret_cast<uint32_t*>(d)[2] = v.z; reinterpret_cast<uint32_t*>(d)[3] = v.w;
                    } else {
                        const int nbyte = min(16, vr);
                        for (int kk = 0; kk < nbyte; ++kk) {
                            const uint32_t w = kk < 8 ? (kk < 4 ? v.x : v.y) : (kk < 12 ? v.z : v.w);
                            d[kk] = (uint8_t)(w >> (8 * (kk & 3)));
                        }
                    }
                }
            }
        }
        __syncwarp();                                             // the tile is free
        cur = nxt; lo_cur = lo_nxt;
        nxt = nn; lo_nxt = lo_nn;
        buf ^= 1;
        g0 = g1; g1 = g2; g2 = g3;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#if MUA_DS_TMA
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#endif
    dec_wait_report(P);
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

// uint8 raster, C % 4 == 0: thread = four channels of one bin (4-byte coalesced loads along channels, four rows in
// flight), sums in 16-bit SWAR lanes, widened every 128 rows (128 x 255 < 2^15), two 16-byte stores of int64 counts.
__global__ void __launch_bounds__(256) k_bin_counts_u8x4(const uint8_t* __restrict__ raster, int64_t T0, int C, int r, int64_t nb,
                                                         int64_t* __restrict__ counts) {
    const int c = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (c >= C) return;
    for (int64_t b = blockIdx.y; b < nb; b += gridDim.y) {
        const int64_t r0 = b * r, r1 = min(r0 + (int64_t)r, T0);
        unsigned long long tot[4] = {0, 0, 0, 0};
        const uint8_t* p = raster + r0 * C + c;
        for (int64_t tb = r0; tb < r1; tb += 128) {
            const int64_t te = min(tb + 128, r1);
            uint32_t ev = 0, od = 0;
            int64_t t = tb;
            for (; t + 4 <= te; t += 4, p += 4 * (int64_t)C) {
                uint32_t v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint32_t*>(p + u * (int64_t)C);
#pragma unroll
                for (int u = 0; u < 4; ++u) { ev += v[u] & 0x00FF00FFu; od += (v[u] >> 8) & 0x00FF00FFu; }
            }
            for (; t < te; ++t, p += C) {
                const uint32_t v = *reinterpret_cast<const uint32_t*>(p);
                ev += v & 0x00FF00FFu; od += (v >> 8) & 0x00FF00FFu;
            }
            tot[0] += ev & 0xFFFFu; tot[1] += od & 0xFFFFu; tot[2] += ev >> 16; tot[3] += od >> 16;
        }
        ulonglong2* o = reinterpret_cast<ulonglong2*>(counts + b * C + c);
        o[0] = make_ulonglong2(tot[0], tot[1]);
        o[1] = make_ulonglong2(tot[2], tot[3]);
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

// Wide variant for C % 8 == 0 and bin_res <= 128 (every bin period of the scripts on a 1 ms raster): tile = 256
// channels x 64 bins.  A warp sums one bin at a time: 8-byte coalesced loads along channels (lane = 8 channels),
// four rows in flight, SWAR accumulation in 16-bit lanes (even / odd bytes of a word: 2 LOP + 1 SHF + 2 ADD per
// 4 raster bytes; 128 rows x 255 < 2^15 keeps the saturation compare carry-free), SWAR min(x, sat), even | odd << 8
// = four saturated bytes in channel order, one conflict-free 8-byte store into the [bin][channel] tile.  The
// transpose happens on the read side: thread = channel, 64 byte loads down its column (a warp reads 32 consecutive
// channels of one tile row: 8 words, broadcast, no conflicts), packed into four 16-byte stores along bins.
constexpr int BW_TC = 256, BW_TB = 64;
__global__ void __launch_bounds__(256) k_bin_sym_wide(const uint8_t* __restrict__ raster, int64_t T0, int C, int r, int64_t nb,
                                                      uint8_t* __restrict__ sym, int64_t stride, int sat) {
    __shared__ __align__(16) uint8_t tile[BW_TB][BW_TC];
    const int c0 = blockIdx.x * BW_TC;
    const int64_t b0 = (int64_t)blockIdx.y * BW_TB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = c0 + 8 * lane;
    const uint32_t satv = (uint32_t)sat * 0x00010001u, satk = (uint32_t)(0x7FFF - sat) * 0x00010001u;
    for (int i = 0; i < BW_TB / 8; ++i) {
        const int bl = warp + 8 * i;
        const int64_t b = b0 + bl;
        uint32_t ev[2] = {0, 0}, od[2] = {0, 0};
        if (b < nb && c < C) {
            const int64_t r0 = b * r, r1 = min(r0 + (int64_t)r, T0);
            const uint8_t* p = raster + r0 * C + c;
            int64_t t = r0;
            for (; t + 4 <= r1; t += 4, p += 4 * (int64_t)C) {
                uint2 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint2*>(p + u * (int64_t)C);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    ev[0] += v[u].x & 0x00FF00FFu; od[0] += (v[u].x >> 8) & 0x00FF00FFu;
                    ev[1] += v[u].y & 0x00FF00FFu; od[1] += (v[u].y >> 8) & 0x00FF00FFu;
                }
            }
            for (; t < r1; ++t, p += C) {
                const uint2 v = *reinterpret_cast<const uint2*>(p);
                ev[0] += v.x & 0x00FF00FFu; od[0] += (v.x >> 8) & 0x00FF00FFu;
                ev[1] += v.y & 0x00FF00FFu; od[1] += (v.y >> 8) & 0x00FF00FFu;
            }
        }
        uint32_t w[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            // 16-bit lanes: bit 15 of x + (0x7FFF - sat) is set  <=>  x > sat
            const uint32_t me = (((ev[j] + satk) >> 15) & 0x00010001u) * 0xFFFFu;
            const uint32_t mo = (((od[j] + satk) >> 15) & 0x00010001u) * 0xFFFFu;
            const uint32_t e = (ev[j] & ~me) | (satv & me), o = (od[j] & ~mo) | (satv & mo);
            w[j] = e | (o << 8);
        }
        *reinterpret_cast<uint2*>(&tile[bl][8 * lane]) = make_uint2(w[0], w[1]);
    }
    __syncthreads();
    const int cc = c0 + threadIdx.x;
    if (cc >= C) return;
    uint8_t* dst = sym + (int64_t)cc * stride + b0;
    const bool vec = (stride % 16 == 0) && b0 + BW_TB <= nb && ((reinterpret_cast<uintptr_t>(sym) & 15) == 0);
#pragma unroll
    for (int q = 0; q < BW_TB / 16; ++q) {
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = 16 * q + 4 * j;
            o[j] = (uint32_t)tile[k][threadIdx.x] | ((uint32_t)tile[k + 1][threadIdx.x] << 8) |
                   ((uint32_t)tile[k + 2][threadIdx.x] << 16) | ((uint32_t)tile[k + 3][threadIdx.x] << 24);
        }
        if (vec) {
            reinterpret_cast<uint4*>(dst)[q] = make_uint4(o[0], o[1], o[2], o[3]);
        } else {
            for (int k = 0; k < 16 && b0 + 16 * q + k < nb; ++k) dst[16 * q + k] = (uint8_t)(o[k >> 2] >> (8 * (k & 3)));
        }
    }
}

// Events -> saturated bin counts, one thread per event (grid stride).  The bin is found by division and then
// corrected against the float64 edges t0 + k*w (separately rounded multiply and add, as NumPy / MATLAB compute
// them), so the result equals a histogram over those edges bit for bit.  Counts live in the output bytes themselves:
// a saturating byte increment by compare-and-swap on the containing word (a byte never exceeds `sat`, so nothing
// carries into its neighbours) -- no workspace, and at MUA rates (a few events per bin) hardly any contention.
__global__ void __launch_bounds__(256) k_bin_events(const double* __restrict__ times, const int32_t* __restrict__ chan, int64_t N,
                                                    double t0, double w, int64_t nb, int C, uint8_t* __restrict__ sym,
                                                    int64_t stride, uint32_t sat) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = times[i];
        const int c = chan[i];
        if (c < 0 || c >= C || !(x >= t0)) continue;                     // also drops NaN
        auto edge = [&](int64_t k) { return __dadd_rn(t0, __dmul_rn((double)k, w)); };
        if (x > edge(nb)) continue;
        double q = floor((x - t0) / w);
        int64_t k = q < 0.0 ? 0 : (q > (double)nb ? nb : (int64_t)q);
        while (k > 0 && x < edge(k)) --k;
        while (k < nb && x >= edge(k + 1)) ++k;
        if (k == nb) k = nb - 1;                                         // x == last edge: the last bin is closed
        uint8_t* bp = sym + (int64_t)c * stride + k;
        uint32_t* wp = reinterpret_cast<uint32_t*>(reinterpret_cast<uintptr_t>(bp) & ~(uintptr_t)3);
        const int sh = 8 * (int)(reinterpret_cast<uintptr_t>(bp) & 3);
        uint32_t old = *wp;
        while (((old >> sh) & 0xFFu) < sat) {
            const uint32_t seen = atomicCAS(wp, old, old + (1u << sh));
            if (seen == old) break;
            old = seen;
        }
    }
}

}  // namespace mua
