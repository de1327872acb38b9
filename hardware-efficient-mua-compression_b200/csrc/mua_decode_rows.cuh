// Stage 6 for recordings of many SHORT rows, codebooks of the variable-count decoder (S > 3 or Lmax > 2): a LANE per channel.
//
// k_decode_var gives a lane one 1024-symbol chunk; a 2 400-bin row has a 960- and a 240-symbol chunk in its window, the warp
// waits for its longest lane and every group pays the staging and write-out bookkeeping: 0.05-0.1 of the HBM roofline.  With
// 100k channels there is a dependent lookup chain per CHANNEL to be had -- as many as the chip can hold -- so here every lane
// decodes its channel's whole stream from bit 0 (the chunk side info is not needed):
//   * same tables as k_decode_var: per codebook row a 2^Wv-entry rank table in shared memory (bit-reversed index, rows skewed),
//     a lookup returns up to four symbols, the lane's peak is applied by the PRMT that unpacks the entry, windows that start
//     with the all-zero entry's bits are answered from a register;
//   * the stream reaches the lane through a lane-private 64-byte ring in shared memory, topped up with one 16-byte cp.async per
//     lane every eight lookups (<= 96 bits are consumed in between; the copy issued at one top-up is awaited at the next, ~1000
//     cycles later).  A first version loaded the stream words straight into registers two words ahead: the 32 lanes of a warp
//     refill at different lookups, so nearly every lookup stalled the whole warp on some lane's global load (8x slower);
//     loads never leave the slot;
//   * symbols queue in a 64-bit register, leave as 4-byte words for a lane-private column of a 4-word ring and as 16-byte stores
//     for the lane's row of the decoded buffer; the units at the two ends of the window are written byte by byte, so nothing
//     outside [start, end) is touched.
// 3 KB of shared memory per warp: 32 warps per SM fit beside the tables.  Results are those of k_decode_var (same tests, both families).
// The loop runs ~16 instructions per symbol (a lookup's bookkeeping is per lane here, not shared by a warp), twice k_decode_var's:
// the dispatch takes this kernel only for windows of at most two chunks, where k_decode_var's lanes mostly idle.
// (The same scheme for the chosen system's codebook class -- four symbols per 8-bit lookup, one 16-byte store per four lookups, no
// symbol queue: bit-exact, but 34 instructions per lookup at 29 % issue utilisation, 0.119 ms against the 0.061 ms of k_decode_lane
// on 100k x 2 400 and 0.58 against 0.18 ms on 100k x 12 000 -- was measured and removed: k_decode_lane's lookup chain is 16
// branch-free instructions on conflict-free lane-private tables.)
#pragma once
#include "mua_decode.cuh"
#include "mua_encode_rows.cuh"

namespace mua {

constexpr int DR_WARPS = 32;
constexpr int DR_PER_WARP = 512 + 32 * 80;

template <bool WIDE>
__global__ void __launch_bounds__(DR_WARPS * 32, 1) k_decode_rows(const __grid_constant__ DecParams P, int wuse) {
    extern __shared__ __align__(1024) uint8_t dsm[];
    const TabHdr* T = reinterpret_cast<const TabHdr*>(P.tab);
    const int K = T->K, Wv = T->Wv;
    if (T->S != P.S || K != P.K || T->Lmax != P.Lmax || T->decv_off == 0 || (WIDE != (T->S > 8))) {
        if (threadIdx.x == 0) dec_flag(P.status, MUA_DEC_BAD_TABLE);   // host view does not match the table block
        return;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // shared memory: [DR_WARPS x (output ring: 4 words x 32 lanes; stream rings: 32 lanes x 80 B)][rank -> symbol maps][tables]
    const uint32_t ring_lane = smem_u32(dsm) + warp * DR_PER_WARP + lane * 4;
    const uint32_t str_lane = smem_u32(dsm) + warp * DR_PER_WARP + 512 + lane * 80;  // 16 words of the lane's stream (80-byte rows: 16 B x odd)
    uint32_t* s_map = reinterpret_cast<uint32_t*>(dsm + DR_WARPS * DR_PER_WARP);     // uint4 [MUA_MAX_S]: idx[p][0..15]
    uint32_t* s_tab = s_map + 4 * MUA_MAX_S;
    const int row_words = (1 << Wv) + DV_ROW_SKEW;
    {
        const uint32_t* g = reinterpret_cast<const uint32_t*>(P.tab + T->decv_off);
        for (int i = threadIdx.x; i < (K << Wv); i += blockDim.x) {                  // layout of k_decode_var
            const int k = i >> Wv, v = i & ((1 << Wv) - 1);
            s_tab[k * row_words + (int)(__brev((uint32_t)v) >> (32 - Wv))] = g[i];
        }
        const uint32_t* gi = reinterpret_cast<const uint32_t*>(&T->idx[0][0]);
        for (int i = threadIdx.x; i < 4 * MUA_MAX_S; i += blockDim.x) s_map[i] = gi[i];
    }
    __syncthreads();
    const uint32_t wmask = (1u << Wv) - 1u;
    const uint32_t nwords = (uint32_t)min((long long)(P.slot_bytes >> 2), 0x7FFFFFFFll);
    const int nblk = (P.C + 31) >> 5;

    if (warp < wuse)
    for (int blk = blockIdx.x + gridDim.x * warp; blk < nblk; blk += gridDim.x * wuse) {
        const int c = blk * 32 + lane;
        int rem = 0, start = 0, end = 0;
        const uint32_t* sp = reinterpret_cast<const uint32_t*>(P.stream);
        uint8_t* row = P.dec;
        uint32_t tab_sa = smem_u32(s_tab), m0 = 0, m1 = 0, m2 = 0, m3 = 0;
        if (c < P.C) {
            start = P.start[c];
            end = P.end[c];
            const int pk = P.peak[c], en = P.enc[c];
            if (end > start && start >= 0) {
                if (pk >= T->S || en >= K) {
                    dec_flag(P.status, MUA_DEC_BAD_TABLE);
                } else {
                    rem = end - start;
                    sp = reinterpret_cast<const uint32_t*>(P.stream + (size_t)c * P.slot_bytes);
                    row = P.dec + (P.off ? P.off[c] : (int64_t)c * P.stride);
                    tab_sa += (uint32_t)(en * row_words) * 4u;
                    const uint4 mp = reinterpret_cast<const uint4*>(s_map)[pk];
                    m0 = mp.x; m1 = mp.y; m2 = mp.z; m3 = mp.w;
                }
            }
        }
        if (rem == 0) { start = 0; end = 0; }
        uint32_t ez;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ez) : "r"(tab_sa));        // entry of the all-zero window
        const uint32_t zm = (ez & 0x40000u) ? ((1u << ((ez >> 20) & 0xFu)) - 1u) : wmask;
        // stream: words [nw, hi) of the slot sit in the lane's ring (word i at slot i & 15); w0 / w1 hold the bits being decoded
        // (stream bit j at register bit j), w2 the next word, raw
        uint32_t hi = 0;
        if (rem > 0) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (hi < nwords) { cp_async16(str_lane + hi * 4, sp + hi); hi += 4; }
        }
        cp_async_commit();
        cp_async_wait<0>();
        uint32_t w0 = stream_rev(lds_u32(str_lane)), w1 = stream_rev(lds_u32(str_lane + 4)), w2 = lds_u32(str_lane + 8);
        uint32_t nw = 3, off = 0;
        // output: bytes [wpos, wpos + qn) of the row sit in the queue (the first word starts at start & ~3: its bytes before the
        // window are never stored); units [fl, fl + 16) still have to leave
        unsigned long long q64 = 0;
        uint32_t qn = (uint32_t)start & 3u, wpos = (uint32_t)start & ~3u, fl = (uint32_t)start & ~15u;
        auto flush_unit = [&](uint32_t upos) {   // 16 bytes of the row at upos (16-byte aligned), only what lies inside the window
            uint4 v;
            v.x = lds_u32(ring_lane); v.y = lds_u32(ring_lane + 128); v.z = lds_u32(ring_lane + 256); v.w = lds_u32(ring_lane + 384);
            if ((int)upos >= start && (int)upos + 16 <= end) {
                *reinterpret_cast<uint4*>(row + upos) = v;
            } else {
                const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int b = 0; b < 16; ++b) {
                    const int p = (int)upos + b;
                    if (p >= start && p < end) row[p] = (uint8_t)(wv[b >> 2] >> (8 * (b & 3)));
                }
            }
        };
        while (__any_sync(FULL, rem > 0)) {
            // top-up: one 16-byte unit per lane whose ring has room; the copy is awaited at the NEXT top-up
            if (rem > 0 && hi - nw < 12u && hi < nwords) { cp_async16(str_lane + (hi & 15u) * 4, sp + hi); hi += 4; }
            cp_async_commit();
            cp_async_wait<1>();
#pragma unroll 2
            for (int it = 0; it < 8; ++it) {
                const uint32_t x = __funnelshift_r(w0, w1, off);                 // next 32 stream bits, first one at bit 0
                const uint32_t e = dv_lookup(tab_sa, x, wmask, zm, ez);
                uint32_t syms = dv_syms<WIDE>(e, m0, m1, m2, m3);
                uint32_t cnt = (e >> 16) & 7u, used = (e >> 20) & 0xFu;
                if ((int)cnt > rem) {                                            // past the window end (or a finished lane)
                    cnt = (uint32_t)rem;
                    syms &= (1u << (8u * cnt)) - 1u;                             // cnt < 4 here
                    if (rem == 0) used = 0;
                }
                q64 |= (unsigned long long)syms << (8u * qn);
                qn += cnt;
                rem -= (int)cnt;
                off += used;
                if (off >= 32u) {                                                // used <= 12: at most one word per lookup
                    w0 = w1; w1 = stream_rev(w2); w2 = lds_u32(str_lane + (nw & 15u) * 4); ++nw; off -= 32u;
                }
                if (qn >= 4u) {
                    sts_u32(ring_lane + ((wpos >> 2) & 3u) * 128u, (uint32_t)q64);
                    q64 >>= 32;
                    qn -= 4u;
                    wpos += 4u;
                    if ((wpos & 15u) == 0u) { flush_unit(fl); fl += 16u; }
                }
            }
        }
        cp_async_wait<0>();
        if (end > start) {
            if (qn > 0u) {                                                       // the last, partial word
                sts_u32(ring_lane + ((wpos >> 2) & 3u) * 128u, (uint32_t)q64);
                wpos += 4u;
            }
            if (fl < wpos) flush_unit(fl);
            if (32ull * (nw - 3u) + off > 8ull * (unsigned long long)P.slot_bytes + 12ull) dec_flag(P.status, MUA_DEC_BAD_OFFSET);   // ran past the slot (the last lookup may count <= 12 bits of symbols past the window)
        }
        __syncwarp();
    }
    dec_wait_report(P);
}

}  // namespace mua
