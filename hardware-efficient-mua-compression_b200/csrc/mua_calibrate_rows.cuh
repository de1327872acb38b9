// Stages 2-4 for recordings of many SHORT rows: a LANE per channel (the calibrate companion of k_encode_rows).
//
// k_calibrate (a warp per channel) serves all history lengths of a channel from one scan, but on a 2 400-bin row nearly every
// 512-byte tile holds a boundary and takes the warp-wide snapshot path (REDUX + lane scan + shuffles): ~1 350 (S = 3) to
// ~3 300 (S = 9) warp instructions per channel, 0.1-0.2 of the HBM roofline.  Here a warp owns 32 channels of a fixed-stride
// recording (all rows T bins long, so every cutoff and window end is the SAME bin for all lanes):
//   * rows are staged by the TMA engine exactly as in k_encode_rows: boxes of 128 bins x 32 channels, 128-byte swizzle, two
//     stages per warp, mbarrier completion; lane l reads its own row conflict-free;
//   * every lane keeps the cumulative counts #{x >= v} (v = 1..S-1) of ITS channel in registers: per 4 bytes and threshold
//     one add, one LOP3 and one DP4A (the SWAR threshold compare of k_calibrate), no cross-lane traffic at all;
//   * the boundaries are a host-sorted list of warp-uniform events.  A boundary inside a 64-bin step costs one masked
//     recount of that step (same three instructions per word, byte mask from the uniform datapath); every boundary
//     stores the lane's counts as uint16 in its own shared-memory column;
//   * after the scan the warp's 32 x nH (channel, history length) pairs are finished one per lane in OUTPUT order, so every
//     [C][nH] result array is written with fully coalesced stores (a lane per channel would scatter 4-byte stores nH elements
//     apart): histograms as count differences, first argmax, SCLV costs with the rank map applied by a data-dependent PRMT on
//     the 16-byte length row (one PRMT + one IMAD per row and symbol), first argmin, bit count -- the epilogue of k_calibrate.
// Results are those of k_calibrate (same oracle, same tests; every GPU test runs with both kernel families).
#pragma once
#include <cuda.h>

#include "mua_calibrate.cuh"
#include "mua_encode_rows.cuh"

namespace mua {

constexpr int CR_MAX_WARPS = 24;
// warps per CTA: the kept counts of S > 6 leave room for fewer than 22 warps (what 100k channels ask for) anyway, and 512 threads
// may use 128 registers
__host__ __device__ constexpr int cr_max_warps(int S) { return S <= 6 ? CR_MAX_WARPS : 16; }
// bins per TMA box: 128 (two steps, 128-byte swizzle) while the scan is memory-bound; S > 3 count more thresholds per byte and
// keep more counts per boundary: half-size boxes (one step, 64-byte swizzle) leave shared memory for twice the warps
#ifndef MUA_CR_BOX3
#define MUA_CR_BOX3 128
#endif
#ifndef MUA_CR_NST3
#define MUA_CR_NST3 2
#endif
__host__ __device__ constexpr int cr_box(int S) { return S <= 3 ? MUA_CR_BOX3 : 64; }
__host__ __device__ constexpr int cr_nst(int S) { return S <= 3 ? MUA_CR_NST3 : 2; }   // stages per warp

struct CalRowsParams {
    CalibParams C;
    int32_t wuse;                        // warps of a CTA that take blocks
    uint32_t zero;                       // == 0, opaque to the compiler (see k_encode_rows: ties a TMA request to the stage's last reads)
    int32_t nev;                         // boundaries (cutoffs and window ends of all history lengths), sorted by position
    int32_t ev_pos[2 * MUA_MAX_H];
    int32_t cutv[MUA_MAX_H], endv[MUA_MAX_H];       // what d_cutoff / d_end receive
    int32_t ev_cut[MUA_MAX_H], ev_end[MUA_MAX_H];   // boundary index of the cutoff / the window end (-1: no post window)
    int32_t warps;                       // warps per CTA; shared memory: [warps x 2 stages][warps x counts][tables]; a stage = one box
    int32_t snap_bytes;                  // counts kept per warp: uint16 [nev][S - 1][32 lanes]
    alignas(64) CUtensorMap tmap;
};

struct CalRowsSmem {                     // after the stages and the kept counts of all warps:
    static constexpr int LEN = 0;                        // MUA_MAX_K x 16 B SCLV rows
    static constexpr int RANK = LEN + MUA_MAX_K * 16;    // MUA_MAX_S x 16 B rank maps
    static constexpr int SEL = RANK + MUA_MAX_S * 16;    // MUA_MAX_S x 16 B: per peak the symbols of ranks 4g..4g+3 as nibbles (PRMT selectors), g = 0..2
    static constexpr int HINFO = SEL + MUA_MAX_S * 16;   // int32 [4][MUA_MAX_H]: cutv, endv, ev_cut, ev_end
    static constexpr int BARS = HINFO + 4 * MUA_MAX_H * 4;   // one mbarrier per warp and stage
    static constexpr int tail(int nst) { return BARS + CR_MAX_WARPS * 8 * nst; }
};

// counts of one step's 64 bytes (four 16-byte pieces in qv), bytes >= `lim` masked away: acc[v] += 0x80 per byte >= v
// (`zz` == 0, opaque and different per boundary: without it the compiler hoists the masked recount's compare-adds, which do not
// depend on the boundary, out of the boundary loop into EVERY step -- 90 of 231 instructions per step at S = 3)
template <int S, bool MASKED>
__device__ __forceinline__ void cr_count(const uint4 (&qv)[4], int lim, uint32_t (&acc)[S], uint32_t zz = 0u) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (MASKED && lim <= 16 * k) break;                                     // pieces past the boundary (uniform branch)
        const uint32_t w[4] = {qv[k].x ^ zz, qv[k].y ^ zz, qv[k].z ^ zz, qv[k].w ^ zz};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t bm = 0x80808080u;
            if (MASKED) {
                const int r = min(max(lim - (16 * k + 4 * j), 0), 4);          // bytes of this word before the boundary (uniform)
                bm = r == 4 ? 0x80808080u : (0x80808080u & ((1u << (8 * r)) - 1u));
            }
            const uint32_t lo7 = w[j] & 0x7F7F7F7Fu;
#pragma unroll
            for (int v = 1; v < S; ++v) {
                const uint32_t t = lo7 + (uint32_t)(0x80 - v) * 0x01010101u;
                const uint32_t m = (t | w[j]) & bm;
                asm("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(acc[v]) : "r"(m), "r"(0x01010101u));
            }
        }
    }
}

// finish (channel c, history length h): gc / ge = #{x >= v} before the cutoff / the window end (index v = 1..S-1)
template <int S>
__device__ __forceinline__ void cr_finish(const CalibParams& P, const CalOut& O, const uint8_t* s_len, const uint8_t* s_rank, const uint8_t* s_sel, int K, int c, int h,
                                          int cut, int end_out, int npost, bool has_post, const int (&gc)[S], const int (&ge)[S]) {
    int hist[S], post[S];
    {
        int g_prev = cut, p_prev = has_post ? npost : 0;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int g_next = s + 1 < S ? gc[s + 1] : 0;
            const int p_next = (has_post && s + 1 < S) ? ge[s + 1] - gc[s + 1] : 0;
            hist[s] = g_prev - g_next;
            post[s] = p_prev - p_next;
            g_prev = g_next;
            p_prev = p_next;
        }
    }
    int p = 0;
    if (P.use_sort) {   // np.argmax: lowest index on ties (functions_1.py:77)
        int best = hist[0];
#pragma unroll
        for (int s = 1; s < S; ++s)
            if (hist[s] > best) { best = hist[s]; p = s; }
    }
    const uint4 rk4 = *reinterpret_cast<const uint4*>(s_rank + 16 * p);   // rank[p][s], s = 0..15
    uint32_t rk[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const uint32_t wsel = s < 4 ? rk4.x : (s < 8 ? rk4.y : rk4.z);
        rk[s] = (wsel >> (8 * (s & 3))) & 0xFFu;
    }
    // byte rk[s] of a 16-byte length row (ranks < 8 live in the first two words: one data-dependent PRMT)
    auto len_of = [&](const uint4& row, int s) -> uint32_t {
        if (S <= 8) return __byte_perm(row.x, row.y, rk[s]) & 0xFFu;
        return (rk[s] < 8 ? __byte_perm(row.x, row.y, rk[s]) : __byte_perm(row.z, row.w, rk[s] - 8)) & 0xFFu;
    };
    // SCLV costs: sum_r len[k][r] * am[r] with am = the histogram in RANK order.  The counts (< 2^16: short rows) are split into
    // a low-byte and a high-byte plane, four symbols per word; one PRMT per word and plane (selectors = the symbols of ranks
    // 4g..4g+3, TabHdr::idx[p], packed as nibbles) brings a plane into rank order; a cost is then 2 * ceil(S / 4) DP4As against
    // the row's packed lengths instead of S data-dependent byte extractions and S multiply-adds
    constexpr int NG = (S + 3) / 4;
    uint32_t hl[3] = {0u, 0u, 0u}, hh[3] = {0u, 0u, 0u};
#pragma unroll
    for (int s = 0; s < S; ++s) {
        hl[s >> 2] |= ((uint32_t)hist[s] & 0xFFu) << (8 * (s & 3));
        hh[s >> 2] |= (((uint32_t)hist[s] >> 8) & 0xFFu) << (8 * (s & 3));
    }
    const uint4 sel4 = *reinterpret_cast<const uint4*>(s_sel + 16 * p);
    const uint32_t selw[3] = {sel4.x, sel4.y, sel4.z};
    uint32_t lo[NG], hi[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        if (S <= 8) {
            lo[g] = __byte_perm(hl[0], hl[1], selw[g]);
            hi[g] = __byte_perm(hh[0], hh[1], selw[g]);
        } else {   // symbols 8, 9 live in the third word: two PRMTs and a select by the selectors' bit 3
            const uint32_t e7 = selw[g] & 0x7777u;
            const uint32_t msk = __byte_perm(0x0000FF00u, 0u, (selw[g] >> 3) & 0x1111u);
            lo[g] = (__byte_perm(hl[0], hl[1], e7) & ~msk) | (__byte_perm(hl[2], 0u, e7) & msk);
            hi[g] = (__byte_perm(hh[0], hh[1], e7) & ~msk) | (__byte_perm(hh[2], 0u, e7) & msk);
        }
    }
    uint32_t best_cost = 0;
    int enc = -1;
    for (int k = 0; k < K; ++k) {   // np.argmin: first minimum (get_BR_no_sort.py:236); rows short: costs fit 32 bits
        if (!((O.active >> k) & 1ull)) continue;
        const uint4 row = *reinterpret_cast<const uint4*>(s_len + 16 * k);
        const uint32_t rw[3] = {row.x, row.y, row.z};
        uint32_t clo = 0, chi = 0;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            asm("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(clo) : "r"(rw[g]), "r"(lo[g]));
            asm("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(chi) : "r"(rw[g]), "r"(hi[g]));
        }
        const uint32_t cost = clo + (chi << 8);
        if (enc < 0 || cost < best_cost) { best_cost = cost; enc = k; }
    }
    if (enc < 0) enc = 0;
    const uint4 erow = *reinterpret_cast<const uint4*>(s_len + 16 * enc);
    uint32_t bits = 0;
#pragma unroll
    for (int s = 0; s < S; ++s) bits += (uint32_t)post[s] * len_of(erow, s);
    const size_t o = (size_t)c * P.nH + h;
    if (O.cutoff) O.cutoff[o] = cut;
    if (O.end) O.end[o] = end_out;
    if (O.peak) O.peak[o] = (uint8_t)p;
    if (O.enc) O.enc[o] = (uint8_t)enc;
    if (O.bits) O.bits[o] = (long long)bits;
    if (O.nsym) O.nsym[o] = has_post ? (long long)npost : 0ll;
    if (O.assign_m) {
#pragma unroll
        for (int s = 0; s < S; ++s) O.assign_m[o * S + rk[s]] = hist[s];
    }
    if (O.post_m) {
#pragma unroll
        for (int s = 0; s < S; ++s) O.post_m[o * S + rk[s]] = post[s];
    }
}

template <int S>
__global__ void __launch_bounds__(cr_max_warps(S) * 32, 1) k_calibrate_rows(const __grid_constant__ CalRowsParams PR) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    using SM = CalRowsSmem;
    const CalibParams& P = PR.C;
    const CalOut& O = P.out[0];
    constexpr int BOXW = cr_box(S), SPB = BOXW / ER_TILE, STAGE = BOXW * 32, NST = cr_nst(S);   // steps per box, bytes per stage, stages
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* tail = smem_raw + (size_t)PR.warps * (NST * STAGE + PR.snap_bytes);
    {
        const uint4* gl = reinterpret_cast<const uint4*>(&O.tab->lens[0][0]);
        const uint4* gr = reinterpret_cast<const uint4*>(&O.tab->rank[0][0]);
        for (int i = threadIdx.x; i < MUA_MAX_K; i += blockDim.x) reinterpret_cast<uint4*>(tail + SM::LEN)[i] = gl[i];
        if (threadIdx.x < MUA_MAX_S) reinterpret_cast<uint4*>(tail + SM::RANK)[threadIdx.x] = gr[threadIdx.x];
        if (threadIdx.x < MUA_MAX_S) {
            const uint8_t* ix = &O.tab->idx[threadIdx.x][0];
            uint32_t w[4] = {0u, 0u, 0u, 0u};
            for (int r = 0; r < 12; ++r) w[r >> 2] |= (uint32_t)(ix[r] & 0xFu) << (4 * (r & 3));
            reinterpret_cast<uint4*>(tail + SM::SEL)[threadIdx.x] = make_uint4(w[0], w[1], w[2], 0u);
        }
        if (threadIdx.x < MUA_MAX_H) {
            int32_t* hi = reinterpret_cast<int32_t*>(tail + SM::HINFO);
            hi[threadIdx.x] = PR.cutv[threadIdx.x];
            hi[MUA_MAX_H + threadIdx.x] = PR.endv[threadIdx.x];
            hi[2 * MUA_MAX_H + threadIdx.x] = PR.ev_cut[threadIdx.x];
            hi[3 * MUA_MAX_H + threadIdx.x] = PR.ev_end[threadIdx.x];
        }
    }
    const uint32_t in0 = smem_u32(smem_raw) + warp * (NST * STAGE);
    // the lane's row of stage 0 with the swizzle of its pieces folded in: 128-byte rows, piece k at 16 (k ^ (l & 7)); 64-byte rows,
    // piece k at 16 (k ^ ((l >> 1) & 3))
    const uint32_t in_lane = BOXW == 128 ? (in0 + lane * 128) | ((lane & 7) * 16) : (in0 + lane * 64) | (((lane >> 1) & 3) * 16);
    uint16_t* snap = reinterpret_cast<uint16_t*>(smem_raw + (size_t)PR.warps * (NST * STAGE) + (size_t)warp * PR.snap_bytes);   // [e][v - 1][32 lanes]
    const uint32_t bar0 = smem_u32(tail + SM::BARS) + warp * (8 * NST);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NST; ++i) mbar_init(reinterpret_cast<uint64_t*>(tail + SM::BARS) + NST * warp + i, 1);
        fence_barrier_init();
    }
    __syncthreads();
    const uint8_t* s_len = tail + SM::LEN;
    const uint8_t* s_rank = tail + SM::RANK;
    const uint8_t* s_sel = tail + SM::SEL;
    const int32_t* s_hinfo = reinterpret_cast<const int32_t*>(tail + SM::HINFO);
    const int K = O.tab->K;
    const int nH = P.nH;
    const int nblk = (P.L.C + 31) >> 5;
    const int nev = PR.nev;
    const int scan_end = nev > 0 ? PR.ev_pos[nev - 1] : 0;
    const int nt = (scan_end + ER_TILE - 1) / ER_TILE, nbox = (nt + SPB - 1) / SPB;
    uint32_t phase = 0;
    if (warp >= PR.wuse) return;

    for (int blk = blockIdx.x + gridDim.x * warp; blk < nblk; blk += gridDim.x * PR.wuse) {
        auto issue_box = [&](int tt, uint32_t s, uint32_t dep) {
            mbar_expect_tx_s(bar0 + 8 * s, STAGE);
            tma_load_2d(in0 + s * STAGE, &PR.tmap, BOXW * tt + (int)dep, blk * 32, bar0 + 8 * s);
        };
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < NST; ++i)
                if (i < nbox) issue_box(i, i, 0u);
        }
        // ---- scan: the lane's cumulative counts, kept at every boundary ----
        uint32_t acc[S];
#pragma unroll
        for (int v = 0; v < S; ++v) acc[v] = 0;
        int ev = 0;
        auto keep = [&](int e, const uint32_t (&g)[S]) {
#pragma unroll
            for (int v = 1; v < S; ++v) snap[(e * (S - 1) + (v - 1)) * 32 + lane] = (uint16_t)(g[v] >> 7);
        };
        int ts = 0;
        for (int t = 0; t < nt; ++t, ts += ER_TILE) {
            const int bb = t / SPB, sub = t % SPB;   // box and step in the box
            const uint32_t s = (uint32_t)bb % NST;
            if (sub == 0) {
                mbar_wait_s(bar0 + 8 * s, (phase >> s) & 1u);
                phase ^= 1u << s;
            }
            const uint32_t tile = (in_lane + s * STAGE) ^ (sub * 64u);
            uint4 qv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) qv[k] = lds_u128(tile ^ (16u * k));
            if (sub == SPB - 1 && bb + NST < nbox) {   // the box is in registers: its stage takes the box NST boxes on
                const uint32_t dep = (qv[0].x | qv[1].x | qv[2].x | qv[3].x) & PR.zero;   // the four loads have been performed
                __syncwarp();
                if (lane == 0) issue_box(bb + NST, s, dep);
            }
            while (ev < nev && PR.ev_pos[ev] < ts + ER_TILE) {   // boundaries inside the step: masked recount
                uint32_t g[S];
#pragma unroll
                for (int v = 0; v < S; ++v) g[v] = acc[v];
                cr_count<S, true>(qv, PR.ev_pos[ev] - ts, g, PR.zero & (uint32_t)ev);
                keep(ev, g);
                ++ev;
            }
            cr_count<S, false>(qv, 0, acc);
            while (ev < nev && PR.ev_pos[ev] == ts + ER_TILE) {  // boundaries at the end of the step
                keep(ev, acc);
                ++ev;
            }
        }
        __syncwarp();
        // ---- (channel, history length) pairs in output order, one per lane ----
        const int npair = min(32, P.L.C - blk * 32) * nH;
        for (int it = lane; it < npair; it += 32) {
            const int cl = it / nH, h = it - cl * nH;
            const int cut = s_hinfo[h], end_out = s_hinfo[MUA_MAX_H + h], ec = s_hinfo[2 * MUA_MAX_H + h], ee = s_hinfo[3 * MUA_MAX_H + h];
            int gc[S], ge[S];
            gc[0] = ge[0] = 0;
#pragma unroll
            for (int v = 1; v < S; ++v) {
                gc[v] = snap[(ec * (S - 1) + (v - 1)) * 32 + cl];
                ge[v] = ee >= 0 ? snap[(ee * (S - 1) + (v - 1)) * 32 + cl] : 0;
            }
            cr_finish<S>(P, O, s_len, s_rank, s_sel, K, blk * 32 + cl, h, cut, end_out, ee >= 0 ? end_out - cut : 0, ee >= 0, gc, ge);
        }
        __syncwarp();
    }
}

}  // namespace mua
