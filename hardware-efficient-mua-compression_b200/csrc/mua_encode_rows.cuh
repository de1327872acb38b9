// Stage 5 for SHORT rows: a LANE per channel.
//
// The warp-per-channel encoders (mua_encode.cuh) pay ~800-1000 warp instructions of per-channel work (window bookkeeping,
// warp scan, placement between lanes, flush, side info) whatever the row length: on the recordings the reference really
// runs on (Flint / Sabes / Brochier: 600 s at 50 ms = 12 000 bins, BASELINE cfg4 at 50 ms = 2 400 bins) that is all they
// do and they reach 0.2 of the HBM roofline.  Here a warp owns 32 channels and every lane codes ITS channel alone:
//   * no warp scan, no partial words travelling between lanes, no select network: the bit writer is lane-private
//     (64-bit hold register, one predicated shared store per 16-symbol piece);
//   * fixed-stride recordings (FIXED): the recording is a 2-D tensor (bin, channel) for the TMA engine; ONE
//     cp.async.bulk.tensor box of 128 bins x 32 channels (4 KB, 128-byte swizzle, mbarrier completion, two stages) feeds a
//     warp for 128 symbols per lane.  The hardware swizzle puts piece k of row l at position k ^ (l & 7): the lanes'
//     16-byte reads of their own rows are conflict-free without padding, rows past the last channel and bins past the row
//     end are zero-filled by the engine, and the loads cost the SM no LSU wavefronts and no address arithmetic.
//     (Measured on the way: a 1-D bulk copy per lane and row is bound by the TMA unit's operation rate -- ~8 cycles per
//     operation and SM whatever its size, 2.2 TB/s for 64-byte rows; 16-byte cp.async copies spend 16-18 shared-memory
//     wavefronts per instruction because a row's 64 bytes arrive as separate 32-byte sectors; L2 prefetches ahead of the boxes --
//     one line per lane, or a whole box by cp.async.bulk.prefetch.tensor -- made the kernel 10-30 % slower.)
//   * ragged sets (per-channel offsets / lengths): 16-byte cp.async copies, 64 bytes per row and stage, four lanes per row and
//     eight rows per instruction, rows padded to 80 bytes (5 x 16 B, odd: conflict-free reads);
//   * the stream leaves the lane through a lane-private column of an 8-word ring ([word][lane]: every access conflict-free)
//     as complete 128-bit units, one 16-byte store per unit;
//   * tiles are aligned to ABSOLUTE multiples of 64 bins, so 128-symbol sub-chunk and 1024-symbol chunk boundaries (the
//     side info of the decoders) always coincide with a tile start: recording them is one predicated store;
//   * windows differ per lane: a tile that is not completely inside the window of EVERY lane is coded with the null-digit
//     table of the fast encoder (symbols outside the window code no bits), all other tiles with the plain base-S table.
// One persistent CTA of ER_WARPS warps per SM; the 32-channel blocks are dealt round-robin to the SMs, and to as many warps of
// an SM as make the rounds even (a block is a long task: a nearly empty extra round would cost as much as a full one).
// Stream bytes, side info, bit counts and flags are those of k_encode_fast (same oracle, same tests).
#pragma once
#include <cuda.h>

#include "mua_decode.cuh"
#include "mua_encode.cuh"

namespace mua {

constexpr int ER_WARPS = 24;            // warps per CTA (one CTA per SM)
constexpr int ER_TILE = 64;             // symbols per lane and step
constexpr int ER_ROWB = ER_TILE + 16;   // padded row of a cp.async stage (ragged sets)
constexpr int ER_LUT_MAX = 2304;        // bytes of all (peak, row) tables of a block (S * K * 768): S = 3 with one row, S = 2 with one

struct EncRowsSmem {
    static constexpr int LUT = 0;                                   // ER_LUT_MAX
    static constexpr int BARS = ER_LUT_MAX;                         // 2 mbarriers per warp
    static constexpr int WARP0 = 3072;                              // 1024-byte aligned (the swizzle pattern uses address bits 7..9)
    static constexpr int STAGE = 4096;                              // one TMA box: 32 rows x 128 bytes
    static constexpr int RING = 2 * STAGE;                          // 8 words x 32 lanes
    static constexpr int PER_WARP = RING + 8 * 32 * 4;              // 9216
    static constexpr int TOTAL = WARP0 + ER_WARPS * PER_WARP;
};
static_assert(ER_LUT_MAX + ER_WARPS * 16 <= EncRowsSmem::WARP0, "tables and barriers share the first 3 KB");

struct EncRowsParams {
    EncParams E;
    int32_t wuse;                       // warps of a CTA that take blocks
    uint32_t zero;                      // == 0, opaque to the compiler: ties a TMA request to the registers loaded from the stage it refills
    alignas(64) CUtensorMap tmap;       // FIXED: uint8 [C][stride] as a 2-D tensor (T bins x C channels), box 128 x 32, 128-byte swizzle
};

__device__ __forceinline__ void sts_u32(uint32_t saddr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory"); }
__device__ __forceinline__ uint4 lds_u128(uint32_t saddr) {
    uint4 q;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(saddr));
    return q;
}
__device__ __forceinline__ void cp_async16(uint32_t dst_saddr, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_saddr), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx_s(uint32_t bar_saddr, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_saddr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar_saddr, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar_saddr), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst_saddr, const CUtensorMap* map, int x, int y, uint32_t bar_saddr) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst_saddr),
                 "l"(map), "r"(x), "r"(y), "r"(bar_saddr)
                 : "memory");
}

// lane-private bit writer: `hold` keeps the stream's last bits right-aligned (the tb & 31 bits that do not fill a word yet
// are its lowest ones; whatever lies above is dead), `tb` counts the channel's bits.  A piece of <= 32 bits completes at
// most one word, which goes to slot (word index & 7) of the lane's ring column.
template <uint32_t RMASK = 7u>
__device__ __forceinline__ void er_append(unsigned long long& hold, uint32_t& tb, uint32_t code, uint32_t len, uint32_t ring_lane) {
    hold = (hold << len) | code;
    const uint32_t tbn = tb + len;
    if ((tbn ^ tb) & ~31u) sts_u32(ring_lane + ((tb >> 5) & RMASK) * 128u, __funnelshift_r((uint32_t)hold, (uint32_t)(hold >> 32), tbn));
    tb = tbn;
}

// FIXED: one row stride and one row length for all channels (no d_off / d_len): TMA tensor boxes; otherwise cp.async.
template <int SV, bool FIXED>
__global__ void __launch_bounds__(ER_WARPS * 32, 1) k_encode_rows(const __grid_constant__ EncRowsParams PR) {
    static_assert(SV == 2 || SV == 3, "the null-digit table needs (S+1)^4 <= 256 entries");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    using SM = EncRowsSmem;
    const EncParams& P = PR.E;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TabHdr* T = reinterpret_cast<const TabHdr*>(P.tab);
    const int K = T->K;
    constexpr int S = SV;
    const uint32_t lut_base = smem_u32(smem_raw + SM::LUT);
    if (T->S != SV || S != P.S || K != P.K || T->Lmax != P.Lmax || T->Lmax > 2 || T->enc4_off == 0 || (lut_base & 1023u) ||
        S * K * EF_LUT_B > ER_LUT_MAX) {
        if (threadIdx.x == 0) *P.overflow = MUA_ENC_BAD_TABLE;   // launch configuration does not match the table block
        return;
    }
    {   // the tables of every (peak, row) pair of this alphabet: lanes of one warp code channels with different pairs.  (Skewing the
        // tables against each other, as k_encode_rows_pair does, costs one LOP3 per word and gained nothing here: three tables.)
        const uint4* g_enc4 = reinterpret_cast<const uint4*>(P.tab + T->enc4_off);
        for (int i = threadIdx.x; i < S * K * (EF_LUT_B / 16); i += ER_WARPS * 32) reinterpret_cast<uint4*>(smem_raw + SM::LUT)[i] = g_enc4[i];
    }
    uint8_t* sm = smem_raw + SM::WARP0 + warp * SM::PER_WARP;
    const uint32_t in0 = smem_u32(sm);                                      // stage 0 of this warp
    // the lane's row of stage 0: 128-byte rows, piece k at (16 k) ^ (16 (lane & 7)) (FIXED) / 80-byte rows (ragged)
    const uint32_t in_lane = FIXED ? (in0 + lane * 128) | ((lane & 7) * 16) : in0 + lane * ER_ROWB;
    const uint32_t ring_lane = in0 + SM::RING + lane * 4;                   // the lane's column of the ring
    const uint32_t bar0 = smem_u32(smem_raw + SM::BARS) + warp * 16;
    if (FIXED && lane == 0) {
        mbar_init(reinterpret_cast<uint64_t*>(smem_raw + SM::BARS) + 2 * warp, 1);
        mbar_init(reinterpret_cast<uint64_t*>(smem_raw + SM::BARS) + 2 * warp + 1, 1);
        fence_barrier_init();
    }
    __syncthreads();
    const uint32_t slot_units = (uint32_t)min((long long)(P.slot_bytes >> 4), 0x7FFFFFFFll);
    const int nblk = (P.L.C + 31) >> 5;
    uint32_t phase = 0;              // FIXED: bit s = parity of the next completion of stage s

    if (warp < PR.wuse)
    for (int blk = blockIdx.x + gridDim.x * warp; blk < nblk; blk += gridDim.x * PR.wuse) {
        const int c = blk * 32 + lane;
        const bool valid = c < P.L.C;
        const int cc = valid ? c : P.L.C - 1;
        const int n = ch_len(P.L, cc);
        int start = P.start[cc];
        int end = min(P.end[cc], n);
        const int pk_c = P.peak[cc], en_c = P.enc[cc];
        const bool bad = pk_c >= P.S || en_c >= K;
        if (valid && bad) *P.overflow = MUA_ENC_BAD_TABLE;         // not a channel state this table block can code
        const bool act = valid && !bad && end > start && start >= 0;
        const int rep_len = max(end - start, 0);                    // what the report row carries
        if (!act) { start = 0; end = 0; }
        const uint32_t lut = lut_base + (act ? (uint32_t)(pk_c * K + en_c) * EF_LUT_B : 0u);
        uint8_t* out = P.stream + (size_t)cc * P.slot_bytes;
        uint32_t tb = 0;

        constexpr int ALIGN = FIXED ? 128 : ER_TILE;               // tile origin: a TMA box holds two steps
        const int tlo = __reduce_min_sync(FULL, act ? (start & ~(ALIGN - 1)) : 0x7FFFFFFF);
        const int thi = __reduce_max_sync(FULL, end);
        if (tlo < thi) {   // warp-uniform
            const int nt = (thi - tlo + ER_TILE - 1) / ER_TILE;    // steps of 64 symbols per lane
            // side info (entry j = chunk start/1024 + j; sub-chunk entry 8 j + i): the first entries are 0
            uint32_t* co = P.chunk_off + (size_t)cc * P.chunk_stride - (start >> 10);
            uint32_t* so = P.sub_off ? P.sub_off + (size_t)cc * P.sub_stride - 8 * (start >> 10) : nullptr;
            if (act) {
                co[start >> 10] = 0;
                if (so) so[start >> 7] = 0;
            }
            // ---- staging ----
            const int q16 = (lane & 3) * 16;
            const uint8_t* src_i[FIXED ? 1 : 4];
            int lim_i[FIXED ? 1 : 4];
            if (!FIXED) {   // copy instruction i moves 16 bytes of row 8 i + lane / 4 (piece lane % 4 of the row's 64-byte tile)
                const int rd_end = (end + 15) & ~15;
                const uint8_t* row = P.L.sym + ch_off(P.L, cc);
#pragma unroll
                for (int i = 0; i < (FIXED ? 1 : 4); ++i) {
                    const int r = 8 * i + (lane >> 2);
                    const unsigned long long rp = __shfl_sync(FULL, (unsigned long long)row, r);
                    src_i[i] = reinterpret_cast<const uint8_t*>(rp) + q16 + tlo;
                    lim_i[i] = __shfl_sync(FULL, rd_end, r) - q16 - tlo;       // the piece is requested while 64 t < lim
                }
            }
            const uint32_t dst_l = in0 + (lane >> 2) * ER_ROWB + q16;
            auto issue = [&](int t, uint32_t s) {   // ragged: request step t of every row into stage s
#pragma unroll
                for (int i = 0; i < (FIXED ? 0 : 4); ++i)
                    if (t * ER_TILE < lim_i[FIXED ? 0 : i]) cp_async16(dst_l + s * SM::STAGE + i * (8 * ER_ROWB), src_i[FIXED ? 0 : i] + t * ER_TILE);
            };
            // FIXED: request box tt (steps 2 tt and 2 tt + 1) into stage s; lane 0 only.  `dep` (== 0) is computed from the registers
            // the stage's last reads filled: a warp-wide LDS retires as a whole, so the request cannot be issued before EVERY lane's
            // reads of the stage have been performed (a __syncwarp alone only orders their issue, and with ~20 warps queueing
            // shared-memory loads an L2-hit box can land before a queued load executes)
            auto issue_box = [&](int tt, uint32_t s, uint32_t dep) {
                mbar_expect_tx_s(bar0 + 8 * s, SM::STAGE);
                tma_load_2d(in0 + s * SM::STAGE, &PR.tmap, tlo + 128 * tt + (int)dep, blk * 32, bar0 + 8 * s);
            };
            const int nbox = (nt + 1) >> 1;
            if (FIXED) {
                if (lane == 0) {
                    issue_box(0, 0, 0u);
                    if (nbox > 1) issue_box(1, 1, 0u);
                }
            } else {
                issue(0, 0);
                cp_async_commit();
            }
            unsigned long long hold = 0;
            int ts = tlo;
            for (int t = 0; t < nt; ++t, ts += ER_TILE) {
                uint32_t tile;
                if (FIXED) {
                    const uint32_t s = (t >> 1) & 1u;
                    if ((t & 1) == 0) {
                        mbar_wait_s(bar0 + 8 * s, (phase >> s) & 1u);
                        phase ^= 1u << s;
                    }
                    tile = (in_lane + s * SM::STAGE) ^ ((t & 1) * 64u);
                } else {
                    cp_async_wait<0>();       // the lane's pieces of step t have landed ...
                    __syncwarp();             // ... everybody's have, and everybody is done with step t - 1, whose stage is requested next
                    if (t + 1 < nt) issue(t + 1, (t & 1) ^ 1u);
                    cp_async_commit();
                    tile = in_lane + (t & 1) * SM::STAGE;
                }
                uint4 qv[4];
                uint32_t any_hi = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    qv[k] = lds_u128(FIXED ? tile ^ (16u * k) : tile + 16u * k);
                    any_hi |= (qv[k].x | qv[k].y) | (qv[k].z | qv[k].w);
                }
                if (FIXED && (t & 1) && (t >> 1) + 2 < nbox) {   // the box is in registers: its stage takes the box after the next one
                    const uint32_t dep = any_hi & PR.zero;
                    __syncwarp();
                    if (lane == 0) issue_box((t >> 1) + 2, (t >> 1) & 1u, dep);
                }
                if (any_hi & 0x80808080u) {   // rare: a count >= 128 (or stale bytes of a row that ended)
#pragma unroll
                    for (int k = 0; k < 4; ++k) qv[k] = clamp127(qv[k]);
                }
                // side info: a sub-chunk / chunk that starts with this step, strictly inside the lane's window
                if ((ts & 127) == 0 && ts > start && ts < end) {
                    if (so) so[ts >> 7] = tb;
                    if ((ts & (TILE - 1)) == 0) co[ts >> 10] = tb;
                }
                const uint32_t tb0 = tb;
                uint32_t pc[4], pl[4];
                if (__all_sync(FULL, ts >= start && ts + ER_TILE <= end)) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) encode16<SV>(qv[k], lut, pc[k], pl[k]);
                } else {
                    const uint32_t vlo4 = (uint32_t)min(max(start - ts, 0), ER_TILE) * 0x01010101u;   // valid symbols of the lane's step: [vlo, vhi)
                    const uint32_t vhi4 = (uint32_t)min(max(end - ts, 0), ER_TILE) * 0x01010101u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) encode16n<SV>(qv[k], lut + 256, 16u * k, vlo4, vhi4, pc[k], pl[k]);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) er_append(hold, tb, pc[k], pl[k], ring_lane);
                // a step adds <= 128 bits: at most one 128-bit unit completes
                if ((tb ^ tb0) & ~127u) {
                    const uint32_t u = tb0 >> 7;
                    const uint32_t rb = ring_lane + (u & 1u) * 512u;
                    uint4 v4;
                    v4.x = bswap32(lds_u32(rb)); v4.y = bswap32(lds_u32(rb + 128)); v4.z = bswap32(lds_u32(rb + 256)); v4.w = bswap32(lds_u32(rb + 384));
                    if (u < slot_units) *reinterpret_cast<uint4*>(out + (size_t)u * 16) = v4;
                    else *P.overflow = MUA_ENC_OVERFLOW;
                }
            }
            if (tb & 127u) {   // last partial unit, zero padded
                const uint32_t u = tb >> 7, nfull = tb >> 5, fill = tb & 31u;
                const uint32_t partial = fill ? ((uint32_t)hold << (32u - fill)) : 0u;
                uint32_t v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t wi = 4 * u + j;
                    v[j] = wi < nfull ? lds_u32(ring_lane + (wi & 7u) * 128u) : (wi == nfull ? partial : 0u);
                    v[j] = bswap32(v[j]);
                }
                if (u < slot_units) *reinterpret_cast<uint4*>(out + (size_t)u * 16) = make_uint4(v[0], v[1], v[2], v[3]);
                else *P.overflow = MUA_ENC_OVERFLOW;
            }
        }
        if (valid) {   // channel epilogue (publish_channel of the warp-per-channel encoders, one lane per channel)
            P.total_bits[c] = tb;
            const int4 rrow = make_int4((int)tb, rep_len, en_c, pk_c);
            for (int p = 0; p < P.n_peers; ++p) *reinterpret_cast<int4*>(P.rep[p] + 4 * (P.row0 + c)) = rrow;
        }
        if (!FIXED) cp_async_wait<0>();
        __syncwarp();   // everybody is done with the stages before the next block's first requests
    }
    signal_when_last(P);
}

// ---------------------------------------------------------------------------------------------
// the same for the pair encoder's codebooks (Lmax <= 8: every SCLV table up to S = 9), fixed-stride recordings
// ---------------------------------------------------------------------------------------------
// A lane codes its channel 64 symbols per step: the pair tables of k_encode_pair (base S+1, null digit for symbols outside the
// window) for ALL (peak, row) pairs of the alphabet sit in shared memory (S x K x 512 B: 7.5 KB at S = 5, 103 KB at S = 9), a
// lane addresses the one of its channel; two symbols per lookup, four lookups merged to an 8-symbol piece (<= 64 bits), pieces
// appended by the lane-private bit writer (the upper 32 bits of a piece only when some lane's piece is that long: warp vote).
// Rows come as TMA boxes of 64 bins x 32 channels (64-byte swizzle: piece k of row l at k ^ ((l >> 1) & 3)), two stages; the ring
// column holds 16 words and is flushed after every 32 symbols (<= 256 bits: at most two 128-bit units).
constexpr int ERP_STAGE = 2048;         // one box: 32 rows x 64 bytes
constexpr int ERP_RING = 16 * 32 * 4;   // 16 words x 32 lanes

// bank skew of the pair table of (peak, row) pair k: its words are stored XOR-ed with this (< 32: a word stays in its 128-byte group)
__host__ __device__ __forceinline__ uint32_t erp_skew(uint32_t k) { return (k * 5u) & 31u; }

struct EncRowsPairParams {
    EncParams E;
    int32_t wuse, warps;                // warps of a CTA that take blocks / that exist
    uint32_t zero;
    int32_t lut_bytes;                  // S * K * 512
    alignas(64) CUtensorMap tmap;       // box 64 x 32, 64-byte swizzle
};

template <int SV>
__global__ void __launch_bounds__(ER_WARPS * 32, 1) k_encode_rows_pair(const __grid_constant__ EncRowsPairParams PR) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const EncParams& P = PR.E;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TabHdr* T = reinterpret_cast<const TabHdr*>(P.tab);
    const int K = T->K;
    constexpr int S = SV;
    // shared memory: [warps x 2 stages][warps x ring][pair tables of all (peak, row) pairs][2 mbarriers per warp]
    uint8_t* s_lut = smem_raw + (size_t)PR.warps * (2 * ERP_STAGE + ERP_RING);
    const uint32_t lut_base = smem_u32(s_lut);
    if (T->S != SV || S != P.S || K != P.K || T->Lmax != P.Lmax || T->Lmax > 8 || T->encp_off == 0 || (lut_base & 255u) ||
        S * K * 512 != PR.lut_bytes) {
        if (threadIdx.x == 0) *P.overflow = MUA_ENC_BAD_TABLE;   // launch configuration does not match the table block
        return;
    }
    {   // word w of the table of pair k goes to word w ^ erp_skew(k): the lanes of a warp code channels with DIFFERENT (peak, row)
        // pairs, and the few hot entries (pairs of small counts) of all tables would otherwise share the same few banks
        // (ncu at 100k x 2 400, S = 5: 64 % of the kernel's shared-memory wavefronts were bank conflicts)
        const uint32_t* g = reinterpret_cast<const uint32_t*>(P.tab + T->encp_off);
        for (int i = threadIdx.x; i < PR.lut_bytes / 4; i += blockDim.x) {
            const int k = i >> 7, w = i & 127;
            reinterpret_cast<uint32_t*>(s_lut)[(k << 7) + (w ^ (int)erp_skew(k))] = g[i];
        }
    }
    const uint32_t in0 = smem_u32(smem_raw) + warp * (2 * ERP_STAGE);
    const uint32_t in_lane = (in0 + lane * 64) | (((lane >> 1) & 3) * 16);
    const uint32_t ring_lane = smem_u32(smem_raw) + PR.warps * (2 * ERP_STAGE) + warp * ERP_RING + lane * 4;
    const uint32_t bar0 = lut_base + PR.lut_bytes + warp * 16;
    if (lane == 0) {
        mbar_init(reinterpret_cast<uint64_t*>(s_lut + PR.lut_bytes) + 2 * warp, 1);
        mbar_init(reinterpret_cast<uint64_t*>(s_lut + PR.lut_bytes) + 2 * warp + 1, 1);
        fence_barrier_init();
    }
    __syncthreads();
    const uint32_t slot_units = (uint32_t)min((long long)(P.slot_bytes >> 4), 0x7FFFFFFFll);
    const int nblk = (P.L.C + 31) >> 5;
    uint32_t phase = 0;
    constexpr uint32_t satk = (uint32_t)(0x7F - (SV - 1)) * 0x01010101u;
    constexpr uint32_t satv = (uint32_t)(SV - 1) * 0x01010101u;
    constexpr uint32_t nullv = (uint32_t)SV * 0x01010101u;
    constexpr uint32_t mult = 2u | ((uint32_t)(2 * (SV + 1)) << 8);

    if (warp < PR.wuse)
    for (int blk = blockIdx.x + gridDim.x * warp; blk < nblk; blk += gridDim.x * PR.wuse) {
        const int c = blk * 32 + lane;
        const bool valid = c < P.L.C;
        const int cc = valid ? c : P.L.C - 1;
        int start = P.start[cc];
        int end = min(P.end[cc], P.L.T);
        const int pk_c = P.peak[cc], en_c = P.enc[cc];
        const bool bad = pk_c >= P.S || en_c >= K;
        if (valid && bad) *P.overflow = MUA_ENC_BAD_TABLE;
        const bool act = valid && !bad && end > start && start >= 0;
        const int rep_len = max(end - start, 0);
        if (!act) { start = 0; end = 0; }
        const uint32_t combo = act ? (uint32_t)(pk_c * K + en_c) : 0u;
        const uint32_t lut = lut_base + combo * 512u;
        const uint32_t skew2 = (erp_skew(combo) * 4u) * 0x01000100u;     // the table's skew as a byte offset, at the two offset bytes of `prod`
        uint8_t* out = P.stream + (size_t)cc * P.slot_bytes;
        uint32_t tb = 0;
        const int tlo = __reduce_min_sync(FULL, act ? (start & ~(ER_TILE - 1)) : 0x7FFFFFFF);
        const int thi = __reduce_max_sync(FULL, end);
        if (tlo < thi) {
            const int nt = (thi - tlo + ER_TILE - 1) / ER_TILE;
            uint32_t* co = P.chunk_off + (size_t)cc * P.chunk_stride - (start >> 10);
            if (act) co[start >> 10] = 0;
            auto issue_box = [&](int t, uint32_t s, uint32_t dep) {   // lane 0: box t (= step t) into stage s
                mbar_expect_tx_s(bar0 + 8 * s, ERP_STAGE);
                tma_load_2d(in0 + s * ERP_STAGE, &PR.tmap, tlo + ER_TILE * t + (int)dep, blk * 32, bar0 + 8 * s);
            };
            if (lane == 0) {
                issue_box(0, 0, 0u);
                if (nt > 1) issue_box(1, 1, 0u);
            }
            unsigned long long hold = 0;
            int ts = tlo;
            for (int t = 0; t < nt; ++t, ts += ER_TILE) {
                const uint32_t s = t & 1u;
                mbar_wait_s(bar0 + 8 * s, (phase >> s) & 1u);
                phase ^= 1u << s;
                const uint32_t tile = in_lane + s * ERP_STAGE;
                uint4 qv[4];
                uint32_t any_hi = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    qv[k] = lds_u128(tile ^ (16u * k));
                    any_hi |= (qv[k].x | qv[k].y) | (qv[k].z | qv[k].w);
                }
                if (t + 2 < nt) {   // the box is in registers: its stage takes the box after the next one
                    const uint32_t dep = any_hi & PR.zero;
                    __syncwarp();
                    if (lane == 0) issue_box(t + 2, s, dep);
                }
                if (any_hi & 0x80808080u) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) qv[k] = clamp127(qv[k]);
                }
                if ((ts & (TILE - 1)) == 0 && ts > start && ts < end) co[ts >> 10] = tb;
                const bool full = __all_sync(FULL, ts >= start && ts + ER_TILE <= end);
                const uint32_t vlo4 = (uint32_t)min(max(start - ts, 0), ER_TILE) * 0x01010101u;
                const uint32_t vhi4 = (uint32_t)min(max(end - ts, 0), ER_TILE) * 0x01010101u;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {   // 32 symbols: four 8-symbol pieces, then the complete units leave
                    const uint32_t tb0 = tb;
                    unsigned long long pc[4];
                    uint32_t pl[4], longest = 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint4 q = qv[2 * hf + (i >> 1)];
                        const uint32_t w2[2] = {(i & 1) ? q.z : q.x, (i & 1) ? q.w : q.y};
                        uint32_t qc[2], ql[2];
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint32_t gg = w2[j] + satk;
                            const uint32_t m = byte_msb_mask(gg);
                            uint32_t ws = (w2[j] & ~m) | (satv & m);
                            if (!full) {
                                const uint32_t iv = (0x83828180u + 0x04040404u * (uint32_t)(8 * hf + 2 * i + j));   // symbol indices in the step, bit 7 set
                                const uint32_t ok = (iv - vlo4) & ~(iv - vhi4);
                                const uint32_t vm = byte_msb_mask(ok);
                                ws = (ws & vm) | (nullv & ~vm);
                            }
                            const uint32_t prod = (ws * mult) ^ skew2;
                            const uint32_t a0 = __byte_perm(prod, lut, 0x7651), a1 = __byte_perm(prod, lut, 0x7653);
                            const uint32_t c0 = lds_u16(a0), l0 = lds_u8_256(a0), c1 = lds_u16(a1), l1 = lds_u8_256(a1);
                            qc[j] = (c0 << l1) | c1;
                            ql[j] = l0 + l1;
                        }
                        pc[i] = ((unsigned long long)qc[0] << ql[1]) | qc[1];
                        pl[i] = ql[0] + ql[1];
                        longest = max(longest, pl[i]);
                    }
                    const bool any_long = __any_sync(FULL, longest > 32u);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint32_t lo_len = pl[i];
                        if (any_long) {   // rare for MUA counts: a piece of more than 32 bits goes in two parts
                            const uint32_t hi_len = pl[i] > 32u ? pl[i] - 32u : 0u;
                            er_append<15u>(hold, tb, hi_len ? (uint32_t)(pc[i] >> 32) : 0u, hi_len, ring_lane);
                            lo_len = pl[i] - hi_len;
                        }
                        er_append<15u>(hold, tb, (uint32_t)pc[i], lo_len, ring_lane);
                    }
#pragma unroll
                    for (int r = 0; r < 2; ++r) {   // <= 256 new bits: at most two units complete
                        const uint32_t u = (tb0 >> 7) + r;
                        if (u < (tb >> 7)) {
                            const uint32_t rb = ring_lane + (u & 3u) * 512u;
                            uint4 v4;
                            v4.x = bswap32(lds_u32(rb)); v4.y = bswap32(lds_u32(rb + 128)); v4.z = bswap32(lds_u32(rb + 256)); v4.w = bswap32(lds_u32(rb + 384));
                            if (u < slot_units) *reinterpret_cast<uint4*>(out + (size_t)u * 16) = v4;
                            else *P.overflow = MUA_ENC_OVERFLOW;
                        }
                    }
                }
            }
            if (tb & 127u) {   // last partial unit, zero padded
                const uint32_t u = tb >> 7, nfull = tb >> 5, fill = tb & 31u;
                const uint32_t partial = fill ? ((uint32_t)hold << (32u - fill)) : 0u;
                uint32_t v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t wi = 4 * u + j;
                    v[j] = wi < nfull ? lds_u32(ring_lane + (wi & 15u) * 128u) : (wi == nfull ? partial : 0u);
                    v[j] = bswap32(v[j]);
                }
                if (u < slot_units) *reinterpret_cast<uint4*>(out + (size_t)u * 16) = make_uint4(v[0], v[1], v[2], v[3]);
                else *P.overflow = MUA_ENC_OVERFLOW;
            }
        }
        if (valid) {
            P.total_bits[c] = tb;
            const int4 rrow = make_int4((int)tb, rep_len, en_c, pk_c);
            for (int p = 0; p < P.n_peers; ++p) *reinterpret_cast<int4*>(P.rep[p] + 4 * (P.row0 + c)) = rrow;
        }
        __syncwarp();
    }
    signal_when_last(P);
}

}  // namespace mua
