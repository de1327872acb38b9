// Stages 2-4: windowed histograms (one scan per channel for all history lengths), approx-sort peak
// and rank map, SCLV cost + first-argmin selection, bit counts.  Also: train histograms, stand-alone
// selection, elimination-round scores, table build.
#pragma once
#include "mua_common.cuh"

namespace mua {

// ---------------------------------------------------------------------------------------------
// table build: one CTA per (peak p, codebook row k)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_build_tables(uint8_t* __restrict__ blob) {
    TabHdr* T = reinterpret_cast<TabHdr*>(blob);
    const int S = T->S, K = T->K, W = T->W;
    const int p = blockIdx.x / K, k = blockIdx.x % K;
    const int tid = threadIdx.x;
    __shared__ uint8_t s_rank[16], s_idx[16], s_len[16];
    __shared__ uint16_t s_code[16];
    if (tid < 16) {
        int s = tid < S ? tid : S - 1;
        int r = rank_of(p, s, S);
        s_rank[tid] = (uint8_t)r;              // raw nibble -> rank (saturation folded in)
        s_len[tid] = tid < S ? T->lens[k][tid] : 0;
        s_code[tid] = tid < S ? T->codes[k][tid] : 0;
    }
    __syncthreads();
    if (tid < S) s_idx[s_rank[tid]] = (uint8_t)tid;
    __syncthreads();
    if (k == 0 && tid < 16) {
        T->rank[p][tid] = s_rank[tid];
        T->idx[p][tid] = tid < S ? s_idx[tid] : 0;
    }
    uint32_t* enc1 = reinterpret_cast<uint32_t*>(blob + T->enc1_off) + (size_t)(p * K + k) * 16;
    uint32_t* enc2 = reinterpret_cast<uint32_t*>(blob + T->enc2_off) + (size_t)(p * K + k) * 256;
    uint32_t* dec = reinterpret_cast<uint32_t*>(blob + T->dec_off) + ((size_t)(p * K + k) << W);
    if (tid < 16) enc1[tid] = ((uint32_t)s_len[s_rank[tid]] << 16) | s_code[s_rank[tid]];
    {
        int r0 = s_rank[tid & 15], r1 = s_rank[tid >> 4];
        uint32_t l1 = s_len[r1];
        enc2[tid] = (((uint32_t)s_code[r0] << l1) | s_code[r1]) | (((uint32_t)s_len[r0] + l1) << 24);
    }
    if (T->enc4_off) {   // Lmax <= 2 and S <= 3: four saturated symbols per entry (fast encoder), 768 bytes per (peak, row) pair
        uint8_t* enc4 = blob + T->enc4_off + (size_t)(p * K + k) * 768;
        if (tid < S * S * S * S) {                    // full tiles: index in base S -> codes [0,128), lengths [128,256)
            uint32_t code = 0, len = 0;
            int rest = tid;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = s_rank[rest % S];
                rest /= S;
                code = (code << s_len[r]) | s_code[r];
                len += s_len[r];
            }
            enc4[tid] = (uint8_t)code;
            enc4[128 + tid] = (uint8_t)len;
        }
        if (tid < (S + 1) * (S + 1) * (S + 1) * (S + 1)) {   // partial tiles: index in base S+1, digit S = outside the window, no bits
            uint32_t code = 0, len = 0;                      // -> codes [256,512), lengths [512,768)
            int rest = tid;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int q = rest % (S + 1);
                rest /= (S + 1);
                if (q < S) {
                    const int r = s_rank[q];
                    code = (code << s_len[r]) | s_code[r];
                    len += s_len[r];
                }
            }
            enc4[256 + tid] = (uint8_t)code;
            enc4[512 + tid] = (uint8_t)len;
        }
    }
    if (T->encp_off && tid < (S + 1) * (S + 1)) {   // Lmax <= 8: symbol pairs, index in base S+1 (digit S = no symbol)
        uint8_t* encp = blob + T->encp_off + (size_t)(p * K + k) * 512;
        const int q0 = tid / (S + 1), q1 = tid % (S + 1);
        uint32_t code = 0, len = 0;
        if (q0 < S) { const int r = s_rank[q0]; code = s_code[r]; len = s_len[r]; }
        if (q1 < S) { const int r = s_rank[q1]; code = (code << s_len[r]) | s_code[r]; len += s_len[r]; }
        reinterpret_cast<uint16_t*>(encp)[tid] = (uint16_t)code;
        encp[256 + 2 * tid] = (uint8_t)len;
    }
    const int nsym = T->nsym;
    for (int v = tid; v < (1 << W); v += blockDim.x) {
        uint32_t e = 0;
        int used = 0;
        for (int n = 0; n < nsym; ++n) {
            int hit = 0;   // the rows are Kraft-complete and W = nsym*Lmax, so a codeword always matches
            for (int r = 0; r < S; ++r) {
                int l = s_len[r];
                if (used + l <= W && (uint32_t)((v >> (W - used - l)) & ((1 << l) - 1)) == s_code[r]) { hit = r; break; }
            }
            e |= (uint32_t)s_idx[hit] << (8 * n);
            used += s_len[hit];
        }
        dec[v] = e | ((uint32_t)used << 28);
    }
    if (p == 0) {   // variable-count rank table of row k (the rank -> symbol map of the lane's peak is applied at decode time)
        const int Wv = T->Wv;
        uint32_t* decv = reinterpret_cast<uint32_t*>(blob + T->decv_off) + ((size_t)k << Wv);
        const uint32_t none = S <= 8 ? 0x8u : 0xFu;
        for (int v = tid; v < (1 << Wv); v += blockDim.x) {
            uint32_t sel = none * 0x1111u;
            int used = 0, n = 0;
            for (; n < 4; ++n) {
                int hit = -1;
                for (int r = 0; r < S; ++r) {
                    const int l = s_len[r];
                    if (used + l <= Wv && (uint32_t)((v >> (Wv - used - l)) & ((1 << l) - 1)) == s_code[r]) { hit = r; break; }
                }
                if (hit < 0) break;   // the next codeword does not fit the window any more
                sel = (sel & ~(0xFu << (4 * n))) | ((uint32_t)hit << (4 * n));
                used += s_len[hit];
            }
            decv[v] = sel | ((uint32_t)n << 16) | ((uint32_t)used << 20);
        }
    }
}

// sum of the four byte counters of a packed word (each <= 252)
__device__ __forceinline__ int bytesum4(uint32_t x) {
    const uint32_t y = (x & 0x00FF00FFu) + ((x >> 8) & 0x00FF00FFu);
    return (int)((y & 0xFFFFu) + (y >> 16));
}

// what a lane's packed streaming counter holds, as a count
__device__ __forceinline__ int cal_flush(uint32_t x);

// ---------------------------------------------------------------------------------------------
// calibrate: warp per channel
// ---------------------------------------------------------------------------------------------
// One scan of a channel serves every history length AND every alphabet size that was asked for: the scan counts
// #{x >= v} for v = 1..SS-1 (SS = largest alphabet), and the histogram of the data saturated at S-1 is
// hist_S[s] = G_s - G_{s+1} (s < S-1), hist_S[S-1] = G_{S-1} -- saturation never changes whether x >= v for v < S.
// The reference re-reads and re-saturates every channel for each S of its sweep (get_BR_no_sort.py:107, :143, :164).
constexpr int CAL_MAX_NS = 9;   // alphabet sizes per launch (2..10)

struct CalOut {                 // outputs and tables of one alphabet size
    int32_t S;
    const TabHdr* tab;
    unsigned long long active;
    int32_t* cutoff;
    int32_t* end;
    uint8_t* peak;
    uint8_t* enc;
    int32_t* assign_m;
    int32_t* post_m;
    int64_t* bits;
    int64_t* nsym;
    int32_t* train_hist;
};

struct CalibParams {
    Layout L;
    int32_t nH, use_sort, mode, train, nS, need_post;
    uint32_t one;               // == 1, opaque to ptxas: multiplier of the adds that are to run on the FMA pipe
    int32_t H[MUA_MAX_H];
    CalOut out[CAL_MAX_NS];
};

#ifndef MUA_CAL_ACC
#define MUA_CAL_ACC 1          // 0: byte counters on the ALU pipe; 1: IMAD compare-add + DP4A flag sum
#endif
#ifndef MUA_CAL_NALU
#define MUA_CAL_NALU 2         // thresholds whose compare-add stays on the ALU pipe (pipe balance)
#endif
__device__ __forceinline__ int cal_flush(uint32_t x) {
#if MUA_CAL_ACC == 1
    return (int)(x >> 7);       // DP4A adds 0x80 per flagged byte
#else
    return bytesum4(x);         // four byte counters
#endif
}
constexpr int CAL_WARPS = 4;
constexpr int CAL_TILE = 512;   // bytes per warp step (16 B per lane)

struct CalSmem {                // per warp
    int bnd[2 * MUA_MAX_H];
    int am[MUA_MAX_H][MUA_MAX_S], pm[MUA_MAX_H][MUA_MAX_S];   // mapped histograms per history length
    long long cost[32];
    int k[32];
    uint4 len[MUA_MAX_K], rank[MUA_MAX_S];                    // SCLV rows and rank maps of the current table block (16 B rows)
};

// Epilogue of one alphabet size S (snapshots counted with SS >= S thresholds):
//   A: lane h < nH turns the snapshots of history length h into the calibration and post-window histograms,
//      finds the peak and maps both histograms through the approx-sort rank map;
//   B: the K cost dot products of every history length are spread over G = 32 / nH lanes each (row k on lane
//      k mod G); the (cost, row) pairs meet in shared memory and the first lane of a group takes the first minimum;
//   C: that lane counts the post-window bits with the chosen row and writes the outputs.
template <int S, int SS>
__device__ __forceinline__ void cal_epilogue(const CalibParams& P, const CalOut& O, int c, int lane, CalSmem& W, const int (*snap)[SS]) {
    const int nH = P.nH;
    if (!P.train) {   // this alphabet's SCLV rows and rank maps -> shared memory of this warp
        __syncwarp();
        const uint4* gl = reinterpret_cast<const uint4*>(&O.tab->lens[0][0]);
        const uint4* gr = reinterpret_cast<const uint4*>(&O.tab->rank[0][0]);
        for (int i = lane; i < MUA_MAX_K; i += 32) W.len[i] = gl[i];
        if (lane < MUA_MAX_S) W.rank[lane] = gr[lane];
        __syncwarp();
    }
    const uint8_t(*s_len)[16] = reinterpret_cast<const uint8_t(*)[16]>(W.len);
    const uint8_t(*s_rank)[16] = reinterpret_cast<const uint8_t(*)[16]>(W.rank);
    int cut = 0, end = 0, p = 0;
    if (lane < nH) {
        const int h = lane;
        cut = W.bnd[h];
        end = W.bnd[nH + h];
        int hist[S], post[S];
        {
            int g_prev = cut;   // G_0 = number of samples
#pragma unroll
            for (int s = 0; s < S; ++s) {
                int g_next = s + 1 < S ? snap[h][s + 1] : 0;
                hist[s] = g_prev - g_next;
                g_prev = g_next;
            }
        }
        const bool has_post = end > 0 && P.mode != MUA_WINDOW_NONE;
        {
            int g_prev = has_post ? end - cut : 0;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                int g_next = (has_post && s + 1 < S) ? snap[nH + h][s + 1] - snap[h][s + 1] : 0;
                post[s] = g_prev - g_next;
                g_prev = g_next;
            }
        }
        if (P.train) {
            // np.flip(np.sort(hist)): descending (get_BR_no_sort.py:147)
#pragma unroll
            for (int i = 1; i < S; ++i) {
#pragma unroll
                for (int j = S - 1; j >= i; --j) {
                    int a = hist[j - 1], b = hist[j];
                    hist[j - 1] = max(a, b);
                    hist[j] = min(a, b);
                }
            }
            const size_t o = (size_t)c * nH + h;
#pragma unroll
            for (int s = 0; s < S; ++s) O.train_hist[o * S + s] = hist[s];
        } else {
            if (P.use_sort) {   // np.argmax: lowest index on ties (functions_1.py:77)
                int best = hist[0];
#pragma unroll
                for (int s = 1; s < S; ++s)
                    if (hist[s] > best) { best = hist[s]; p = s; }
            }
#pragma unroll
            for (int s = 0; s < S; ++s) {   // mapped histograms: m[rank[s]] = hist[s]
                const int r = s_rank[p][s];
                W.am[h][r] = hist[s];
                W.pm[h][r] = post[s];
            }
        }
    }
    if (P.train) return;
    __syncwarp();
    const int K = O.tab->K;
    const int G = 32 / nH;                       // lanes per history length (>= 2)
    const int gh = lane / G, gj = lane - gh * G;
    {
        long long best_cost = 0x7FFFFFFFFFFFFFFFll;
        int enc = -1;
        if (gh < nH) {
            int am[S];
#pragma unroll
            for (int r = 0; r < S; ++r) am[r] = W.am[gh][r];
            for (int k = gj; k < K; k += G) {
                if (!((O.active >> k) & 1ull)) continue;
                long long cost = 0;
#pragma unroll
                for (int r = 0; r < S; ++r) cost += (long long)am[r] * s_len[k][r];
                if (enc < 0 || cost < best_cost) { best_cost = cost; enc = k; }
            }
        }
        W.cost[lane] = best_cost;
        W.k[lane] = enc;
    }
    __syncwarp();
    if (lane < nH) {
        const int h = lane;
        long long best_cost = 0;
        int enc = -1;
        for (int j = 0; j < G; ++j) {   // first minimum over the rows: lowest cost, then lowest row index
            const int k = W.k[h * G + j];
            const long long cost = W.cost[h * G + j];
            if (k >= 0 && (enc < 0 || cost < best_cost || (cost == best_cost && k < enc))) { best_cost = cost; enc = k; }
        }
        if (enc < 0) enc = 0;
        long long bits = 0, ns = 0;
#pragma unroll
        for (int r = 0; r < S; ++r) { const int pmr = W.pm[h][r]; bits += (long long)pmr * s_len[enc][r]; ns += pmr; }
        const size_t o = (size_t)c * nH + h;
        if (O.cutoff) O.cutoff[o] = cut;
        if (O.end) O.end[o] = (P.mode == MUA_WINDOW_NONE) ? cut : end;
        if (O.peak) O.peak[o] = (uint8_t)p;
        if (O.enc) O.enc[o] = (uint8_t)enc;
        if (O.bits) O.bits[o] = bits;
        if (O.nsym) O.nsym[o] = ns;
        if (O.assign_m) {
#pragma unroll
            for (int r = 0; r < S; ++r) O.assign_m[o * S + r] = W.am[h][r];
        }
        if (O.post_m) {
#pragma unroll
            for (int r = 0; r < S; ++r) O.post_m[o * S + r] = W.pm[h][r];
        }
    }
    __syncwarp();
}

template <int S, int SS>
__device__ __forceinline__ void cal_epilogue_if(const CalibParams& P, const CalOut& O, int c, int lane, CalSmem& W, const int (*snap)[SS]) {
    if constexpr (S <= SS) cal_epilogue<S, SS>(P, O, c, lane, W, snap);
}

// SS = number of symbol values the scan distinguishes (thresholds 1..SS-1); MULTI: the epilogue runs for every
// alphabet size in P.out[0..nS) (each <= SS), otherwise for P.out[0] with S == SS.
template <int SS, bool MULTI>
__global__ void __launch_bounds__(CAL_WARPS * 32) k_calibrate(const __grid_constant__ CalibParams P) {
    constexpr int S = SS;
    __shared__ int s_snap[CAL_WARPS][2 * MUA_MAX_H][S];   // [boundary][v] = #{t < b : x_t >= v}, v = 1..S-1
    __shared__ CalSmem s_w[CAL_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * CAL_WARPS + warp;
    if (c >= P.L.C) return;
    CalSmem& W = s_w[warp];
    const int nH = P.nH, nB = 2 * nH;
    const int n = ch_len(P.L, c);
    const uint8_t* row = P.L.sym + ch_off(P.L, c);

    // boundaries: cutoffs then window ends (0 = unused)
    if (lane < nH) {
        int H = P.H[lane];
        int cut = n > 0 ? min(max(H, 1), n) : 0;
        int end = 0;
        if (P.mode == MUA_WINDOW_SKIP) {
            end = cut + n / 2;
            if (end > n) end = -1;
        } else if (P.mode == MUA_WINDOW_TRUNCATE) {
            end = min(cut + n / 2, n);
        }
        W.bnd[lane] = cut;
        W.bnd[nH + lane] = end;
    }
    for (int i = lane; i < nB * S; i += 32) (&s_snap[warp][0][0])[i] = 0;
    __syncwarp();
    // the post window is only scanned when a post-window output was asked for
    const bool need_post = P.need_post != 0;
    // lane bi owns boundary bi (nB <= 32): 0 or negative = unused
    const int myb = (lane < (need_post ? nB : nH)) ? W.bnd[lane] : 0;
    const int scan_end = __reduce_max_sync(FULL, max(myb, 0));

    // Counters: per threshold v a packed word of four byte counters (flags of 4 bytes x 4 words per step,
    // flushed into 32-bit counters before they can overflow) -- no POPC on the streaming path.
    uint32_t accb[S];
    int acc[S];
#pragma unroll
    for (int v = 0; v < S; ++v) { acc[v] = 0; accb[v] = 0; }
    const int rd_end = (n + 15) & ~15;   // rows are readable up to round_up(len, 16)
    auto next_boundary = [&](int after) {   // smallest boundary > after that still matters (INT_MAX if none)
        return __reduce_min_sync(FULL, myb > after ? myb : 0x7FFFFFFF);
    };
    int nextb = next_boundary(0);
    int since_flush = 0;
    constexpr int UNR = 4;               // 16-byte loads in flight per lane (2 KB per warp step)
    for (int t0 = 0; t0 < scan_end; t0 += UNR * CAL_TILE) {
        uint4 qv[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int p0 = t0 + u * CAL_TILE + lane * 16;
            qv[u] = make_uint4(0, 0, 0, 0);
            if (p0 < rd_end && t0 + u * CAL_TILE < scan_end) qv[u] = *reinterpret_cast<const uint4*>(row + p0);
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int ts = t0 + u * CAL_TILE;
            if (ts >= scan_end) break;
            const uint32_t w[4] = {qv[u].x, qv[u].y, qv[u].z, qv[u].w};
            uint32_t lo7[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) lo7[j] = w[j] & 0x7F7F7F7Fu;
            if (nextb <= ts + CAL_TILE) {
                // Boundaries inside (ts, ts + CAL_TILE]: snapshot the cumulative counts (rare for long rows, everything
                // for short ones).  All boundaries of the tile are served at once: per threshold the warp total up to the
                // tile start (one REDUX) and the lane's 16 flag bits; an exclusive scan over lanes of the flag counts,
                // three thresholds packed per word (a warp total is <= 512 < 2^10); then the lane that owns a boundary
                // fetches the scan value and the flags of the lane the boundary falls into by shuffle and finishes alone.
                int base[S];
                uint32_t u[S];                     // flag of byte i of word j at bit 8i + j
                constexpr int NPK = (S - 1 + 2) / 3;
                uint32_t ex[NPK];
#pragma unroll
                for (int v = 1; v < S; ++v) {
                    base[v] = __reduce_add_sync(FULL, acc[v] + cal_flush(accb[v]));
                    u[v] = (ge_mask(w[0], lo7[0], v) >> 7) | (ge_mask(w[1], lo7[1], v) >> 6) | (ge_mask(w[2], lo7[2], v) >> 5) |
                           (ge_mask(w[3], lo7[3], v) >> 4);
                }
#pragma unroll
                for (int g = 0; g < NPK; ++g) {
                    uint32_t pk = 0;
#pragma unroll
                    for (int i = 0; i < 3; ++i)
                        if (1 + 3 * g + i < S) pk |= (uint32_t)__popc(u[1 + 3 * g + i]) << (10 * i);
                    uint32_t incl = pk;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t t = __shfl_up_sync(FULL, incl, d);
                        if (lane >= d) incl += t;
                    }
                    ex[g] = incl - pk;
                }
                const bool mine = myb > ts && myb <= ts + CAL_TILE;
                const int q = mine ? myb - ts : 1;            // bytes of the tile before the boundary: 1..512
                const int L = (q - 1) >> 4;                   // the lane the boundary falls into
                const int nvalid = q - 16 * L;                // ... and how many of its bytes count: 1..16
                const int full = nvalid >> 2, part = nvalid & 3;
                const uint32_t vm = ((1u << full) - 1u) * 0x01010101u | ((1u << full) * (0x01010101u & ((1u << (8 * part)) - 1u)));
#pragma unroll
                for (int g = 0; g < NPK; ++g) {
                    const uint32_t exL = __shfl_sync(FULL, ex[g], L);
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const int v = 1 + 3 * g + i;
                        if (v < S) {
                            const uint32_t uL = __shfl_sync(FULL, u[v], L);
                            if (mine) s_snap[warp][lane][v] = base[v] + (int)((exL >> (10 * i)) & 1023u) + __popc(uL & vm);
                        }
                    }
                }
                nextb = next_boundary(ts + CAL_TILE);
            }
#if MUA_CAL_ACC == 0
#pragma unroll
            for (int v = 1; v < S; ++v) {
                accb[v] += (ge_mask(w[0], lo7[0], v) >> 7) + (ge_mask(w[1], lo7[1], v) >> 7) +
                           (ge_mask(w[2], lo7[2], v) >> 7) + (ge_mask(w[3], lo7[3], v) >> 7);
            }
            if (++since_flush == 63) {   // byte counters hold at most 63 * 4 = 252
#pragma unroll
                for (int v = 1; v < S; ++v) { acc[v] += bytesum4(accb[v]); accb[v] = 0; }
                since_flush = 0;
            }
#else
            // With byte counters the loop is bound by the ALU pipe (per word and threshold VIADD, LOP3, SHF, IADD, all
            // issued every other cycle per scheduler) while the FMA pipe idles.  Here the compare-add t = lo7 * 1 + c
            // is an IMAD (the multiplier is a kernel parameter ptxas cannot fold) and the four flag bytes of a word
            // are summed into a 32-bit counter by one DP4A (0x80 per flag; no shift, no overflow flushes): per word
            // and threshold ONE ALU op (LOP3: (t | w) & 0x80808080).  MUA_CAL_NALU thresholds keep their compare-add
            // on the ALU pipe.  Measured (100k channels x 120k bins, nine history lengths, 6.1 GB scanned): S=5
            // 1.62 -> 1.01 ms, S=7 2.26 -> 1.35, S=10 3.34 -> 1.97; IMAD.HI (x * 2^25 >> 32) instead of DP4A is slower.
#pragma unroll
            for (int v = 1; v < S; ++v) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t t;
                    const uint32_t cv = (uint32_t)(0x80 - v) * 0x01010101u;
                    if (v <= MUA_CAL_NALU) t = lo7[j] + cv;
                    else asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(lo7[j]), "r"(P.one), "r"(cv));
                    const uint32_t m = (t | w[j]) & 0x80808080u;
                    asm("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(accb[v]) : "r"(m), "r"(0x01010101u));
                }
            }
            if (++since_flush == (1 << 19)) {   // 2048 per step at most: a 32-bit counter lasts 2^21 steps
#pragma unroll
                for (int v = 1; v < S; ++v) { acc[v] += cal_flush(accb[v]); accb[v] = 0; }
                since_flush = 0;
            }
#endif
        }
    }
    __syncwarp();

    // ---- epilogue(s) ----
    const int (*snap)[S] = s_snap[warp];
    if (!MULTI) {
        cal_epilogue<S, S>(P, P.out[0], c, lane, W, snap);
    } else {
        for (int i = 0; i < P.nS; ++i) {
            const CalOut& O = P.out[i];
            switch (O.S) {
                case 2: cal_epilogue_if<2, S>(P, O, c, lane, W, snap); break;
                case 3: cal_epilogue_if<3, S>(P, O, c, lane, W, snap); break;
                case 4: cal_epilogue_if<4, S>(P, O, c, lane, W, snap); break;
                case 5: cal_epilogue_if<5, S>(P, O, c, lane, W, snap); break;
                case 6: cal_epilogue_if<6, S>(P, O, c, lane, W, snap); break;
                case 7: cal_epilogue_if<7, S>(P, O, c, lane, W, snap); break;
                case 8: cal_epilogue_if<8, S>(P, O, c, lane, W, snap); break;
                case 9: cal_epilogue_if<9, S>(P, O, c, lane, W, snap); break;
                case 10: cal_epilogue_if<10, S>(P, O, c, lane, W, snap); break;
                default: break;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// calibrate, calibration window only: G lanes per channel
// ---------------------------------------------------------------------------------------------
// The streaming system (test_chosen_system.py:80-97, one history length, bit counts taken from the encoder)
// needs only the first cutoff <= H <= 1024 samples of a channel: cutoff, window end, peak and SCLV row.  A whole
// warp per channel (k_calibrate) spends ~1000 instructions on those 64 bytes; here G lanes share a channel, each
// counting #{x >= v} over its 16-byte pieces, the counts meet by xor-shuffles, every lane of the group rebuilds
// the histogram / peak / rank map, the SCLV rows are dealt round-robin to the lanes and the (cost, row) pairs are
// min-reduced (lowest cost, then lowest row = np.argmin's first minimum, get_BR_no_sort.py:236).
constexpr int CALH_THREADS = 128;

template <int S, int G>
__global__ void __launch_bounds__(CALH_THREADS) k_calibrate_head(const __grid_constant__ CalibParams P) {
    __shared__ uint4 s_len4[MUA_MAX_K], s_rank4[MUA_MAX_S];
    const CalOut& O = P.out[0];
    {
        const uint4* gl = reinterpret_cast<const uint4*>(&O.tab->lens[0][0]);
        const uint4* gr = reinterpret_cast<const uint4*>(&O.tab->rank[0][0]);
        for (int i = threadIdx.x; i < MUA_MAX_K; i += CALH_THREADS) s_len4[i] = gl[i];
        if (threadIdx.x < MUA_MAX_S) s_rank4[threadIdx.x] = gr[threadIdx.x];
    }
    __syncthreads();
    const uint8_t(*s_len)[16] = reinterpret_cast<const uint8_t(*)[16]>(s_len4);
    const uint8_t(*s_rank)[16] = reinterpret_cast<const uint8_t(*)[16]>(s_rank4);
    const int c_raw = (int)(((long long)blockIdx.x * CALH_THREADS + threadIdx.x) / G);
    const int g = threadIdx.x % G;
    const bool valid = c_raw < P.L.C;
    const int c = valid ? c_raw : P.L.C - 1;          // idle groups shadow the last channel (shuffles stay warp-wide)
    const int n = ch_len(P.L, c);
    const uint8_t* row = P.L.sym + ch_off(P.L, c);
    const int cut = n > 0 ? min(max(P.H[0], 1), n) : 0;

    int cnt[S];                                        // cnt[v] = #{t < cut : x_t >= v}
#pragma unroll
    for (int v = 0; v < S; ++v) cnt[v] = 0;
    for (int p0 = g * 16; p0 < cut; p0 += G * 16) {
        const uint4 q = *reinterpret_cast<const uint4*>(row + p0);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        const int nvalid = min(16, cut - p0);          // bytes of this piece inside the window: 1..16
        const int full = nvalid >> 2, part = nvalid & 3;
        // flag of byte i of word j sits at bit 8i + j: words below `full` count whole, word `full` its first `part` bytes
        const uint32_t vm = ((1u << full) - 1u) * 0x01010101u | ((1u << full) * (0x01010101u & ((1u << (8 * part)) - 1u)));
        uint32_t lo7[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) lo7[j] = w[j] & 0x7F7F7F7Fu;
#pragma unroll
        for (int v = 1; v < S; ++v) {
            const uint32_t u = (ge_mask(w[0], lo7[0], v) >> 7) | (ge_mask(w[1], lo7[1], v) >> 6) |
                               (ge_mask(w[2], lo7[2], v) >> 5) | (ge_mask(w[3], lo7[3], v) >> 4);
            cnt[v] += __popc(u & vm);
        }
    }
#pragma unroll
    for (int v = 1; v < S; ++v) {
#pragma unroll
        for (int d = G / 2; d; d >>= 1) cnt[v] += __shfl_xor_sync(FULL, cnt[v], d);
    }
    int hist[S];
    {
        int g_prev = cut;                              // G_0 = number of samples
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int g_next = s + 1 < S ? cnt[s + 1] : 0;
            hist[s] = g_prev - g_next;
            g_prev = g_next;
        }
    }
    int p = 0;
    if (P.use_sort) {                                  // np.argmax: lowest index on ties (functions_1.py:77)
        int best = hist[0];
#pragma unroll
        for (int s = 1; s < S; ++s)
            if (hist[s] > best) { best = hist[s]; p = s; }
    }
    int rk[S];
#pragma unroll
    for (int s = 0; s < S; ++s) rk[s] = s_rank[p][s];
    const int K = O.tab->K;
    long long best_cost = 0;
    int enc = -1;
    for (int k = g; k < K; k += G) {
        if (!((O.active >> k) & 1ull)) continue;
        long long cost = 0;
#pragma unroll
        for (int s = 0; s < S; ++s) cost += (long long)hist[s] * s_len[k][rk[s]];
        if (enc < 0 || cost < best_cost) { best_cost = cost; enc = k; }
    }
#pragma unroll
    for (int d = G / 2; d; d >>= 1) {
        const long long oc = __shfl_xor_sync(FULL, best_cost, d);
        const int ok = __shfl_xor_sync(FULL, enc, d);
        if (ok >= 0 && (enc < 0 || oc < best_cost || (oc == best_cost && ok < enc))) { best_cost = oc; enc = ok; }
    }
    if (!valid || g != 0) return;
    int end = cut;
    if (P.mode == MUA_WINDOW_SKIP) {
        end = cut + n / 2;
        if (end > n) end = -1;
    } else if (P.mode == MUA_WINDOW_TRUNCATE) {
        end = min(cut + n / 2, n);
    }
    if (O.cutoff) O.cutoff[c] = cut;
    if (O.end) O.end[c] = end;
    if (O.peak) O.peak[c] = (uint8_t)p;
    if (O.enc) O.enc[c] = (uint8_t)(enc < 0 ? 0 : enc);
    if (O.assign_m) {
#pragma unroll
        for (int s = 0; s < S; ++s) O.assign_m[(size_t)c * S + rk[s]] = hist[s];
    }
}

// ---------------------------------------------------------------------------------------------
// stand-alone selection / bit counts / elimination scores (thread per histogram)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_select(const int32_t* __restrict__ hist, int64_t N, const TabHdr* __restrict__ T,
                                                unsigned long long active, uint8_t* __restrict__ enc,
                                                int64_t* __restrict__ min1, int64_t* __restrict__ min2) {
    __shared__ uint8_t s_len[MUA_MAX_K][16];
    for (int i = threadIdx.x; i < MUA_MAX_K * 16; i += blockDim.x) (&s_len[0][0])[i] = (&T->lens[0][0])[i];
    __syncthreads();
    const int S = T->S, K = T->K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        int hv[MUA_MAX_S];
#pragma unroll
        for (int s = 0; s < MUA_MAX_S; ++s) hv[s] = s < S ? hist[i * S + s] : 0;
        long long m1 = 0, m2 = 0;
        int e = -1, cnt = 0;
        for (int k = 0; k < K; ++k) {
            if (!((active >> k) & 1ull)) continue;
            long long cost = 0;
#pragma unroll
            for (int r = 0; r < MUA_MAX_S; ++r) cost += (long long)hv[r] * s_len[k][r];
            if (cnt == 0) { m1 = cost; e = k; }
            else if (cost < m1) { m2 = m1; m1 = cost; e = k; }
            else if (cnt == 1 || cost < m2) { m2 = cost; }
            ++cnt;
        }
        if (cnt < 2) m2 = m1;
        enc[i] = (uint8_t)(e < 0 ? 0 : e);
        if (min1) min1[i] = m1;
        if (min2) min2[i] = m2;
    }
}

__global__ void __launch_bounds__(256) k_bit_counts(const int32_t* __restrict__ hist, const uint8_t* __restrict__ enc, int64_t N,
                                                    const TabHdr* __restrict__ T, int64_t* __restrict__ bits,
                                                    int64_t* __restrict__ nsym) {
    const int S = T->S;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = enc[i];
        long long b = 0, n = 0;
        for (int r = 0; r < S; ++r) {
            int hv = hist[i * S + r];
            b += (long long)hv * T->lens[k][r];
            n += hv;
        }
        bits[i] = b;
        nsym[i] = n;
    }
}

__global__ void __launch_bounds__(256) k_elim_scores(const uint8_t* __restrict__ enc, const int64_t* __restrict__ min1,
                                                     const int64_t* __restrict__ min2, int64_t N, int K,
                                                     unsigned long long* __restrict__ assign_hist,
                                                     unsigned long long* __restrict__ score) {
    __shared__ unsigned long long s_delta[MUA_MAX_K], s_cnt[MUA_MAX_K], s_tot;
    if (threadIdx.x < MUA_MAX_K) { s_delta[threadIdx.x] = 0; s_cnt[threadIdx.x] = 0; }
    if (threadIdx.x == 0) s_tot = 0;
    __syncthreads();
    unsigned long long tot = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        int e = enc[i];
        tot += (unsigned long long)min1[i];
        atomicAdd(&s_delta[e], (unsigned long long)(min2[i] - min1[i]));
        atomicAdd(&s_cnt[e], 1ull);
    }
    atomicAdd(&s_tot, tot);
    __syncthreads();
    if (threadIdx.x < K) {
        atomicAdd(&score[threadIdx.x], s_tot + s_delta[threadIdx.x]);
        atomicAdd(&assign_hist[threadIdx.x], s_cnt[threadIdx.x]);
    }
}

}  // namespace mua
