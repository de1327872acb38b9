// C ABI of libmua_b200.so (see include/mua_b200.h): argument validation + kernel launches.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "mua_calibrate.cuh"
#include "mua_calibrate_rows.cuh"
#include "mua_decode.cuh"
#include "mua_decode_rows.cuh"
#include "mua_dropin.cuh"
#include "mua_encode.cuh"
#include "mua_encode_rows.cuh"

using namespace mua;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what) { return fail(MUA_E_CUDA, "%s: %s", what, cudaGetErrorString(e)); }

#define CHECK_LAUNCH(what)                                   \
    do {                                                     \
        cudaError_t e__ = cudaGetLastError();                \
        if (e__ != cudaSuccess) return cuda_fail(e__, what); \
    } while (0)

#define REQUIRE(cond, ...)                                    \
    do {                                                      \
        if (!(cond)) return fail(MUA_E_INVALID, __VA_ARGS__); \
    } while (0)

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        cached_dev = dev;
    }
    return cached > 0 ? cached : 148;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// cuTensorMapEncodeTiled through the runtime (no link against libcuda): the tensor view of a decoded buffer for k_decode_sub
typedef CUresult (*TensorMapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                           const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TensorMapEncodeTiledFn tensor_map_encoder() {
    static TensorMapEncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<TensorMapEncodeTiledFn>(p);
    }();
    return fn;
}

// rows of at most this many bins are encoded by the lane-per-channel kernels (MUA_ROWS_T overrides it for A/B measurements)
int rows_t_max() {
    static const int v = [] {
        const char* e = getenv("MUA_ROWS_T");
        return e ? atoi(e) : 16384;
    }();
    return v;
}

// ... and only for recordings with at least this many channels: a warp takes 32 channels, so the lane-per-channel kernels need
// ~8 blocks per SM to keep the SMs busy (MUA_ROWS_MIN_C overrides it, read at every call: the tests switch kernel families with it)
int rows_min_channels() {
    const char* e = getenv("MUA_ROWS_MIN_C");
    return e ? atoi(e) : 8 * 32 * sm_count();
}

int check_layout(const uint8_t* d_sym, const int64_t* d_off, const int32_t* d_len, int64_t stride, int32_t T, int32_t C) {
    REQUIRE(C >= 0, "C < 0");
    REQUIRE(d_sym != nullptr || C == 0, "d_sym is NULL");      // a recording without channels has no buffer
    REQUIRE(aligned16(d_sym), "d_sym must be 16-byte aligned");
    REQUIRE(T >= 0 && T <= (1 << 28), "T out of range [0, 2^28]");
    if (!d_off) REQUIRE(stride % 16 == 0 && stride >= (int64_t)((T + 15) & ~15), "stride must be a multiple of 16 and >= round_up(T,16)");
    (void)d_len;
    return MUA_OK;
}

void layout_sizes(int S, int K, int Lmax, TabHdr* h) {
    h->S = S;
    h->K = K;
    h->Lmax = Lmax;
    h->nsym = dec_nsym(S, K, Lmax);
    h->W = h->nsym * Lmax;
    h->enc1_off = align_up((int)sizeof(TabHdr), 512);
    h->enc2_off = align_up(h->enc1_off + S * K * 16 * 4, 512);
    int next = align_up(h->enc2_off + S * K * 256 * 4, 512);
    h->enc4_off = 0;
    if (Lmax <= 2 && S <= 3) {
        h->enc4_off = next;
        next = align_up(next + S * K * 768, 512);
    }
    h->encp_off = 0;
    if (Lmax <= 8) {
        h->encp_off = next;
        next = align_up(next + S * K * 512, 512);
    }
    h->dec_off = next;
    next = align_up(h->dec_off + ((S * K) << h->W) * 4, 512);
    h->Wv = decv_window(K, Lmax);
    h->decv_off = next;
    h->total_bytes = align_up(h->decv_off + (K << h->Wv) * 4, 512);
}

template <int SS, bool MULTI>
void launch_calibrate(const CalibParams& P, cudaStream_t st) {
    const int grid = (P.L.C + CAL_WARPS - 1) / CAL_WARPS;
    k_calibrate<SS, MULTI><<<grid, CAL_WARPS * 32, 0, st>>>(P);
}

template <int S>
void launch_calibrate_head(const CalibParams& P, cudaStream_t st) {
    if (P.H[0] <= 128) {
        const long long threads = (long long)P.L.C * 4;
        k_calibrate_head<S, 4><<<(unsigned)((threads + CALH_THREADS - 1) / CALH_THREADS), CALH_THREADS, 0, st>>>(P);
    } else {
        const long long threads = (long long)P.L.C * 32;
        k_calibrate_head<S, 32><<<(unsigned)((threads + CALH_THREADS - 1) / CALH_THREADS), CALH_THREADS, 0, st>>>(P);
    }
}

template <int S>
int launch_calibrate_rows(const CalibParams& P, cudaStream_t st) {
    CalRowsParams PR;
    memset(&PR, 0, sizeof(PR));
    PR.C = P;
    const int T = P.L.T, nH = P.nH;
    // the boundaries of every channel (all rows are T bins long), sorted by position
    struct Ev { int pos, h, is_end; } ev[2 * MUA_MAX_H];
    int nev = 0;
    for (int h = 0; h < nH; ++h) {
        const int cut = T < (P.H[h] > 1 ? P.H[h] : 1) ? T : (P.H[h] > 1 ? P.H[h] : 1);   // functions_1.py:59-68
        int end = 0;
        if (P.mode == MUA_WINDOW_SKIP) {
            end = cut + T / 2;                                                           // get_BR_no_sort.py:178-183
            if (end > T) end = -1;
        } else if (P.mode == MUA_WINDOW_TRUNCATE) {
            end = cut + T / 2 < T ? cut + T / 2 : T;                                     // test_chosen_system.py:99-103
        }
        const bool has_post = end > 0 && P.mode != MUA_WINDOW_NONE && P.need_post;
        PR.cutv[h] = cut;
        PR.endv[h] = P.mode == MUA_WINDOW_NONE ? cut : end;
        ev[nev++] = Ev{cut, h, 0};
        if (has_post) ev[nev++] = Ev{end, h, 1};
    }
    for (int i = 1; i < nev; ++i)   // insertion sort
        for (int j = i; j > 0 && ev[j].pos < ev[j - 1].pos; --j) {
            const Ev t = ev[j]; ev[j] = ev[j - 1]; ev[j - 1] = t;
        }
    PR.nev = nev;
    for (int h = 0; h < MUA_MAX_H; ++h) PR.ev_cut[h] = 0, PR.ev_end[h] = -1;
    for (int i = 0; i < nev; ++i) {
        PR.ev_pos[i] = ev[i].pos;
        if (ev[i].is_end) PR.ev_end[ev[i].h] = i; else PR.ev_cut[ev[i].h] = i;
    }
    PR.snap_bytes = nev * (S - 1) * 64;
    constexpr int stage = cr_box(S) * 32 * cr_nst(S);   // all stages of a warp
    int warps = (227 * 1024 - CalRowsSmem::tail(cr_nst(S))) / (stage + PR.snap_bytes);
    if (warps > cr_max_warps(S)) warps = cr_max_warps(S);
    REQUIRE(warps >= 1, "calibrate: too many boundaries for shared memory");
    PR.warps = warps;
    const cuuint64_t gdim[2] = {(cuuint64_t)T, (cuuint64_t)P.L.C};
    const cuuint64_t gstr[1] = {(cuuint64_t)P.L.stride};
    const cuuint32_t box[2] = {(cuuint32_t)cr_box(S), 32}, estr[2] = {1, 1};
    const CUresult r = tensor_map_encoder()(&PR.tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(P.L.sym), gdim, gstr, box, estr,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, cr_box(S) == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    const long long nblk = ((long long)P.L.C + 31) / 32;
    const int grid = (int)(nblk < sm_count() ? nblk : sm_count());
    const long long per_sm = (nblk + grid - 1) / grid, rounds = (per_sm + warps - 1) / warps;
    PR.wuse = (int32_t)((per_sm + rounds - 1) / rounds);
    const int smem = warps * (stage + PR.snap_bytes) + CalRowsSmem::tail(cr_nst(S));
    cudaError_t e = cudaFuncSetAttribute(k_calibrate_rows<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return cuda_fail(e, "calibrate smem attribute");
    k_calibrate_rows<S><<<grid, warps * 32, smem, st>>>(PR);
    CHECK_LAUNCH("k_calibrate_rows");
    return MUA_OK;
}

// one alphabet size: the scan counts exactly its thresholds; several: one scan with the thresholds of the largest
int dispatch_calibrate(const CalibParams& P, cudaStream_t st) {
    if (P.L.C == 0) return MUA_OK;
    if (P.nS == 1 && P.nH == 1 && !P.train && !P.need_post && P.H[0] <= 1024) {
        // calibration window only (the streaming system): a few lanes per channel
        switch (P.out[0].S) {
            case 2: launch_calibrate_head<2>(P, st); break;
            case 3: launch_calibrate_head<3>(P, st); break;
            case 4: launch_calibrate_head<4>(P, st); break;
            case 5: launch_calibrate_head<5>(P, st); break;
            case 6: launch_calibrate_head<6>(P, st); break;
            case 7: launch_calibrate_head<7>(P, st); break;
            case 8: launch_calibrate_head<8>(P, st); break;
            case 9: launch_calibrate_head<9>(P, st); break;
            case 10: launch_calibrate_head<10>(P, st); break;
            default: return fail(MUA_E_INVALID, "S=%d outside 2..10", P.out[0].S);
        }
        CHECK_LAUNCH("k_calibrate_head");
        return MUA_OK;
    }
    if (P.nS == 1 && !P.train && !P.L.off && !P.L.len && P.L.T > 0 && P.L.T <= (rows_t_max() < 16384 ? rows_t_max() : 16384) &&
        P.L.C >= rows_min_channels() && tensor_map_encoder() != nullptr) {
        // many short rows of one length: a lane per channel (k_calibrate_rows); all boundaries are the same bins for every channel
        switch (P.out[0].S) {
            case 2: return launch_calibrate_rows<2>(P, st);
            case 3: return launch_calibrate_rows<3>(P, st);
            case 4: return launch_calibrate_rows<4>(P, st);
            case 5: return launch_calibrate_rows<5>(P, st);
            case 6: return launch_calibrate_rows<6>(P, st);
            case 7: return launch_calibrate_rows<7>(P, st);
            case 8: return launch_calibrate_rows<8>(P, st);
            case 9: return launch_calibrate_rows<9>(P, st);
            case 10: return launch_calibrate_rows<10>(P, st);
            default: return fail(MUA_E_INVALID, "S=%d outside 2..10", P.out[0].S);
        }
    }
    if (P.nS == 1) {
        switch (P.out[0].S) {
            case 2: launch_calibrate<2, false>(P, st); break;
            case 3: launch_calibrate<3, false>(P, st); break;
            case 4: launch_calibrate<4, false>(P, st); break;
            case 5: launch_calibrate<5, false>(P, st); break;
            case 6: launch_calibrate<6, false>(P, st); break;
            case 7: launch_calibrate<7, false>(P, st); break;
            case 8: launch_calibrate<8, false>(P, st); break;
            case 9: launch_calibrate<9, false>(P, st); break;
            case 10: launch_calibrate<10, false>(P, st); break;
            default: return fail(MUA_E_INVALID, "S=%d outside 2..10", P.out[0].S);
        }
    } else {
        int smax = 0;
        for (int i = 0; i < P.nS; ++i) smax = P.out[i].S > smax ? P.out[i].S : smax;
        if (smax <= 6) launch_calibrate<6, true>(P, st);
        else launch_calibrate<10, true>(P, st);
    }
    CHECK_LAUNCH("k_calibrate");
    return MUA_OK;
}

}  // namespace

extern "C" {

int mua_abi_version(void) { return MUA_ABI_VERSION; }
const char* mua_last_error(void) { return g_err; }

int mua_canonical_codebook(const uint8_t* h_lens, int K, int S, uint16_t* h_codes_out) {
    REQUIRE(h_lens && h_codes_out, "NULL argument");
    REQUIRE(S >= 2 && S <= MUA_MAX_S && K >= 1 && K <= MUA_MAX_K, "S/K out of range");
    for (int k = 0; k < K; ++k) {
        const uint8_t* L = h_lens + (size_t)k * S;
        uint32_t code = 0;
        for (int r = 0; r < S; ++r) {
            REQUIRE(L[r] >= 1 && L[r] <= 15, "length out of range");
            if (r) {
                REQUIRE(L[r] >= L[r - 1], "rows must be ascending (SCLV)");
                code = (code + 1) << (L[r] - L[r - 1]);
            }
            REQUIRE(code < (1u << L[r]), "row %d is not a valid prefix code (Kraft sum > 1)", k);
            h_codes_out[(size_t)k * S + r] = (uint16_t)code;
        }
    }
    return MUA_OK;
}

size_t mua_tables_bytes(int S, int K) {
    if (S < 2 || S > MUA_MAX_S || K < 1 || K > MUA_MAX_K) return 0;
    int worst = 0;   // the size must not depend on the rows: take the maximum over Lmax
    for (int L = 1; L <= 9; ++L) {
        TabHdr h;
        layout_sizes(S, K, L, &h);
        if (h.total_bytes > worst) worst = h.total_bytes;
    }
    return (size_t)worst;
}

int mua_build_tables(void* d_tables, const uint8_t* h_lens, const uint16_t* h_codes, int S, int K, void* stream) {
    REQUIRE(d_tables && h_lens && h_codes, "NULL argument");
    REQUIRE(aligned16(d_tables), "d_tables must be 16-byte aligned");
    REQUIRE(S >= 2 && S <= MUA_MAX_S && K >= 1 && K <= MUA_MAX_K, "S/K out of range");
    cudaStream_t st = (cudaStream_t)stream;
    TabHdr h;
    memset(&h, 0, sizeof(h));
    int Lmax = 0;
    for (int k = 0; k < K; ++k) {
        double kraft = 0;
        for (int r = 0; r < S; ++r) {
            int L = h_lens[(size_t)k * S + r];
            REQUIRE(L >= 1 && L <= 9, "codeword length %d outside 1..9", L);
            REQUIRE(h_codes[(size_t)k * S + r] < (1u << L), "codeword wider than its length");
            h.lens[k][r] = (uint8_t)L;
            h.codes[k][r] = h_codes[(size_t)k * S + r];
            kraft += 1.0 / (double)(1u << L);
            if (L > Lmax) Lmax = L;
        }
        REQUIRE(kraft == 1.0, "row %d is not Kraft-complete (decode LUT needs a complete prefix code)", k);
        for (int a = 0; a < S; ++a)
            for (int b = 0; b < S; ++b) {
                if (a == b) continue;
                int la = h.lens[k][a], lb = h.lens[k][b];
                if (la <= lb) REQUIRE((h.codes[k][b] >> (lb - la)) != h.codes[k][a], "row %d is not prefix-free", k);
            }
    }
    layout_sizes(S, K, Lmax, &h);
    cudaError_t e = cudaMemcpyAsync(d_tables, &h, sizeof(h), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return cuda_fail(e, "upload table header");
    k_build_tables<<<S * K, 256, 0, st>>>(reinterpret_cast<uint8_t*>(d_tables));
    CHECK_LAUNCH("k_build_tables");
    // the header lives on this stack frame: make sure the (pageable) upload has been consumed
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return cuda_fail(e, "build tables");
    return MUA_OK;
}

int mua_bin_raster(const void* d_raster, int dtype, int64_t T0, int32_t C, int32_t bin_res, int64_t* d_counts, uint8_t* d_sym,
                   int64_t sym_stride, int32_t S, void* stream) {
    REQUIRE(d_raster, "d_raster is NULL");
    REQUIRE(T0 >= 0 && C >= 0 && bin_res >= 1, "bad T0/C/bin_res");
    REQUIRE(d_counts || d_sym, "no output requested");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t nb = (T0 + bin_res - 1) / bin_res;
    if (nb == 0 || C == 0) return MUA_OK;
    if (d_counts) {
        dim3 grid((C + 255) / 256, (unsigned)(nb < 32768 ? nb : 32768));
        switch (dtype) {
            case MUA_DT_U8:
                if (C % 4 == 0 && (reinterpret_cast<uintptr_t>(d_raster) & 3) == 0 && (reinterpret_cast<uintptr_t>(d_counts) & 15) == 0) {
                    dim3 grid4((C / 4 + 255) / 256, (unsigned)(nb < 32768 ? nb : 32768));
                    k_bin_counts_u8x4<<<grid4, 256, 0, st>>>((const uint8_t*)d_raster, T0, C, bin_res, nb, d_counts);
                } else {
                    k_bin_counts<uint8_t, unsigned long long><<<grid, 256, 0, st>>>((const uint8_t*)d_raster, T0, C, bin_res, nb, d_counts);
                }
                break;
            case MUA_DT_I32: k_bin_counts<int32_t, long long><<<grid, 256, 0, st>>>((const int32_t*)d_raster, T0, C, bin_res, nb, d_counts); break;
            case MUA_DT_I64: k_bin_counts<int64_t, long long><<<grid, 256, 0, st>>>((const int64_t*)d_raster, T0, C, bin_res, nb, d_counts); break;
            case MUA_DT_F32: k_bin_counts<float, float><<<grid, 256, 0, st>>>((const float*)d_raster, T0, C, bin_res, nb, d_counts); break;
            case MUA_DT_F64: k_bin_counts<double, double><<<grid, 256, 0, st>>>((const double*)d_raster, T0, C, bin_res, nb, d_counts); break;
            default: return fail(MUA_E_INVALID, "unknown dtype %d", dtype);
        }
        CHECK_LAUNCH("k_bin_counts");
    }
    if (d_sym) {
        REQUIRE(dtype == MUA_DT_U8, "symbol output needs a uint8 raster");
        REQUIRE(S == 0 || (S >= 2 && S <= MUA_MAX_S), "S outside {0, 2..10}");
        REQUIRE(sym_stride >= nb, "sym_stride < number of bins");
        REQUIRE(nb <= (int64_t)65535 * BIN_TB, "too many bins for one launch");
        dim3 grid((C + BIN_TC - 1) / BIN_TC, (unsigned)((nb + BIN_TB - 1) / BIN_TB));
        REQUIRE(aligned16(d_sym) || sym_stride % 16 != 0, "d_sym must be 16-byte aligned when sym_stride is a multiple of 16");
        if (C % 8 == 0 && bin_res <= 128 && (reinterpret_cast<uintptr_t>(d_raster) & 7) == 0 && nb <= (int64_t)65535 * BW_TB) {
            dim3 gridw((C + BW_TC - 1) / BW_TC, (unsigned)((nb + BW_TB - 1) / BW_TB));
            k_bin_sym_wide<<<gridw, 256, 0, st>>>((const uint8_t*)d_raster, T0, C, bin_res, nb, d_sym, sym_stride, S ? S - 1 : 255);
        } else {
            k_bin_sym<<<grid, 256, 0, st>>>((const uint8_t*)d_raster, T0, C, bin_res, nb, d_sym, sym_stride, S ? S - 1 : 255);
        }
        CHECK_LAUNCH("k_bin_sym");
    }
    return MUA_OK;
}

int mua_bin_events(const double* d_times, const int32_t* d_chan, int64_t N, double t0, double w, int64_t nb, int32_t C,
                   uint8_t* d_sym, int64_t sym_stride, int32_t S, void* stream) {
    REQUIRE(N >= 0 && nb >= 0 && C >= 0, "bad N/nb/C");
    REQUIRE(w > 0.0 && t0 == t0, "bin width must be positive and t0 a number");
    REQUIRE(S == 0 || (S >= 2 && S <= MUA_MAX_S), "S outside {0, 2..10}");
    if (nb == 0 || C == 0) return MUA_OK;
    REQUIRE(d_sym, "d_sym is NULL");
    REQUIRE(sym_stride >= nb && sym_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(d_sym) & 3) == 0,
            "sym_stride must be >= nb and a multiple of 4, d_sym 4-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(d_sym, 0, (size_t)C * (size_t)sym_stride, st);
    if (e != cudaSuccess) return cuda_fail(e, "zero the symbol buffer");
    if (N == 0) return MUA_OK;
    REQUIRE(d_times && d_chan, "NULL argument");
    const long long blocks = (N + 255) / 256, cap = (long long)sm_count() * 16;
    k_bin_events<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, st>>>(d_times, d_chan, N, t0, w, nb, C, d_sym, sym_stride,
                                                                      S ? (uint32_t)(S - 1) : 255u);
    CHECK_LAUNCH("k_bin_events");
    return MUA_OK;
}

int mua_calibrate(const uint8_t* d_sym, const int64_t* d_off, const int32_t* d_len, int64_t stride, int32_t T, int32_t C, int32_t S,
                  const int32_t* h_H, int32_t nH, int32_t use_sort, int32_t window_mode, const void* d_tables, uint32_t active_lo,
                  uint32_t active_hi, int32_t* d_cutoff, int32_t* d_end, uint8_t* d_peak, uint8_t* d_enc, int32_t* d_assign_m,
                  int32_t* d_post_m, int64_t* d_bits, int64_t* d_nsym, void* stream) {
    mua_calib_out o;
    memset(&o, 0, sizeof(o));
    o.S = S; o.d_tables = d_tables; o.active_lo = active_lo; o.active_hi = active_hi;
    o.d_cutoff = d_cutoff; o.d_end = d_end; o.d_peak = d_peak; o.d_enc = d_enc;
    o.d_assign_m = d_assign_m; o.d_post_m = d_post_m; o.d_bits = d_bits; o.d_nsym = d_nsym;
    return mua_calibrate_multi(d_sym, d_off, d_len, stride, T, C, h_H, nH, use_sort, window_mode, &o, 1, stream);
}

int mua_calibrate_multi(const uint8_t* d_sym, const int64_t* d_off, const int32_t* d_len, int64_t stride, int32_t T, int32_t C,
                        const int32_t* h_H, int32_t nH, int32_t use_sort, int32_t window_mode, const mua_calib_out* h_outs,
                        int32_t nS, void* stream) {
    int rc = check_layout(d_sym, d_off, d_len, stride, T, C);
    if (rc) return rc;
    REQUIRE(h_H && nH >= 1 && nH <= MUA_MAX_H, "nH outside 1..%d", MUA_MAX_H);
    REQUIRE(window_mode >= 0 && window_mode <= 2, "bad window_mode");
    REQUIRE(h_outs && nS >= 1 && nS <= CAL_MAX_NS, "nS outside 1..%d", CAL_MAX_NS);
    CalibParams P;
    memset(&P, 0, sizeof(P));
    P.L = Layout{d_sym, d_off, d_len, stride, T, C};
    P.nH = nH; P.use_sort = use_sort; P.mode = window_mode; P.train = 0; P.nS = nS; P.one = 1;
    for (int i = 0; i < nH; ++i) P.H[i] = h_H[i];
    for (int i = 0; i < nS; ++i) {
        const mua_calib_out& o = h_outs[i];
        REQUIRE(o.S >= 2 && o.S <= MUA_MAX_S, "S=%d outside 2..10", o.S);
        REQUIRE(o.d_tables, "d_tables is NULL");
        CalOut& q = P.out[i];
        q.S = o.S;
        q.tab = reinterpret_cast<const TabHdr*>(o.d_tables);
        q.active = ((unsigned long long)o.active_hi << 32) | o.active_lo;
        REQUIRE(q.active != 0, "no active SCLV row");
        q.cutoff = o.d_cutoff; q.end = o.d_end; q.peak = o.d_peak; q.enc = o.d_enc;
        q.assign_m = o.d_assign_m; q.post_m = o.d_post_m; q.bits = o.d_bits; q.nsym = o.d_nsym;
        if (o.d_post_m || o.d_bits || o.d_nsym) P.need_post = 1;
    }
    return dispatch_calibrate(P, (cudaStream_t)stream);
}

int mua_train_hist(const uint8_t* d_sym, const int64_t* d_off, const int32_t* d_len, int64_t stride, int32_t T, int32_t C, int32_t S,
                   int32_t* d_hist_sorted, void* stream) {
    mua_calib_out o;
    memset(&o, 0, sizeof(o));
    o.S = S; o.d_train_hist = d_hist_sorted;
    return mua_train_hist_multi(d_sym, d_off, d_len, stride, T, C, &o, 1, stream);
}

int mua_train_hist_multi(const uint8_t* d_sym, const int64_t* d_off, const int32_t* d_len, int64_t stride, int32_t T, int32_t C,
                         const mua_calib_out* h_outs, int32_t nS, void* stream) {
    int rc = check_layout(d_sym, d_off, d_len, stride, T, C);
    if (rc) return rc;
    REQUIRE(h_outs && nS >= 1 && nS <= CAL_MAX_NS, "nS outside 1..%d", CAL_MAX_NS);
    CalibParams P;
    memset(&P, 0, sizeof(P));
    P.L = Layout{d_sym, d_off, d_len, stride, T, C};
    P.nH = 1; P.use_sort = 0; P.mode = MUA_WINDOW_NONE; P.train = 1; P.nS = nS; P.one = 1;
    P.H[0] = 0x7FFFFFFF;
    for (int i = 0; i < nS; ++i) {
        REQUIRE(h_outs[i].S >= 2 && h_outs[i].S <= MUA_MAX_S, "S=%d outside 2..10", h_outs[i].S);
        REQUIRE(h_outs[i].d_train_hist || C == 0, "d_train_hist is NULL");
        P.out[i].S = h_outs[i].S;
        P.out[i].active = 1;
        P.out[i].train_hist = h_outs[i].d_train_hist;
    }
    return dispatch_calibrate(P, (cudaStream_t)stream);
}

int mua_select_sclv(const int32_t* d_hist, int64_t N, const void* d_tables, uint32_t active_lo, uint32_t active_hi, uint8_t* d_enc,
                    int64_t* d_min1, int64_t* d_min2, void* stream) {
    REQUIRE(d_hist && d_tables && d_enc, "NULL argument");
    REQUIRE(N >= 0, "N < 0");
    const unsigned long long active = ((unsigned long long)active_hi << 32) | active_lo;
    REQUIRE(active != 0, "no active SCLV row");
    if (N == 0) return MUA_OK;
    const int grid = (int)((N + 255) / 256 < (int64_t)sm_count() * 8 ? (N + 255) / 256 : (int64_t)sm_count() * 8);
    k_select<<<grid, 256, 0, (cudaStream_t)stream>>>(d_hist, N, reinterpret_cast<const TabHdr*>(d_tables), active, d_enc, d_min1, d_min2);
    CHECK_LAUNCH("k_select");
    return MUA_OK;
}

int mua_bit_counts(const int32_t* d_hist, const uint8_t* d_enc, int64_t N, const void* d_tables, int64_t* d_bits, int64_t* d_nsym,
                   void* stream) {
    REQUIRE(d_hist && d_enc && d_tables && d_bits && d_nsym, "NULL argument");
    REQUIRE(N >= 0, "N < 0");
    if (N == 0) return MUA_OK;
    const int grid = (int)((N + 255) / 256 < (int64_t)sm_count() * 8 ? (N + 255) / 256 : (int64_t)sm_count() * 8);
    k_bit_counts<<<grid, 256, 0, (cudaStream_t)stream>>>(d_hist, d_enc, N, reinterpret_cast<const TabHdr*>(d_tables), d_bits, d_nsym);
    CHECK_LAUNCH("k_bit_counts");
    return MUA_OK;
}

int mua_elim_scores(const uint8_t* d_enc, const int64_t* d_min1, const int64_t* d_min2, int64_t N, int32_t K, int64_t* d_assign_hist,
                    int64_t* d_score, void* stream) {
    REQUIRE(d_enc && d_min1 && d_min2 && d_assign_hist && d_score, "NULL argument");
    REQUIRE(K >= 1 && K <= MUA_MAX_K && N >= 0, "bad K/N");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(d_assign_hist, 0, sizeof(int64_t) * K, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_score, 0, sizeof(int64_t) * K, st);
    if (e != cudaSuccess) return cuda_fail(e, "memset");
    if (N == 0) return MUA_OK;
    const int grid = (int)((N + 255) / 256 < (int64_t)sm_count() * 4 ? (N + 255) / 256 : (int64_t)sm_count() * 4);
    k_elim_scores<<<grid, 256, 0, st>>>(d_enc, d_min1, d_min2, N, K, reinterpret_cast<unsigned long long*>(d_assign_hist),
                                        reinterpret_cast<unsigned long long*>(d_score));
    CHECK_LAUNCH("k_elim_scores");
    return MUA_OK;
}

int mua_peer_alloc(size_t bytes, void** d_ptr, uint8_t* h_handle) {
    REQUIRE(d_ptr && h_handle && bytes > 0, "NULL argument / zero size");
    static_assert(sizeof(cudaIpcMemHandle_t) == MUA_IPC_HANDLE_BYTES, "IPC handle size");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc (peer buffer)");
    e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return cuda_fail(e, "cudaIpcGetMemHandle");
    }
    memcpy(h_handle, &h, sizeof(h));
    *d_ptr = p;
    return MUA_OK;
}

int mua_peer_open(const uint8_t* h_handle, void** d_ptr) {
    REQUIRE(d_ptr && h_handle, "NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handle, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return cuda_fail(e, "cudaIpcOpenMemHandle");
    *d_ptr = p;
    return MUA_OK;
}

int mua_peer_close(void* d_ptr) {
    if (!d_ptr) return MUA_OK;
    cudaError_t e = cudaIpcCloseMemHandle(d_ptr);
    return e == cudaSuccess ? MUA_OK : cuda_fail(e, "cudaIpcCloseMemHandle");
}

int mua_peer_free(void* d_ptr) {
    if (!d_ptr) return MUA_OK;
    cudaError_t e = cudaFree(d_ptr);
    return e == cudaSuccess ? MUA_OK : cuda_fail(e, "cudaFree (peer buffer)");
}

static int check_sink(const mua_report_sink* s) {
    REQUIRE(s, "report sink is NULL");
    REQUIRE(s->n_peers >= 1 && s->n_peers <= MUA_MAX_PEERS && s->rank >= 0 && s->rank < s->n_peers, "bad report sink ranks");
    for (int i = 0; i < s->n_peers; ++i) REQUIRE(s->d_flags[i], "report sink: d_flags[%d] is NULL", i);
    return MUA_OK;
}

int mua_report_signal(const mua_report_sink* h_sink, int32_t step, void* stream) {
    int rc = check_sink(h_sink);
    if (rc) return rc;
    PeerFlags F;
    F.n = h_sink->n_peers; F.rank = h_sink->rank;
    for (int i = 0; i < MUA_MAX_PEERS; ++i) F.flags[i] = i < F.n ? h_sink->d_flags[i] : nullptr;
    k_report_signal<<<1, 32, 0, (cudaStream_t)stream>>>(F, step);
    CHECK_LAUNCH("k_report_signal");
    return MUA_OK;
}

int mua_report_wait(const mua_report_sink* h_sink, int32_t step, void* stream) {
    int rc = check_sink(h_sink);
    if (rc) return rc;
    PeerFlags F;
    F.n = h_sink->n_peers; F.rank = h_sink->rank;
    for (int i = 0; i < MUA_MAX_PEERS; ++i) F.flags[i] = i < F.n ? h_sink->d_flags[i] : nullptr;
    k_report_wait<<<1, 32, 0, (cudaStream_t)stream>>>(F, step, 2000000000ll);   // give up after 2 s (%globaltimer, ns)
    CHECK_LAUNCH("k_report_wait");
    return MUA_OK;
}

int mua_encode(const uint8_t* d_sym, const int64_t* d_off, const int32_t* d_len, int64_t stride, int32_t T, int32_t C, int32_t S,
               const int32_t* d_start, const int32_t* d_end, const uint8_t* d_peak, const uint8_t* d_enc, const void* d_tables,
               int32_t K, int32_t Lmax, uint8_t* d_stream, int64_t slot_bytes, uint32_t* d_chunk_off, int32_t chunk_stride,
               uint32_t* d_sub_off, int32_t sub_stride, int64_t* d_total_bits, int32_t* d_overflow, const mua_report_sink* h_sink,
               void* stream) {
    int rc = check_layout(d_sym, d_off, d_len, stride, T, C);
    if (rc) return rc;
    if (C == 0) return MUA_OK;                                  // nothing to encode: per-channel arrays may be empty (NULL)
    REQUIRE(d_start && d_end && d_peak && d_enc && d_tables && d_stream && d_chunk_off && d_total_bits && d_overflow, "NULL argument");
    REQUIRE(aligned16(d_stream) && slot_bytes > 0 && slot_bytes % 16 == 0, "d_stream/slot_bytes must be 16-byte aligned");
    REQUIRE(chunk_stride >= (T + TILE - 1) / TILE && chunk_stride >= 1, "chunk_stride < ceil(T/%d)", TILE);
    REQUIRE(!d_sub_off || sub_stride >= 8 * ((T + TILE - 1) / TILE), "sub_stride < 8 * ceil(T/%d)", TILE);
    REQUIRE((long long)C * chunk_stride < (1ll << 29) && (!d_sub_off || (long long)C * sub_stride < (1ll << 29)),
            "side-info arrays must be smaller than 2 GiB");
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 0) return MUA_OK;
    REQUIRE(S >= 2 && S <= MUA_MAX_S && K >= 1 && K <= MUA_MAX_K && Lmax >= 1 && Lmax <= 9, "bad S/K/Lmax");
    TabHdr h;
    layout_sizes(S, K, Lmax, &h);
    EncParams P;
    P.L = Layout{d_sym, d_off, d_len, stride, T, C};
    P.S = S; P.start = d_start; P.end = d_end; P.peak = d_peak; P.enc = d_enc;
    P.tab = reinterpret_cast<const uint8_t*>(d_tables); P.K = K; P.Lmax = Lmax;
    P.stream = d_stream; P.slot_bytes = slot_bytes; P.chunk_off = d_chunk_off; P.chunk_stride = chunk_stride;
    P.sub_off = d_sub_off; P.sub_stride = sub_stride;
    P.total_bits = d_total_bits; P.overflow = d_overflow;
    P.n_peers = 0; P.row0 = 0; P.signal_step = 0; P.rank = 0;
    for (int i = 0; i < MUA_MAX_PEERS; ++i) { P.rep[i] = nullptr; P.flags[i] = nullptr; }
    if (h_sink && h_sink->n_peers > 0) {
        REQUIRE(h_sink->n_peers <= MUA_MAX_PEERS && h_sink->row0 >= 0, "bad report sink");
        REQUIRE((long long)T * Lmax < (1ll << 31), "report sink rows are int32: T * Lmax must be < 2^31");
        for (int i = 0; i < h_sink->n_peers; ++i) {
            REQUIRE(h_sink->d_report[i] && aligned16(h_sink->d_report[i]), "report sink: d_report[%d] NULL or not 16-byte aligned", i);
            P.rep[i] = h_sink->d_report[i];
        }
        P.n_peers = h_sink->n_peers;
        P.row0 = h_sink->row0;
        if (h_sink->signal_step > 0) {
            REQUIRE(h_sink->rank >= 0 && h_sink->rank < h_sink->n_peers, "report sink: bad rank");
            for (int i = 0; i < h_sink->n_peers; ++i) {
                REQUIRE(h_sink->d_flags[i], "report sink: d_flags[%d] is NULL", i);
                P.flags[i] = h_sink->d_flags[i];
            }
            P.signal_step = h_sink->signal_step;
            P.rank = h_sink->rank;
        }
    }
    const int ctas_needed = (C + ENC_WARPS - 1) / ENC_WARPS;
    if (h.Lmax <= 2 && S <= 3 && T <= rows_t_max() && C >= rows_min_channels() && S * K * EF_LUT_B <= ER_LUT_MAX) {
        // short rows: a lane per channel.  All 32-channel blocks of a wave are resident at once; the warps are spread evenly
        // over the waves (a block is a long task: an extra, nearly empty wave would cost as much as a full one)
        const int smem = EncRowsSmem::TOTAL;
        const bool fixed = !d_off && !d_len && T > 0 && tensor_map_encoder() != nullptr;
        EncRowsParams PR;
        PR.E = P;
        PR.zero = 0;
        memset(&PR.tmap, 0, sizeof(PR.tmap));
        if (fixed) {   // the recording as a 2-D tensor for the TMA engine: bin x channel, boxes of 128 bins x 32 channels
            const cuuint64_t gdim[2] = {(cuuint64_t)T, (cuuint64_t)C};
            const cuuint64_t gstr[1] = {(cuuint64_t)stride};
            const cuuint32_t box[2] = {128, 32}, estr[2] = {1, 1};
            const CUresult r = tensor_map_encoder()(&PR.tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(d_sym), gdim, gstr, box, estr,
                                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
        }
        const long long nblk = ((long long)C + 31) / 32;
        const int grid = (int)(nblk < sm_count() ? nblk : sm_count());
        const long long per_sm = (nblk + grid - 1) / grid, rounds = (per_sm + ER_WARPS - 1) / ER_WARPS;
        PR.wuse = (int32_t)((per_sm + rounds - 1) / rounds);
#define MUA_LAUNCH_ENCR(SV, FX)                                                                                         \
    do {                                                                                                                \
        cudaError_t e = cudaFuncSetAttribute(k_encode_rows<SV, FX>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
        if (e != cudaSuccess) return cuda_fail(e, "encode smem attribute");                                            \
        k_encode_rows<SV, FX><<<grid, ER_WARPS * 32, smem, st>>>(PR);                                                   \
    } while (0)
        if (S == 2) { if (fixed) MUA_LAUNCH_ENCR(2, true); else MUA_LAUNCH_ENCR(2, false); }
        else { if (fixed) MUA_LAUNCH_ENCR(3, true); else MUA_LAUNCH_ENCR(3, false); }
#undef MUA_LAUNCH_ENCR
    } else if (h.Lmax <= 2 && S <= 3) {
        const int smem = EncFastSmem::PER_WARP * EF_WARPS;
        const int per_sm = 4;
        const int need = (C + EF_WARPS - 1) / EF_WARPS;
        const int grid = need < sm_count() * per_sm ? need : sm_count() * per_sm;
        // fixed row stride, a multiple of 64: tiles arrive as TMA tensor boxes with the 64-byte swizzle (k_encode_fast<.., true>)
        const bool tensor = !d_off && stride % 64 == 0 && (long long)C * stride < (1ll << 37) && tensor_map_encoder() != nullptr &&
                            (reinterpret_cast<uintptr_t>(d_sym) & 63) == 0 && !getenv("MUA_ENC_NO_TENSOR");
        EncFastParams PF;
        PF.E = P;
        memset(&PF.tmap, 0, sizeof(PF.tmap));
        if (tensor) {
            const cuuint64_t gdim[2] = {64, (cuuint64_t)(((long long)C * stride + 63) / 64)};
            const cuuint64_t gstr[1] = {64};
            const cuuint32_t box[2] = {64, 32}, estr[2] = {1, 1};
            const CUresult r = tensor_map_encoder()(&PF.tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(d_sym), gdim, gstr, box, estr,
                                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
        }
#define MUA_LAUNCH_ENC(SV, TN)                                                                                          \
    do {                                                                                                                \
        cudaError_t e = cudaFuncSetAttribute(k_encode_fast<SV, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
        if (e != cudaSuccess) return cuda_fail(e, "encode smem attribute");                                            \
        k_encode_fast<SV, TN><<<grid, EF_WARPS * 32, smem, st>>>(PF);                                                   \
    } while (0)
        if (S == 2) { if (tensor) MUA_LAUNCH_ENC(2, true); else MUA_LAUNCH_ENC(2, false); }
        else { if (tensor) MUA_LAUNCH_ENC(3, true); else MUA_LAUNCH_ENC(3, false); }
#undef MUA_LAUNCH_ENC
    } else if (h.Lmax <= 8 && S >= 3 && S <= 9 && T > 0 && T <= rows_t_max() && C >= rows_min_channels() && !d_off && !d_len &&
               tensor_map_encoder() != nullptr && (227 * 1024 - S * K * 512) / (2 * ERP_STAGE + ERP_RING + 16) >= 8) {
        // many short rows, pair-table codebooks: a lane per channel (k_encode_rows_pair), all pair tables in shared memory
        EncRowsPairParams PR;
        PR.E = P;
        PR.zero = 0;
        PR.lut_bytes = S * K * 512;
        int warps = (227 * 1024 - PR.lut_bytes) / (2 * ERP_STAGE + ERP_RING + 16);
        if (warps > ER_WARPS) warps = ER_WARPS;
        PR.warps = warps;
        memset(&PR.tmap, 0, sizeof(PR.tmap));
        const cuuint64_t gdim[2] = {(cuuint64_t)T, (cuuint64_t)C};
        const cuuint64_t gstr[1] = {(cuuint64_t)stride};
        const cuuint32_t box[2] = {ER_TILE, 32}, estr[2] = {1, 1};
        const CUresult r = tensor_map_encoder()(&PR.tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(d_sym), gdim, gstr, box, estr,
                                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
        const long long nblk = ((long long)C + 31) / 32;
        const int grid = (int)(nblk < sm_count() ? nblk : sm_count());
        const long long per_sm = (nblk + grid - 1) / grid, rounds = (per_sm + warps - 1) / warps;
        PR.wuse = (int32_t)((per_sm + rounds - 1) / rounds);
        const int smem = warps * (2 * ERP_STAGE + ERP_RING + 16) + PR.lut_bytes;
#define MUA_LAUNCH_ENCRP(SV)                                                                                            \
    do {                                                                                                                \
        cudaError_t e = cudaFuncSetAttribute(k_encode_rows_pair<SV>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
        if (e != cudaSuccess) return cuda_fail(e, "encode smem attribute");                                            \
        k_encode_rows_pair<SV><<<grid, warps * 32, smem, st>>>(PR);                                                     \
    } while (0)
        switch (S) {
            case 3: MUA_LAUNCH_ENCRP(3); break;
            case 4: MUA_LAUNCH_ENCRP(4); break;
            case 5: MUA_LAUNCH_ENCRP(5); break;
            case 6: MUA_LAUNCH_ENCRP(6); break;
            case 7: MUA_LAUNCH_ENCRP(7); break;
            case 8: MUA_LAUNCH_ENCRP(8); break;
            default: MUA_LAUNCH_ENCRP(9); break;
        }
#undef MUA_LAUNCH_ENCRP
    } else if (h.Lmax <= 8 && S >= 3 && S <= 9) {
        const int smem = EncPairSmem::PER_WARP * ENC_WARPS;
        const int grid = ctas_needed < sm_count() * 3 ? ctas_needed : sm_count() * 3;
        const bool tensor = !d_off && stride % 64 == 0 && (long long)C * stride < (1ll << 37) && tensor_map_encoder() != nullptr &&
                            (reinterpret_cast<uintptr_t>(d_sym) & 63) == 0 && !getenv("MUA_ENC_NO_TENSOR");
        EncFastParams PF;
        PF.E = P;
        memset(&PF.tmap, 0, sizeof(PF.tmap));
        if (tensor) {   // the recording's bytes as 64-byte rows, one box of 64 x 32 per 2048-symbol tile (see k_encode_fast)
            const cuuint64_t gdim[2] = {64, (cuuint64_t)(((long long)C * stride + 63) / 64)};
            const cuuint64_t gstr[1] = {64};
            const cuuint32_t box[2] = {64, 32}, estr[2] = {1, 1};
            const CUresult r = tensor_map_encoder()(&PF.tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t*>(d_sym), gdim, gstr, box, estr,
                                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
        }
#define MUA_LAUNCH_ENCP1(SV, TN)                                                                                        \
    do {                                                                                                                \
        cudaError_t e = cudaFuncSetAttribute(k_encode_pair<SV, TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
        if (e != cudaSuccess) return cuda_fail(e, "encode smem attribute");                                            \
        k_encode_pair<SV, TN><<<grid, ENC_WARPS * 32, smem, st>>>(PF);                                                  \
    } while (0)
#define MUA_LAUNCH_ENCP(SV)                                \
    do {                                                   \
        if (tensor) MUA_LAUNCH_ENCP1(SV, true);            \
        else MUA_LAUNCH_ENCP1(SV, false);                  \
    } while (0)
        switch (S) {
            case 3: MUA_LAUNCH_ENCP(3); break;
            case 4: MUA_LAUNCH_ENCP(4); break;
            case 5: MUA_LAUNCH_ENCP(5); break;
            case 6: MUA_LAUNCH_ENCP(6); break;
            case 7: MUA_LAUNCH_ENCP(7); break;
            case 8: MUA_LAUNCH_ENCP(8); break;
            default: MUA_LAUNCH_ENCP(9); break;
        }
#undef MUA_LAUNCH_ENCP1
#undef MUA_LAUNCH_ENCP
    } else {
        const int smem = EncGenSmem::PER_WARP * ENC_WARPS;
        cudaError_t e = cudaFuncSetAttribute(k_encode_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return cuda_fail(e, "encode smem attribute");
        const int grid = ctas_needed < sm_count() * 3 ? ctas_needed : sm_count() * 3;
        k_encode_gen<<<grid, ENC_WARPS * 32, smem, st>>>(P);
    }
    CHECK_LAUNCH("k_encode");
    return MUA_OK;
}

int mua_pack_streams(const uint8_t* d_stream, int64_t slot_bytes, const int64_t* d_total_bits, int32_t C, int64_t* d_unit_off,
                     uint8_t* d_dense, int64_t dense_bytes, void* stream) {
    REQUIRE(C >= 0, "C < 0");
    REQUIRE(d_unit_off, "d_unit_off is NULL");
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 0) {
        cudaError_t e = cudaMemsetAsync(d_unit_off, 0, sizeof(int64_t), st);
        return e == cudaSuccess ? MUA_OK : cuda_fail(e, "memset");
    }
    REQUIRE(d_stream && d_total_bits && d_dense, "NULL argument");
    REQUIRE(aligned16(d_stream) && aligned16(d_dense) && slot_bytes > 0 && slot_bytes % 16 == 0 && dense_bytes >= 0,
            "d_stream/d_dense/slot_bytes must be 16-byte aligned");
    k_pack_offsets<<<1, 1024, 0, st>>>(d_total_bits, C, slot_bytes >> 4, d_unit_off);
    CHECK_LAUNCH("k_pack_offsets");
    const int need = (C + 7) / 8;
    const int grid = need < sm_count() * 8 ? need : sm_count() * 8;
    k_pack_copy<<<grid, 256, 0, st>>>(d_stream, slot_bytes, C, d_unit_off, d_dense, dense_bytes >> 4);
    CHECK_LAUNCH("k_pack_copy");
    return MUA_OK;
}

int mua_decode(const uint8_t* d_stream, int64_t slot_bytes, const uint32_t* d_chunk_off, int32_t chunk_stride,
               const uint32_t* d_sub_off, int32_t sub_stride, const int64_t* d_off,
               int64_t stride, int32_t C, int32_t S, const int32_t* d_start, const int32_t* d_end, const uint8_t* d_peak,
               const uint8_t* d_enc, const void* d_tables, int32_t K, int32_t Lmax, int32_t max_end, uint8_t* d_dec,
               int32_t* d_status, const mua_report_sink* h_wait_sink, int32_t wait_step, void* stream) {
    REQUIRE(C >= 0, "C < 0");
    if (C == 0) return MUA_OK;                                  // nothing to decode: per-channel arrays may be empty (NULL)
    REQUIRE(d_stream && d_chunk_off && d_start && d_end && d_peak && d_enc && d_tables && d_dec && d_status, "NULL argument");
    REQUIRE(aligned16(d_stream) && slot_bytes > 0 && slot_bytes % 16 == 0, "d_stream/slot_bytes must be 16-byte aligned");
    REQUIRE(aligned16(d_dec), "d_dec must be 16-byte aligned");
    REQUIRE(d_off || stride % 16 == 0, "stride must be a multiple of 16");
    REQUIRE(chunk_stride >= 1 && C >= 0, "bad chunk_stride/C");
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 0) return MUA_OK;
    REQUIRE(S >= 2 && S <= MUA_MAX_S && K >= 1 && K <= MUA_MAX_K && Lmax >= 1 && Lmax <= 9, "bad S/K/Lmax");
    TabHdr h;
    layout_sizes(S, K, Lmax, &h);
    DecParams P;
    P.stream = d_stream; P.slot_bytes = slot_bytes; P.chunk_off = d_chunk_off; P.chunk_stride = chunk_stride;
    P.off = d_off; P.stride = stride; P.C = C; P.S = S; P.start = d_start; P.end = d_end; P.peak = d_peak; P.enc = d_enc;
    P.tab = reinterpret_cast<const uint8_t*>(d_tables); P.K = K; P.Lmax = Lmax; P.dec = d_dec; P.status = d_status;
    P.sub_off = d_sub_off; P.sub_stride = sub_stride;
    P.wait_flags = nullptr; P.wait_n = 0; P.wait_step = 0;
    if (h_wait_sink && h_wait_sink->n_peers > 0 && wait_step > 0) {
        int rcs = check_sink(h_wait_sink);
        if (rcs) return rcs;
        P.wait_flags = h_wait_sink->d_flags[h_wait_sink->rank]; P.wait_n = h_wait_sink->n_peers; P.wait_step = wait_step;
    }
    REQUIRE(slot_bytes < (1ll << 32), "slot_bytes must be < 4 GiB");
    // chunks per channel that can be non-empty: ceil(max_end / CHUNK) when the caller bounds the window end
    P.item_chunks = chunk_stride;
    if (max_end > 0 && (max_end + TILE - 1) / TILE < chunk_stride) P.item_chunks = (max_end + TILE - 1) / TILE;
#ifndef MUA_DEC_BY_CHUNK_MAX
#define MUA_DEC_BY_CHUNK_MAX 4       // rows of at most this many chunks are decoded chunk-major (see dec_item)
#endif
    // (lane decoder only: 100k x 2 400 decodes in 57 instead of 68 us; the general decoder, whose stages are paced by the warp's
    // slowest lane either way, got slower with it: S=5 157 -> 181 us)
    P.by_chunk = h.nsym == 4 && h.Lmax <= 2 && h.K <= DL_MAX_ROWS && h.S <= 8 && h.W == 8 && P.item_chunks >= 2 &&
                 P.item_chunks <= MUA_DEC_BY_CHUNK_MAX;
    const long long nitems = P.by_chunk ? (long long)((C + 31) / 32) * 32 * P.item_chunks : (long long)C * P.item_chunks;
    const long long groups = (nitems + 31) / 32;
    const int lut_bytes = ((h.S * h.K) << h.W) * 4;
    const bool smem_lut = lut_bytes <= 32 * 1024;
    if (h.nsym == 4 && h.Lmax <= 2) {
        REQUIRE((long long)C * (slot_bytes >> 4) < (1ll << 32), "stream buffer must be < 64 GiB");
        if (h.K <= DL_MAX_ROWS && h.S <= 8 && h.W == 8) {   // lane-private LUT banks: one persistent CTA per SM
#ifndef MUA_DL_NC
#define MUA_DL_NC 1
#endif
#ifndef MUA_DV_EXTRA
#define MUA_DV_EXTRA 256      // bits a staged stream row of k_decode_var holds beyond one worst-case period
#endif
#ifndef MUA_DEC_SUB
#define MUA_DEC_SUB 1
#endif
#ifndef MUA_DEC_SUB_MIN_CHUNKS
#define MUA_DEC_SUB_MIN_CHUNKS 8     // shorter windows leave most of a 32-sub-chunk group idle: lane decoder
#endif
            // with the encoder's 128-symbol side info (codebooks of the fast encoder's class, fixed row stride, long windows): a lane per
            // sub-chunk, 32 consecutive sub-chunks per warp (k_decode_sub)
            const long long sgroups = (long long)C * ((8 * P.item_chunks + 31) / 32);
            if (MUA_DEC_SUB && d_sub_off && !d_off && h.S <= 3 && P.item_chunks >= MUA_DEC_SUB_MIN_CHUNKS && sgroups < (1ll << 31)) {
                REQUIRE(sub_stride >= 8 * P.item_chunks, "sub_stride < 8 * chunks per channel");
                DecSubParams PS;
                PS.D = P;
                memset(&PS.tmap, 0, sizeof(PS.tmap));
#if MUA_DS_TMA
                REQUIRE(tensor_map_encoder() != nullptr, "cuTensorMapEncodeTiled is not available");
                const cuuint64_t gdim[3] = {128, (cuuint64_t)((stride + 127) / 128), (cuuint64_t)C};
                const cuuint64_t gstr[2] = {128, (cuuint64_t)stride};
                const cuuint32_t box[3] = {128, 32, 1}, estr[3] = {1, 1, 1};
                const CUresult r = tensor_map_encoder()(&PS.tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d_dec, gdim, gstr, box, estr,
                                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
#endif
                cudaError_t e = cudaFuncSetAttribute(k_decode_sub, cudaFuncAttributeMaxDynamicSharedMemorySize, DL_SMEM);
                if (e != cudaSuccess) return cuda_fail(e, "decode smem attribute");
                const long long blocks_needed = (sgroups + DS_WARPS - 1) / DS_WARPS;
                const int grid = (int)(blocks_needed < sm_count() ? blocks_needed : sm_count());
                k_decode_sub<<<grid, DS_WARPS * 32, DL_SMEM, st>>>(PS);
                CHECK_LAUNCH("k_decode");
                return MUA_OK;
            }
            constexpr int NC = MUA_DL_NC;
            cudaError_t e = cudaFuncSetAttribute(k_decode_lane<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, DL_SMEM);
            if (e != cudaSuccess) return cuda_fail(e, "decode smem attribute");
            const long long lgroups = (nitems + 32 * NC - 1) / (32 * NC);
            const long long blocks_needed = (lgroups + DL_WARPS / NC - 1) / (DL_WARPS / NC);
            const int grid = (int)(blocks_needed < sm_count() ? blocks_needed : sm_count());
            k_decode_lane<NC><<<grid, DL_WARPS / NC * 32, DL_SMEM, st>>>(P);
            CHECK_LAUNCH("k_decode");
            return MUA_OK;
        }
        const bool fast_smem_lut = smem_lut;
        const int smem = DF_WARPS * DF_PER_WARP + (fast_smem_lut ? lut_bytes : 0);
        const long long blocks_needed = (groups + DF_WARPS - 1) / DF_WARPS;
        const long long cap = (long long)sm_count() * 4;
        const int grid = (int)(blocks_needed < cap ? blocks_needed : cap);
        if (fast_smem_lut) {
            cudaError_t e = cudaFuncSetAttribute(k_decode_fast<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return cuda_fail(e, "decode smem attribute");
            k_decode_fast<true><<<grid, DF_WARPS * 32, smem, st>>>(P);
        } else {
            cudaError_t e = cudaFuncSetAttribute(k_decode_fast<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return cuda_fail(e, "decode smem attribute");
            k_decode_fast<false><<<grid, DF_WARPS * 32, smem, st>>>(P);
        }
    } else if (P.item_chunks <= 2 && C >= rows_min_channels()) {
        // many rows whose windows span at most two chunks (the 2 400-bin rows of cfg4 at 50 ms): a lane per channel, the whole stream
        // of a channel decoded by its lane (k_decode_rows).  Its lookup loop costs ~16 instructions per symbol against ~8 of
        // k_decode_var, so it only pays where k_decode_var's lanes idle behind a 960- and a 240-symbol chunk per channel
        // (100k x 2 400: S=5 0.157 -> 0.13 ms, S=9 0.39 -> 0.18 ms; 100k x 12 000: 0.31 -> 0.60 ms, not taken)
        const int smem = DR_WARPS * DR_PER_WARP + 4 * MUA_MAX_S * 4 + h.K * ((1 << h.Wv) + DV_ROW_SKEW) * 4;
        const long long nblk = ((long long)C + 31) / 32;
        const int grid = (int)(nblk < sm_count() ? nblk : sm_count());
        const long long per_sm = (nblk + grid - 1) / grid, rounds = (per_sm + DR_WARPS - 1) / DR_WARPS;
        const int wuse = (int)((per_sm + rounds - 1) / rounds);
        if (h.S > 8) {
            cudaError_t e = cudaFuncSetAttribute(k_decode_rows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return cuda_fail(e, "decode smem attribute");
            k_decode_rows<true><<<grid, DR_WARPS * 32, smem, st>>>(P, wuse);
        } else {
            cudaError_t e = cudaFuncSetAttribute(k_decode_rows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return cuda_fail(e, "decode smem attribute");
            k_decode_rows<false><<<grid, DR_WARPS * 32, smem, st>>>(P, wuse);
        }
    } else {
        // variable-count lookups from per-row rank tables in shared memory: one persistent CTA per SM
        const int fixed = 4 * MUA_MAX_S * 4 + 16 + DV_LENS_B + h.K * ((1 << h.Wv) + DV_ROW_SKEW) * 4;      // rank maps, ticket, SCLV rows, tables
        // staged stream row per lane: alignment slack + one worst-case period + look-ahead + MUA_DV_EXTRA bits (a stage serves
        // periods for as long as a worst-case period still fits: ~MUA_DV_EXTRA / 130 more periods of typical MUA counts)
        P.var_pps = 1;
        P.var_str_w = ((127 + 96 + 128 * h.Lmax + MUA_DV_EXTRA + 31) / 32 + 3) / 4 * 4;     // whole 16-byte units
        if ((P.var_str_w / 4) % 2 == 0) P.var_str_w += 4;     // row stride = 16 B x odd: the lanes' refill words spread over 8 banks (a 128-byte stride puts all 32 lanes on one)
        const int DV_PER_WARP = 32 * P.var_str_w * 4 + 32 * DG_OUT_B + 16;
        int nw = (227 * 1024 - fixed) / DV_PER_WARP;
        nw = nw > DV_WARPS ? DV_WARPS : nw;
        REQUIRE(nw >= 1, "decode tables do not fit shared memory");
        const int smem = nw * DV_PER_WARP + fixed;
        const long long blocks_needed = (groups + nw - 1) / nw;
        const int grid = (int)(blocks_needed < sm_count() ? blocks_needed : sm_count());
        if (h.S > 8) {
            cudaError_t e = cudaFuncSetAttribute(k_decode_var<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return cuda_fail(e, "decode smem attribute");
            k_decode_var<true><<<grid, nw * 32, smem, st>>>(P);
        } else {
            cudaError_t e = cudaFuncSetAttribute(k_decode_var<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
            if (e != cudaSuccess) return cuda_fail(e, "decode smem attribute");
            k_decode_var<false><<<grid, nw * 32, smem, st>>>(P);
        }
    }
    CHECK_LAUNCH("k_decode");
    return MUA_OK;
}

int mua_verify(const uint8_t* d_sym, const uint8_t* d_dec, const int64_t* d_off, int64_t stride, int32_t C, int32_t S,
               const int32_t* d_start, const int32_t* d_end, unsigned long long* d_mismatch, void* stream) {
    REQUIRE(d_sym && d_dec && d_start && d_end && d_mismatch, "NULL argument");
    REQUIRE(S >= 2 && S <= MUA_MAX_S && C >= 0, "bad S/C");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(d_mismatch, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return cuda_fail(e, "memset");
    if (C == 0) return MUA_OK;
    const int grid = (C + 7) / 8 < sm_count() * 8 ? (C + 7) / 8 : sm_count() * 8;
    k_verify<<<grid, 256, 0, st>>>(d_sym, d_dec, d_off, stride, C, S, d_start, d_end, d_mismatch);
    CHECK_LAUNCH("k_verify");
    return MUA_OK;
}

int mua_online_histogram(uint8_t* d_x, int64_t n, int64_t H, int32_t max_firing_rate, uint32_t* d_counts, int32_t* d_first,
                         void* stream) {
    REQUIRE(d_x && d_counts && d_first, "NULL argument");
    REQUIRE(n >= 1, "empty input (the reference raises IndexError, functions_1.py:45)");
    REQUIRE(max_firing_rate >= 0 && max_firing_rate <= 255, "max_firing_rate outside 0..255");
    k_online_hist<<<1, 256, 0, (cudaStream_t)stream>>>(d_x, n, H, max_firing_rate, d_counts, d_first);
    CHECK_LAUNCH("k_online_hist");
    return MUA_OK;
}

int mua_approx_sort(const void* d_hist, int dtype, int32_t n, int64_t count, int64_t* d_idx, void* stream) {
    REQUIRE(d_hist && d_idx, "NULL argument");
    REQUIRE(n >= 1 && count >= 0, "bad n/count");
    if (count == 0) return MUA_OK;
    const int grid = (int)((count + 127) / 128);
    if (dtype == MUA_DT_I64) k_approx_sort<long long><<<grid, 128, 0, (cudaStream_t)stream>>>((const long long*)d_hist, n, count, d_idx);
    else if (dtype == MUA_DT_F64) k_approx_sort<double><<<grid, 128, 0, (cudaStream_t)stream>>>((const double*)d_hist, n, count, d_idx);
    else return fail(MUA_E_INVALID, "histogram dtype must be int64 or float64");
    CHECK_LAUNCH("k_approx_sort");
    return MUA_OK;
}

int mua_copy_rows(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width, int64_t rows, int32_t direction,
                  void* stream) {
    REQUIRE(dst && src, "NULL argument");
    REQUIRE(width >= 0 && rows >= 0 && dst_pitch >= width && src_pitch >= width, "bad pitch/width/rows");
    REQUIRE(direction == 0 || direction == 1, "direction must be 0 (H2D) or 1 (D2H)");
    if (width == 0 || rows == 0) return MUA_OK;
    cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src, (size_t)src_pitch, (size_t)width, (size_t)rows,
                                      direction == 0 ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpy2DAsync");
    return MUA_OK;
}

int mua_synth(uint8_t* d_sym, int64_t stride, int32_t T, int32_t C, int64_t c0, uint32_t seed, const uint32_t* d_thr, int32_t bursty,
              void* stream) {
    REQUIRE(d_sym && d_thr, "NULL argument");
    REQUIRE(aligned16(d_sym) && stride % 16 == 0 && stride >= (int64_t)((T + 15) & ~15), "stride must be a multiple of 16 and >= round_up(T,16)");
    REQUIRE(T >= 0 && C >= 0, "bad T/C");
    if (T == 0 || C == 0) return MUA_OK;
    const long long total = (long long)C * ((T + 15) / 16);
    const long long cap = (long long)sm_count() * 16;
    const long long need = (total + 255) / 256;
    k_synth<<<(int)(need < cap ? need : cap), 256, 0, (cudaStream_t)stream>>>(d_sym, stride, T, C, c0, seed, d_thr, bursty);
    CHECK_LAUNCH("k_synth");
    return MUA_OK;
}

}  // extern "C"
