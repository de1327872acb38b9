/*
 * mua_b200.h -- C ABI of the B200-native MUA compression hot path (libmua_b200.so).
 *
 * Drop-in boundary for the reference's `Compressing data/functions_1.py` path
 * (zhengzhang96/Hardware-efficient-MUA-compression).  The reference has no FFI of its own (it is
 * plain Python/NumPy); these entry points are what a ctypes binding for that path binds, one per
 * stage of the path.  Each declaration cites the reference lines it replaces (paths relative to the
 * reference root).
 *
 * Conventions
 *   - plain pointers and sizes only; every `d_*` pointer is DEVICE memory owned by the caller, every
 *     `h_*` pointer is HOST memory.  No hidden allocation, no retained state between calls.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work is
 *     enqueued asynchronously on it; no per-step call synchronises the device.  The one exception is the
 *     one-off setup call mua_build_tables, which waits for `stream` before it returns (its table header is
 *     uploaded from the caller's stack frame).
 *   - return value: 0 = ok, <0 = error (MUA_E_*); mua_last_error() returns a thread-local message.
 *   - re-entrant across distinct streams/devices.
 *   - kernel choice: mua_calibrate / mua_encode / mua_decode pick between two kernel families by the shape of the recording (a warp
 *     per channel; a lane per channel for recordings of many short rows) -- results are identical.  Three
 *     environment variables override the choice for A/B measurements and tests (MUA_ROWS_MIN_C, MUA_ROWS_T, MUA_ENC_NO_TENSOR);
 *     nothing else is read from the environment.
 *
 * Channel layout ("recording"): symbols are uint8, one row per channel:
 *     channel c occupies d_sym[off(c) .. off(c)+len(c)),  off(c) = d_off ? d_off[c] : c*stride,
 *     len(c) = d_len ? d_len[c] : T.
 *   d_sym and every off(c) must be 16-byte aligned and the buffer must be readable up to
 *   off(c) + round_up(len(c), 16) (rows padded to 16 B) -- the encoder stages rows with TMA bulk copies.
 *
 * Stream format (defined by oracle/mua_oracle.py:encode_channel; the Python reference emits no
 * bitstream): per channel the window [start,end) is saturated to S-1, mapped through the
 * approx-sort rank map of its calibration peak, coded with codebook row enc(c); codewords are
 * appended MSB-first, stream bit i lives in byte i/8 at bit 7-(i%8); zero-padded to 128 bits.
 * Side info: uint32 bit offset of every MUA_CHUNK-symbol chunk, chunks aligned to absolute bin index.
 */
#ifndef MUA_B200_H
#define MUA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MUA_ABI_VERSION 2
#define MUA_CHUNK 1024        /* symbols per decode chunk */
#define MUA_MAX_S 10          /* symbols per alphabet: 2..10 (get_BR_no_sort.py:104) */
#define MUA_MAX_K 35          /* candidate SCLVs for S=10 (Stored_SCLVs_S_10.pkl) */
#define MUA_MAX_H 16          /* history lengths per calibrate call (scripts use 9: 2^2..2^10) */
#define MUA_MAX_PEERS 16      /* GPUs of one node that can share a report sink */
#define MUA_IPC_HANDLE_BYTES 64

#define MUA_OK 0
#define MUA_E_INVALID (-1)    /* bad argument (message says which) */
#define MUA_E_CUDA (-2)       /* CUDA runtime error (message holds cudaGetErrorString) */

/* device-side flags: *d_overflow of mua_encode, *d_status of mua_decode (int32, zeroed by the caller, 0 = ok) */
#define MUA_ENC_OVERFLOW 1    /* a stream did not fit its slot (nothing was written past the slot) */
#define MUA_ENC_BAD_TABLE 2   /* table block does not match S/K/Lmax, or a channel's peak >= S / SCLV row >= K
                                 (that channel is not encoded: total_bits 0) */
#define MUA_DEC_BAD_OFFSET 1  /* a chunk's side-info bit offset lies past its slot (overflowed encode, corrupt side
                                 info): the chunk is skipped, nothing outside the stream buffer is read */
#define MUA_DEC_BAD_TABLE 2   /* table block does not match S/K/Lmax (nothing is decoded), or a channel's peak >= S /
                                 SCLV row >= K (that channel is skipped) */

/* window rule for the post-calibration ("to be compressed") window */
#define MUA_WINDOW_NONE 0     /* calibration only: no post window is scanned */
#define MUA_WINDOW_SKIP 1     /* end = cutoff + len/2; if end > len the channel is skipped
                                 (get_BR_no_sort.py:178-183) */
#define MUA_WINDOW_TRUNCATE 2 /* end = min(cutoff + len/2, len) (test_chosen_system.py:99-103) */

/* raster element types accepted by mua_bin_raster */
#define MUA_DT_U8 0
#define MUA_DT_I32 1
#define MUA_DT_I64 2
#define MUA_DT_F32 3
#define MUA_DT_F64 4

int mua_abi_version(void);
const char* mua_last_error(void);

/* ---- codebooks and lookup tables --------------------------------------------------------- */

/* Canonical Huffman codes for K ascending length rows [K][S] (host helper).  [1,2,2] -> 0,10,11,
 * the only codeword table in the Python hot path (test_chosen_system.py:26-27). */
int mua_canonical_codebook(const uint8_t* h_lens, int K, int S, uint16_t* h_codes_out);

/* Bytes of the device table block for alphabet size S with K codebook rows. */
size_t mua_tables_bytes(int S, int K);

/* Build the device table block: SCLV lengths (Stored_SCLVs_S_<S>.pkl rows, get_BR_no_sort.py:119-124),
 * codewords, and for every (peak p, row k): encode LUTs (saturate + approx_sort rank map
 * functions_1.py:75-90 + codeword) and the multi-symbol decode LUT. */
int mua_build_tables(void* d_tables, const uint8_t* h_lens, const uint16_t* h_codes, int S, int K,
                     void* stream);

/* ---- stage 1: binning --------------------------------------------------------------------- */

/* bin_MUA_data (functions_1.py:11-24): raster [T0][C] (row-major, element type `dtype`) summed over
 * `bin_res` consecutive rows -> ceil(T0/bin_res) bins, last bin partial.
 *   d_counts : int64 [nb][C] (reference layout/dtype) or NULL
 *   d_sym    : uint8 [C][sym_stride] channel-major symbols saturated at S-1 (saturation
 *              get_BR_no_sort.py:143,164), or NULL; S = 0 means clamp at 255 only. */
int mua_bin_raster(const void* d_raster, int dtype, int64_t T0, int32_t C, int32_t bin_res,
                   int64_t* d_counts, uint8_t* d_sym, int64_t sym_stride, int32_t S, void* stream);

/* MUA events (threshold-crossing times) -> binned count symbols: the MATLAB formatters' histogram2 over
 * `time_bins = min(t):BP/1000:max(t)` and one bin per channel, cast to uint8
 * (Data/Load_and_bin_Sabes_store_as_mat_file.m:49-54; same in the Flint and Brochier formatters).
 *   d_times : float64 [N] event times in seconds, any order;  d_chan : int32 [N] channel of every event (0-based)
 *   edge k  = t0 + k*w, computed in float64 as add(t0, mul(k, w)); bin k = [edge k, edge k+1), the last bin is
 *             closed on the right (histogram / np.histogram convention); events outside [edge 0, edge nb],
 *             events of channels outside [0, C) and NaN times are dropped
 *   d_sym   : uint8 [C][sym_stride] channel-major counts saturated at S-1 (S = 0: at 255, the uint8 cast);
 *             the call zeroes d_sym first (sym_stride >= nb, multiple of 4, d_sym 4-byte aligned) */
int mua_bin_events(const double* d_times, const int32_t* d_chan, int64_t N, double t0, double w, int64_t nb,
                   int32_t C, uint8_t* d_sym, int64_t sym_stride, int32_t S, void* stream);

/* ---- stages 2-4: calibration windows, histograms, approx-sort, SCLV selection ------------- */

/* For every channel c and history length H_h (h < nH):
 *   cutoff = min(max(H,1), len)                          functions_1.py:59-68 (the value callers use)
 *   assign = bincount(min(x,S-1)[:cutoff])               get_BR_no_sort.py:171
 *   peak   = first argmax(assign) (0 when use_sort == 0) functions_1.py:77
 *   end    per `window_mode`; post = bincount(min(x,S-1)[cutoff:end])   get_BR_no_sort.py:178-189
 *   *_m    = histogram permuted by approx_sort's idx     get_BR_with_approx_sort.py:175-176,193
 *   enc    = first argmin_k sum_r assign_m[r]*SCLV[k][r] over rows k with bit k of
 *            active_lo/active_hi set                     get_BR_no_sort.py:252,279
 *   bits   = sum_r SCLV[enc][r]*post_m[r]; nsym = sum post                  get_BR_no_sort.py:282-287
 * All outputs are optional (NULL skips), shaped [C][nH] (histograms [C][nH][S]).
 * d_end receives -1 for a skipped channel (MUA_WINDOW_SKIP). */
int mua_calibrate(const uint8_t* d_sym, const int64_t* d_off, const int32_t* d_len, int64_t stride,
                  int32_t T, int32_t C, int32_t S, const int32_t* h_H, int32_t nH, int32_t use_sort,
                  int32_t window_mode, const void* d_tables, uint32_t active_lo, uint32_t active_hi,
                  int32_t* d_cutoff, int32_t* d_end, uint8_t* d_peak, uint8_t* d_enc,
                  int32_t* d_assign_m, int32_t* d_post_m, int64_t* d_bits, int64_t* d_nsym,
                  void* stream);

/* Full-recording histogram, exactly sorted descending: np.flip(np.sort(hist))
 * (get_BR_no_sort.py:140-147).  d_hist_sorted int32 [C][S]. */
int mua_train_hist(const uint8_t* d_sym, const int64_t* d_off, const int32_t* d_len, int64_t stride,
                   int32_t T, int32_t C, int32_t S, int32_t* d_hist_sorted, void* stream);

/* One pass over the recording for SEVERAL alphabet sizes: the reference re-reads and re-saturates every channel for
 * each S of its sweep (`for S in range(2, 11)`, get_BR_no_sort.py:107 around :140-147 and :171-287); saturation at
 * S-1 never changes whether a count is >= v for v < S, so one scan with the thresholds of the largest S yields the
 * windowed histograms of all of them.  h_outs[i] names alphabet size i's table block and output buffers (same
 * meaning and shapes as the arguments of mua_calibrate / mua_train_hist; NULL outputs are skipped). */
typedef struct mua_calib_out {
    int32_t S;
    const void* d_tables;            /* table block of this S (unused by mua_train_hist_multi) */
    uint32_t active_lo, active_hi;
    int32_t* d_cutoff;
    int32_t* d_end;
    uint8_t* d_peak;
    uint8_t* d_enc;
    int32_t* d_assign_m;
    int32_t* d_post_m;
    int64_t* d_bits;
    int64_t* d_nsym;
    int32_t* d_train_hist;           /* mua_train_hist_multi only */
} mua_calib_out;
int mua_calibrate_multi(const uint8_t* d_sym, const int64_t* d_off, const int32_t* d_len, int64_t stride,
                        int32_t T, int32_t C, const int32_t* h_H, int32_t nH, int32_t use_sort,
                        int32_t window_mode, const mua_calib_out* h_outs, int32_t nS, void* stream);
int mua_train_hist_multi(const uint8_t* d_sym, const int64_t* d_off, const int32_t* d_len, int64_t stride,
                         int32_t T, int32_t C, const mua_calib_out* h_outs, int32_t nS, void* stream);

/* SCLV cost + selection on N histograms int32 [N][S] (np.matmul + np.argmin,
 * get_BR_no_sort.py:229-236): d_enc[n] = first argmin over active rows; optional d_min1/d_min2 =
 * smallest and second-smallest cost (for the elimination score, :307-316). */
int mua_select_sclv(const int32_t* d_hist, int64_t N, const void* d_tables, uint32_t active_lo,
                    uint32_t active_hi, uint8_t* d_enc, int64_t* d_min1, int64_t* d_min2, void* stream);

/* bits[n] = sum_r SCLV[enc[n]][r]*hist[n][r]; nsym[n] = sum_r hist[n][r] (get_BR_no_sort.py:282-287) */
int mua_bit_counts(const int32_t* d_hist, const uint8_t* d_enc, int64_t N, const void* d_tables,
                   int64_t* d_bits, int64_t* d_nsym, void* stream);

/* One greedy-elimination round over N train channels (get_BR_no_sort.py:237-240,307-316):
 *   d_assign_hist[k] = #channels whose argmin is row k      (int64 [K], zeroed by the call)
 *   d_score[j]       = sum_n (enc[n]==j ? min2[n] : min1[n]) (int64 [K], zeroed by the call) */
int mua_elim_scores(const uint8_t* d_enc, const int64_t* d_min1, const int64_t* d_min2, int64_t N,
                    int32_t K, int64_t* d_assign_hist, int64_t* d_score, void* stream);

/* ---- multi-GPU report sink ("only the per-channel bit counts and chosen-SCLV indices are gathered") ---------
 *
 * Channels shard over the GPUs of a node; what the BR report needs from every channel is 16 bytes: {bit count, window
 * symbols, SCLV row, peak} (get_BR_no_sort.py:282-291).  Instead of gathering them with a collective after the local
 * kernels, the encoder stores each channel's row straight into EVERY peer's report buffer through NVLink peer memory (one
 * 16-byte store per peer in the channel epilogue), so the transfer rides on the encode; a flag per source rank tells a
 * consumer when a step's rows have landed.
 *
 *   mua_peer_alloc   cudaMalloc + zero + cudaIpcGetMemHandle: the ONE allocating entry point of the library (IPC handles
 *                    need a whole allocation).  h_handle receives MUA_IPC_HANDLE_BYTES to hand to the other processes.
 *   mua_peer_open    map another process's buffer (cudaIpcOpenMemHandle, peer access enabled lazily); mua_peer_close unmaps.
 *   mua_peer_free    cudaFree of an own buffer.
 *   mua_report_signal  enqueue after mua_encode: rank `sink->rank` tells every peer "my rows of step `step` are written"
 *                    (system-scope release store of `step` into d_flags[peer][rank]).  Not needed when the sink passed to
 *                    mua_encode carries signal_step > 0: then the encoder's last block to retire does the same stores.
 *   mua_report_wait  enqueue where the report is needed: returns (on the stream) once d_flags[rank][p] >= step for every peer
 *                    p; gives up after ~2 s and stores 1 into d_flags[rank][MUA_MAX_PEERS] instead of hanging the device.
 * Steps are numbered 1, 2, ...; callers alternate two report buffers by step parity, so a peer that is one step ahead
 * never overwrites rows that are still being read (a rank cannot be two steps ahead: its wait needs every peer's signal). */
typedef struct mua_report_sink {
    int32_t n_peers;                       /* ranks sharing the report (<= MUA_MAX_PEERS); 0 = no sink */
    int32_t rank;                          /* this process's index */
    int64_t row0;                          /* global index of this rank's first channel */
    int32_t signal_step;                   /* mua_encode only: > 0 = the encoder itself signals this step when its last
                                              block retires (no mua_report_signal launch needed); 0 = it does not */
    int32_t reserved;
    int32_t* d_report[MUA_MAX_PEERS];      /* int32 [C_total][4] on every peer (entry `rank` = the own buffer) */
    int32_t* d_flags[MUA_MAX_PEERS];       /* int32 [MUA_MAX_PEERS + 2] on every peer: [p] = last step signalled by rank p,
                                              [MUA_MAX_PEERS] = wait timed out, [MUA_MAX_PEERS + 1] = block counter (zero it once) */
} mua_report_sink;
int mua_peer_alloc(size_t bytes, void** d_ptr, uint8_t* h_handle);
int mua_peer_open(const uint8_t* h_handle, void** d_ptr);
int mua_peer_close(void* d_ptr);
int mua_peer_free(void* d_ptr);
int mua_report_signal(const mua_report_sink* h_sink, int32_t step, void* stream);
int mua_report_wait(const mua_report_sink* h_sink, int32_t step, void* stream);

/* ---- stage 5: Huffman encode -------------------------------------------------------------- */

/* Encode window [d_start[c], d_end[c]) (d_end <= len; d_end <= d_start encodes nothing) of every
 * channel with rank map of d_peak[c] and codebook row d_enc[c].
 *   d_stream      : C slots of `slot_bytes` (multiple of 16); slot c holds the padded stream
 *   d_chunk_off   : uint32 [C][chunk_stride]; entry j = bit offset of chunk j (see header comment); this array and
 *                   d_sub_off must each be smaller than 2 GiB
 *   d_total_bits  : int64 [C]  (== SCLV[enc] . mapped post histogram, get_BR_no_sort.py:287)
 *   d_overflow    : int32 [1], MUA_ENC_OVERFLOW when a stream did not fit its slot, MUA_ENC_BAD_TABLE when
 *                   S/K/Lmax do not match the table block or a channel's peak/row is out of range
 *                   (caller zeroes it)
 *   K, Lmax       : rows and longest codeword of the table block (as passed to mua_build_tables);
 *                   host-side launch configuration only, the kernels cross-check them
 *   d_sub_off     : NULL, or uint32 [C][sub_stride], sub_stride >= 8 * chunk_stride: finer side info for the sub-chunk decoder.
 *                   Entry 8 j + i = bit offset of the first window symbol of the 128-symbol sub-chunk i of chunk j (absolute
 *                   bins [1024 (start/1024 + j) + 128 i, + 128); entry 8 j == d_chunk_off entry j); written for the
 *                   sub-chunks that intersect the window, and only by the encoder of codebooks with Lmax <= 2 and S <= 3 (the
 *                   chosen system); mua_decode uses it for exactly those codebooks and ignores it otherwise
 *   h_sink        : NULL, or the multi-GPU report sink: row (row0 + c) = {total_bits, max(end - start, 0), enc, peak}
 *                   as int32 is also stored into every peer's d_report (bit counts must fit int32: T * Lmax < 2^31) */
int mua_encode(const uint8_t* d_sym, const int64_t* d_off, const int32_t* d_len, int64_t stride,
               int32_t T, int32_t C, int32_t S, const int32_t* d_start, const int32_t* d_end,
               const uint8_t* d_peak, const uint8_t* d_enc, const void* d_tables, int32_t K,
               int32_t Lmax, uint8_t* d_stream, int64_t slot_bytes, uint32_t* d_chunk_off,
               int32_t chunk_stride, uint32_t* d_sub_off, int32_t sub_stride, int64_t* d_total_bits,
               int32_t* d_overflow, const mua_report_sink* h_sink, void* stream);

/* Pack the used part of every channel's slot back to back (what a caller ships off the device or stores: slots are sized for
 * the worst case, a stream uses ceil(total_bits / 128) 16-byte units of its slot).
 *   d_unit_off : int64 [C + 1] out: 16-byte unit offset of channel c in d_dense; [C] = total units
 *   d_dense    : dense_bytes bytes (units past the end are dropped; size it C * slot_bytes to be safe) */
int mua_pack_streams(const uint8_t* d_stream, int64_t slot_bytes, const int64_t* d_total_bits, int32_t C,
                     int64_t* d_unit_off, uint8_t* d_dense, int64_t dense_bytes, void* stream);

/* ---- stage 6: table-driven chunk-parallel decode ------------------------------------------ */

/* Inverse of mua_encode: symbols are written back at their absolute bin index, i.e. d_dec uses the
 * same layout (d_off/stride) as the input; bytes outside [start,end) are not touched.
 * max_end: an upper bound of every d_end[c] known to the host (e.g. H + T/2), or 0 if unknown; it
 * only sizes the launch (chunks past it are never scheduled).
 * d_status: int32 [1], zeroed by the caller; MUA_DEC_BAD_OFFSET / MUA_DEC_BAD_TABLE (the larger one wins) when
 * something was not decoded -- a decode after an encode that set MUA_ENC_OVERFLOW reports MUA_DEC_BAD_OFFSET
 * for the chunks that start past the slot and garbage-free output cannot be assumed for the truncated channel.
 * d_sub_off / sub_stride: NULL / 0, or the sub-chunk side info mua_encode wrote (same codebook class only): the decoder then
 * works on 128-symbol sub-chunks, 32 consecutive ones per warp, and writes 4 KB of consecutive symbols per warp and pass
 * instead of 32 rows of 128 bytes 1 KB apart (fixed row stride only, windows of at least eight chunks).
 * h_wait_sink / wait_step: NULL / 0, or the multi-GPU report sink: the decode's first block ends by polling the own flag
 * block until every rank has signalled `wait_step` (what mua_report_wait does as a separate launch), so a round-trip decode
 * that follows the encode also completes the gathered report. */
int mua_decode(const uint8_t* d_stream, int64_t slot_bytes, const uint32_t* d_chunk_off,
               int32_t chunk_stride, const uint32_t* d_sub_off, int32_t sub_stride, const int64_t* d_off, int64_t stride, int32_t C, int32_t S,
               const int32_t* d_start, const int32_t* d_end, const uint8_t* d_peak,
               const uint8_t* d_enc, const void* d_tables, int32_t K, int32_t Lmax, int32_t max_end,
               uint8_t* d_dec, int32_t* d_status, const mua_report_sink* h_wait_sink, int32_t wait_step,
               void* stream);

/* Round-trip check on the device: counts positions in [start,end) where d_dec != min(d_sym, S-1).
 * d_mismatch : uint64 [1] (zeroed by the call). */
int mua_verify(const uint8_t* d_sym, const uint8_t* d_dec, const int64_t* d_off, int64_t stride,
               int32_t C, int32_t S, const int32_t* d_start, const int32_t* d_end,
               unsigned long long* d_mismatch, void* stream);

/* ---- the literal functions_1.py signatures (batch-of-1; used by the drop-in module) --------- */

/* online_histogram_w_sat_based_nb_of_samples (functions_1.py:27-68) on one channel of n >= 1 uint8
 * samples: i = min(max(H,1), n); d_x[:i] is saturated IN PLACE (`>= max_firing_rate`, :45-46);
 * d_counts[v] (uint32 [256]) = occurrences of value v in d_x[:i]; d_first[v] (int32 [256]) = first
 * position of v (0x7FFFFFFF if absent) -- the dict's insertion order (:48-53). */
int mua_online_histogram(uint8_t* d_x, int64_t n, int64_t H, int32_t max_firing_rate,
                         uint32_t* d_counts, int32_t* d_first, void* stream);

/* approx_sort (functions_1.py:75-90) on `count` histograms of length n (dtype MUA_DT_I64 or
 * MUA_DT_F64): d_idx int64 [count][n] = rank -> symbol permutation around the first argmax. */
int mua_approx_sort(const void* d_hist, int dtype, int32_t n, int64_t count, int64_t* d_idx,
                    void* stream);

/* ---- host <-> device row copies (for callers that keep recordings in host memory) ----------- */

/* Copy `rows` rows of `width` bytes between pitched buffers (cudaMemcpy2DAsync on `stream`).
 * direction: 0 = host -> device, 1 = device -> host.  The path only reads bins [0, cutoff + len/2) of a
 * channel (test_chosen_system.py:99-103), so a caller uploads just that prefix of every row. */
int mua_copy_rows(void* dst, int64_t dst_pitch, const void* src, int64_t src_pitch, int64_t width,
                  int64_t rows, int32_t direction, void* stream);

/* ---- synthetic MUA (bench/test input; integer-only counter RNG, mirrored by the oracle) ---- */

/* d_sym[c][t] for c in [c0, c0+C): see oracle/mua_oracle.py:synth_symbols.  d_thr: uint32 [256][24]. */
int mua_synth(uint8_t* d_sym, int64_t stride, int32_t T, int32_t C, int64_t c0, uint32_t seed,
              const uint32_t* d_thr, int32_t bursty, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MUA_B200_H */
