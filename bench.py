#!/usr/bin/env python
"""bench.py -- channel-bins/s encoded+decoded of the MUA compression hot path on N B200s.

One "step" = one pass of the hot path (calibrate -> Huffman encode -> chunk-parallel decode) over
this rank's shard of the cfg5 synthetic stream (BASELINE.json configs[4]: 1M channels x 1 h @ 50 ms
sharded over 8 GPUs = 125 000 channels x 72 000 bins per GPU, S=3, H=64, codebook 0/10/11);
weak scaling: every rank processes its own 125k-channel shard; the per-channel report (bits, symbols,
SCLV index, peak) reaches every rank inside the step through peer-memory stores of the encoder
(--report p2p, default) or one NCCL all-gather (--report nccl).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
  torchrun --nproc-per-node N bench.py --gpus N ...          (N > 1)

Prints ONE JSON line (rank 0).  `--impl reference` times the reference's CPU path instead: the
UNMODIFIED reference loop staged under oracle/_ref (oracle/make_ref.py), one process per host core."""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "channel-bins/s encoded+decoded"
UNIT = "channel-bins/s"
S, H, BP, SCLV = 3, 64, 50, (1, 2, 2)
SEED = 6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--channels", type=int, default=125000, help="channels per GPU")
    ap.add_argument("--bins", type=int, default=72000, help="bins per channel (1 h at 50 ms)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--e2e-blocks", type=int, default=10)
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="loop-time budget of the CPU baseline leg")
    ap.add_argument("--report", default="p2p", choices=["p2p", "nccl"], help="N > 1: how the per-channel report reaches every rank")
    ap.add_argument("--workload", default="cfg5", choices=["cfg5", "cfg4"],
                    help="cfg5: BASELINE configs[4], the headline (default); cfg4: one cell of the configs[3] sweep (--alphabet, --bp)")
    ap.add_argument("--alphabet", type=int, default=5, help="cfg4: alphabet size S (3, 5, 7, 9)")
    ap.add_argument("--bp", type=int, default=1, help="cfg4: bin period in ms (1, 10, 50); bins = 120 s / bp")
    a = ap.parse_args()
    if a.workload == "cfg4":
        a.bins = 120000 // a.bp
        if a.channels == 125000:
            a.channels = 100000
    return a


def workload(a):
    """the two bench workloads as data: alphabet, history lengths, window rule, SCLV rows, kernels of a step"""
    if a.workload == "cfg4":
        Sx = a.alphabet
        kern = {3: ["k_calibrate<3>", "k_encode_fast<3,tensor>", "k_decode_lane"]}.get(Sx, ["k_calibrate<%d>" % Sx, "k_encode_pair<%d>" % Sx if Sx < 10 else "k_encode_gen", "k_decode_var"])
        if a.bins <= 16384 and a.channels >= 8 * 32 * 148:
            # many short rows: the library picks its lane-per-channel kernels (mua_abi.cu: rows_t_max / rows_min_channels)
            kern[0] = "k_calibrate_rows<%d>" % Sx
            if Sx <= 3:
                kern[1] = "k_encode_rows<3,true>"
            elif Sx < 10:
                kern[1] = "k_encode_rows_pair<%d>" % Sx
            if Sx > 3 and a.bins <= 2048 + 1024:
                kern[2] = "k_decode_rows"
        return {"name": "cfg4", "S": Sx, "BP": a.bp, "T": a.bins, "H": [2 ** e for e in range(2, 11)], "h_enc": 4, "window": "skip",
                "sclv": None, "seed": 5, "kernels": kern,
                "want": ("cutoff", "end", "peak", "enc", "assign_m", "post_m", "bits", "nsym")}
    return {"name": "cfg5", "S": S, "BP": BP, "T": a.bins, "H": [H], "h_enc": 0, "window": "truncate", "sclv": np.array([SCLV]),
            "seed": SEED, "kernels": ["k_calibrate_head<3,4>", "k_encode_fast<3,tensor>", "k_decode_lane<1>"],
            "want": ("cutoff", "end", "peak", "enc")}


def workload_name(a):
    if a.workload == "cfg4":
        return ("cfg4 cell: %d channels x %d bins (120 s at %d ms), S=%d, all SCLV rows of Stored_SCLVs_S_%d, nine history lengths "
                "2^2..2^10 in one calibrate pass (skip rule), encode + decode at H=64, bursty Poisson" % (a.channels, a.bins, a.bp, a.alphabet, a.alphabet))
    return "cfg5 shard: %d channels x %d bins per GPU (1M-channel x 1-hour stream over 8 GPUs), S=3 H=64 BP=50ms, codebook 0/10/11, bursty Poisson" % (a.channels, a.bins)


def config_dict(a, world):
    """the same `config` keys in both arms (the driver compares them)"""
    C, T = a.channels, a.bins
    return {"workload": workload_name(a), "channels_per_gpu": C, "bins": T, "total_channels": C * world,
            "l2": "inputs (%.1f GB per GPU) are larger than L2" % (C * T / 1e9), "sharding": "channels, contiguous blocks",
            "synthetic": "integer counter RNG: 256 quantised Gamma(2,10) Hz rate classes, Poisson counts, independent 16-bin burst "
                         "blocks with p = 1/11 (rate class + 96) -- NOT the per-channel Gamma rate + 2-state Markov process of "
                         "SURVEY 8(d); chosen so that the CPU oracle can regenerate any channel of the device-generated stream"}


# ------------------------------------------------------------------------------------------------
# CPU baseline legs (the only place bench.py executes oracle/)
# ------------------------------------------------------------------------------------------------
# kind "reference": the UNMODIFIED reference loop (test_chosen_system.py:66-131 calling functions_1.py:27-68,75-90),
#   exec()ed from the staged copy oracle/_ref/ (oracle/make_ref.py; built by __graft_entry__.build()) under the
#   SURVEY Appendix-C harness (oracle/ref_harness.py); only the per-dataset loop is timed, the script's own imports,
#   path parsing and pickle.load run before the timer.
# kind "port": oracle/ref_port.py, a literal restatement with the same cost structure -- used only when no staged
#   reference is present (a checkout that never ran build() next to /root/reference).
_CPU_CACHE = {}


def _cpu_kind():
    from oracle import ref_harness
    return "reference" if ref_harness.reference_dir() else "port"


def _cpu_block(args):
    """worker: the reference loop on one block of channels; returns (seconds, post-window bins).
    The synthetic block (and the script workspace holding it) is made once per process, outside the timer."""
    seed, c0, nch, T, reps = args
    from oracle import mua_oracle as O, ref_port as R, ref_harness as RH
    key = (seed, c0, nch, T)
    if key not in _CPU_CACHE:
        thr = O.synth_threshold_table(float(BP))
        x = O.synth_symbols(seed, np.arange(c0, c0 + nch), T, thr, True)
        ws = None
        ref_dir = RH.reference_dir()
        if ref_dir:
            import atexit, shutil, tempfile
            ws = tempfile.mkdtemp(prefix="mua_refarm_")
            atexit.register(shutil.rmtree, ws, True)
            # all_binned_data[-2] is what the script reads (BP_counter = -2, test_chosen_system.py:23,55): one dataset
            chans = [np.ascontiguousarray(x[i]) for i in range(nch)]
            RH.write_workspace(ws, [[chans], [[]]], [BP, 100], os.path.join(ref_dir, "Produce SCLVs"), which=("test",))
        _CPU_CACHE[key] = (x, ws, ref_dir)
    x, ws, ref_dir = _CPU_CACHE[key]
    best, nsym = None, 0
    for _ in range(reps):
        if ws:
            dt, _, n = RH.chosen_system_timed(ref_dir, ws)       # the script re-loads (and then clips) its own copy
        else:
            ch = [x[i].copy() for i in range(nch)]
            t = time.perf_counter()
            _, nn = R.chosen_system_loop(ch, S=S, H=H, sclv=SCLV)
            dt = time.perf_counter() - t
            n = int(nn.sum())
        best = dt if best is None else min(best, dt)
        nsym = n
    return best, nsym


_WHAT = {"reference": "the UNMODIFIED reference loop test_chosen_system.py:66-131 + functions_1.py (staged copy oracle/_ref, "
                      "exec()ed under the Appendix-C harness; imports/path parsing/pickle.load outside the timer)",
         "port": "literal port of test_chosen_system.py:80-106 (oracle/ref_port.py; no staged reference found)"}


def cpu_baseline_single(T, nch=96, budget_s=10.0):
    """reference loop on ONE core over successive 96-channel x T blocks of the same synthetic stream until about
    `budget_s` seconds of loop time have been spent (synthetic generation is outside the timer)."""
    kind = _cpu_kind()
    loop_s, nsym, nblocks = 0.0, 0, 0
    t_wall = time.perf_counter()
    while loop_s < budget_s and time.perf_counter() - t_wall < 3 * budget_s + 20 and nblocks < 256:
        dt, n = _cpu_block((SEED, (nblocks % 4) * nch, nch, T, 1))      # 4 distinct blocks, cycled
        loop_s += dt
        nsym += n
        nblocks += 1
    return {"value": nsym / loop_s, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "%d passes over 96-channel blocks (4 distinct, cycled) of %d channels x %d bins of the workload (%.1f s of loop time), %s, "
                      "1 process (the reference is single-threaded); counts bits like the reference "
                      "(no bitstream, no decoder)" % (nblocks, nch, T, loop_s, _WHAT[kind]),
            "host_cores": os.cpu_count()}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    kind = _cpu_kind()
    cores = os.cpu_count() or 1
    nch = 96
    ctx = mp.get_context("fork")
    vals = []
    with ctx.Pool(cores) as pool:
        jobs = [(SEED, i * nch, nch, a.bins, 1) for i in range(cores)]
        for it in range(a.warmup + a.steps):
            t = time.perf_counter()
            res = pool.map(_cpu_block, jobs, chunksize=1)
            wall = time.perf_counter() - t
            # every core runs the loop on its own block in parallel; step time = slowest worker's loop time
            slow = max(r[0] for r in res)
            nsym = sum(r[1] for r in res)
            if it >= a.warmup:
                vals.append((nsym / slow, slow, wall))
    v = float(np.mean([x[0] for x in vals]))
    ms = float(np.mean([x[1] for x in vals]) * 1e3)
    one = cpu_baseline_single(a.bins, budget_s=min(a.cpu_seconds, 5.0))
    sample = ("%d processes (one per host core; NOT the reference's behaviour, which is one thread) x %d channels x %d bins per step, %s; "
              "synthetic generation outside the timer; the reference counts bits from histograms, it emits no bitstream "
              "and has no decoder; the same loop on ONE core (the reference as shipped): %.4g %s"
              % (cores, nch, a.bins, _WHAT[kind], one["value"], UNIT))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": config_dict(a, max(a.gpus, 1)),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "single_core_value": one["value"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons, pw = [], [], set(), []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # keep the samples taken under load (upper half of the power readings): the sampler also sees the
        # few idle milliseconds before the first and after the last step
        if pw:
            thr = 0.5 * (max(pw) + min(pw))
            load = [s_ for s_, p_ in zip(sm, pw) if p_ >= thr] or sm
        else:
            load = sm
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "samples_under_load": len(load),
                "window": "warm-up + timed steps (same load)", "reasons": sorted(reasons)}


def _timed(fn, n):
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def run_b200(a):
    import torch
    import torch.distributed as dist
    import mua_b200
    from mua_b200 import pipeline as P, dist as D, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = D.bind_to_gpu_numa_node(local)          # host buffers of this rank on the GPU's own NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = workload(a)
    C, T, S_, BP_ = a.channels, W["T"], W["S"], W["BP"]
    HS, h_enc = W["H"], W["h_enc"]
    H_ = HS[h_enc]
    C_total = C * world

    thr = P.synth_threshold_table(float(BP_))
    rec = P.synth_recording(C, T, seed=W["seed"], BP_ms=float(BP_), bursty=True, c0=rank * C, device=dev, thr=thr)
    cb = mua_b200.Codebook(S_, W["sclv"], device=dev)
    # cfg5 (the streaming system): the timed path only needs the calibration window (64 samples/channel): cutoff, window
    # end, peak, SCLV row; bit counts come out of the encoder itself.  cfg4 (the sweep): all nine history lengths with
    # their post-window histograms, selection and bit counts in one pass, then encode + decode at H = 64.
    want = W["want"]
    cal = P.calibrate(rec, cb, HS, use_sort=True, window=W["window"], want=want)
    sel = lambda k: cal[k][:, h_enc].contiguous() if len(HS) > 1 else cal[k][:, 0]
    max_end = H_ + T // 2
    slot = cb.worst_case_slot_bytes(T // 2 + 16)
    st0, en0, pk0, ec0 = sel("cutoff"), sel("end"), sel("peak"), sel("enc")
    es = P.encode(rec, cb, st0, en0, pk0, ec0, slot_bytes=slot)
    dec = torch.zeros_like(rec.sym)
    dec_status = torch.zeros(1, dtype=torch.int32, device=dev)
    # ---- the per-channel report (bit count, window symbols, SCLV row, peak) of all ranks on every rank ----
    # p2p : the encoder stores every channel's 16-byte row straight into all peers' report buffers over NVLink (dist.PeerReport:
    #       IPC-mapped peer memory, one flag per source rank); no collective, nothing left after the decode but a flag poll
    # nccl: one all_gather_into_tensor on a side stream after the encoder (round 1)
    report_mode, peer, peer_err = "none", None, None
    if world > 1:
        report_mode = a.report
        if report_mode == "p2p" and T * cb.Lmax >= 2 ** 31:
            report_mode = "nccl"
        if report_mode == "p2p":
            try:
                peer = D.PeerReport(C, C_total)
            except Exception as e:                       # no peer access between these GPUs: fall back, and say so
                peer_err = repr(e)[:200]
                peer = None
            ok = torch.tensor([1 if peer is not None else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                if peer is not None:
                    peer.close()
                    peer = None
                report_mode = "nccl"
    rep_dtype = torch.int32 if T * 16 < 2 ** 31 else torch.int64
    rep_buf = torch.empty((C_total, 4), dtype=rep_dtype, device=dev) if report_mode == "nccl" else None
    comm = torch.cuda.Stream(device=dev) if report_mode == "nccl" else None
    torch.cuda.synchronize()
    step_no = [0]
    launches_per_step = 3       # k_calibrate*, k_encode*, k_decode* (p2p: the encoder signals the report, the decoder waits for it)

    def step(ev=None):
        step_no[0] += 1
        k = step_no[0]
        P.calibrate(rec, cb, HS, use_sort=True, window=W["window"], want=want, out=cal)
        st, en, pk, ec = sel("cutoff"), sel("end"), sel("peak"), sel("enc")
        if ev: ev[1].record()
        P.encode(rec, cb, st, en, pk, ec, out=es, sink=peer.sink(k) if peer is not None else None)
        if ev: ev[2].record()
        rep = None
        if report_mode == "nccl":
            main = torch.cuda.current_stream()
            comm.wait_stream(main)
            with torch.cuda.stream(comm):
                rep = D.gather_channel_report(es.total_bits, en - st, ec, pk, C_total, out=rep_buf, dtype=rep_dtype)
        # p2p: the decoder's first block ends by polling the report flags (all set long before: the peers signalled at the end
        # of their encode), so the step needs no fourth launch
        P.decode(es, rec, cb, st, en, pk, ec, out=dec, max_end=max_end, status=dec_status,
                 wait_sink=peer.sink(k, signal=False) if peer is not None else None, wait_step=k if peer is not None else 0)
        if ev: ev[3].record()
        if report_mode == "nccl":
            main.wait_stream(comm)
        elif peer is not None:
            rep = k
        if ev: ev[4].record()
        return rep

    sampler = ClockSampler(local) if rank == 0 else None
    nw = 0
    for _ in range(max(a.warmup, 3)):                     # >= 3 warm-up steps (the first one also sets NCCL up) ...
        step()
        nw += 1
    torch.cuda.synchronize()
    # ... then ~1 s under load so that the clock sampler sees it.  Every step of an N > 1 run exchanges the report, so
    # the number of extra steps must be the same on all ranks: sized from three timed steps, MAX over ranks.
    t_w = time.perf_counter()
    for _ in range(3):
        step()
        nw += 1
    torch.cuda.synchronize()
    n_extra = int(min(1000, max(0, 1.0 / max((time.perf_counter() - t_w) / 3, 1e-5))))
    if world > 1:
        ne = torch.tensor([n_extra], dtype=torch.int64, device=dev)
        dist.all_reduce(ne, op=dist.ReduceOp.MAX)
        n_extra = int(ne.item())
    for i in range(n_extra):
        step()
        nw += 1
        if i % 8 == 7:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(a.steps)]
    t_end = torch.cuda.Event(enable_timing=True)
    rep = None
    for k in range(a.steps):
        evs[k][0].record()
        rep = step(evs[k])
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    total_ms = evs[0][0].elapsed_time(t_end)
    stage_all = np.array([[evs[k][i].elapsed_time(evs[k][i + 1]) for i in range(4)] for k in range(a.steps)])
    stage_ms = stage_all.mean(axis=0)
    if world > 1:
        tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        total_ms = float(tmax.item())

    # ---- checks outside the timed region: lossless, stream length == SCLV . histogram, the report ----
    st, en, pk, ec = sel("cutoff"), sel("end"), sel("peak"), sel("enc")
    mism = int(P.verify(rec, dec, S_, st, en).item())
    assert mism == 0, "decode is not lossless: %d mismatches" % mism
    assert int(es.overflow.item()) == 0 and int(dec_status.item()) == 0
    # the reference's bit count (histogram of the post window . SCLV, get_BR_no_sort.py:287) from a separate full scan
    ref = P.calibrate(rec, cb, [H_], use_sort=True, window=W["window"], want=("bits", "nsym"))
    assert torch.equal(es.total_bits, ref["bits"][:, 0]), "encoded length != SCLV . post histogram"
    assert torch.equal(ref["nsym"][:, 0], (en - st).clamp(min=0).to(torch.int64))
    nsym_local = int(ref["nsym"][:, 0].sum().item())
    bits_local = int(es.total_bits.sum().item())
    tot = torch.tensor([nsym_local, bits_local], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    nsym_all, bits_all = int(tot[0].item()), int(tot[1].item())
    local_rep = torch.stack([es.total_bits, (en - st).clamp(min=0).to(torch.int64), ec.to(torch.int64), pk.to(torch.int64)], dim=1)
    report_verified = None
    if world == 1:
        rep_t = local_rep
    else:
        if peer is not None:
            assert not peer.timed_out(), "a peer's report rows never arrived"
            rep_t = peer.report(rep).clone()
        else:
            rep_t = rep
        lo, hi = D.shard_range(C_total, rank, world)
        assert torch.equal(rep_t[lo:hi].to(torch.int64), local_rep), "gathered report != local slice"
        # identical on every rank: position-weighted checksums, MIN and MAX over ranks must agree
        w = torch.arange(1, C_total + 1, dtype=torch.int64, device=dev)
        r64 = rep_t.to(torch.int64)
        chk = torch.stack([(r64[:, j] * w).sum() for j in range(4)] + [r64.sum()])
        cmin, cmax = chk.clone(), chk.clone()
        dist.all_reduce(cmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(cmax, op=dist.ReduceOp.MAX)
        assert torch.equal(cmin, cmax), "the gathered report differs between ranks"
        assert int(r64[:, 0].sum().item()) == bits_all and int(r64[:, 1].sum().item()) == nsym_all
        report_verified = True
    br = D.br_report(rep_t, BP_) if rank == 0 else None

    ms_per_step = total_ms / a.steps
    value = nsym_all / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel, algorithmic bytes per launch ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak_gbs, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak_gbs, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    n_chunks = int(((en + 1023) // 1024 - st // 1024).clamp(min=0).sum().item())
    n_sub = int(((en + 127) // 128 - st // 128).clamp(min=0).sum().item()) if es.sub_off is not None and cb.S <= 3 and cb.Lmax <= 2 else 0
    sub_decode = n_sub > 0 and (max_end + 1023) // 1024 >= 8  # mua_decode's rule for the sub-chunk decoder (needs the 128-symbol offsets)
    side = 4 * n_chunks + 4 * n_sub + 8 * C                   # chunk offsets, 128-symbol sub-chunk offsets (fast encoder), bit counts
    enc_bytes = nsym_local + bits_local / 8 + side            # symbols read + stream written + side info
    dec_bytes = bits_local / 8 + nsym_local + (4 * n_sub if sub_decode else 4 * n_chunks)   # stream read + symbols written + offsets read
    cal_bytes = float(C) * min(T, (max(HS) + T // 2) if W["name"] == "cfg4" else H_)   # bins scanned by the calibrate pass
    stages = {"calibrate_ms": float(stage_ms[0]), "encode_ms": float(stage_ms[1]), "decode_ms": float(stage_ms[2]),
              "gather_ms": float(stage_ms[3]),   # N > 1: what is left of the report exchange after the decode it overlaps
              "calibrate_gbs": cal_bytes / stage_ms[0] / 1e6, "calibrate_frac": cal_bytes / stage_ms[0] / 1e6 / peak_gbs,
              "encode_gbs": enc_bytes / stage_ms[1] / 1e6, "decode_gbs": dec_bytes / stage_ms[2] / 1e6,
              "encode_frac": enc_bytes / stage_ms[1] / 1e6 / peak_gbs, "decode_frac": dec_bytes / stage_ms[2] / 1e6 / peak_gbs,
              "combined_gbs": (enc_bytes + dec_bytes) / (stage_ms[1] + stage_ms[2]) / 1e6,
              "combined_frac": (enc_bytes + dec_bytes) / (stage_ms[1] + stage_ms[2]) / 1e6 / peak_gbs,
              "bits_per_symbol": bits_local / max(nsym_local, 1),
              "per_step_ms": {"calibrate": [round(float(v), 4) for v in stage_all[:, 0]], "encode": [round(float(v), 4) for v in stage_all[:, 1]],
                              "decode": [round(float(v), 4) for v in stage_all[:, 2]], "gather": [round(float(v), 4) for v in stage_all[:, 3]]}}
    names = list(W["kernels"])
    if sub_decode:
        names[2] = "k_decode_sub"
    order = np.argsort([-stage_ms[0], -stage_ms[1], -stage_ms[2]])
    di = int(order[0])
    dom = names[di]
    ach = [stages["calibrate_gbs"], stages["encode_gbs"], stages["decode_gbs"]][di]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        key = ["k_calibrate", "k_encode", "k_decode"][di]
        if tj.get("channels") == C and tj.get("bins") == T and tj.get("workload", "cfg5") == W["name"]:
            traffic = tj.get(key)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak_gbs, "unit": "GB/s", "frac": ach / peak_gbs,
                "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": [cal_bytes, enc_bytes, dec_bytes][di]}

    # ---- e2e: host buffers in, compressed streams + report out ----
    e2e = None
    if not a.no_e2e:
        e2e = run_e2e(a, W, rec, cb, dev, world, rank)

    cpu = cpu_baseline_single(T, budget_s=a.cpu_seconds) if (rank == 0 and W["name"] == "cfg5") else None
    if rank == 0:
        cfg = config_dict(a, world)
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "warmup_steps_run": nw, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "config": cfg,
            "roofline": roofline, "stages": stages, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches_per_step * a.steps,
            "launches_per_step": {"own_kernels": launches_per_step, "kernels": names,
                                  "other": "none" if report_mode != "nccl" else "torch stack/cast kernels + NCCL all_gather on the side stream"},
            "report": {"mode": report_mode, "verified": report_verified, "fallback_reason": peer_err,
                       "what": {"p2p": "rows stored by the encoder into every peer's buffer over NVLink, its last block signals the step, the decoder's first block ends by polling the flags",
                                "nccl": "all_gather_into_tensor on a side stream", "none": "single GPU"}[report_mode]},
            "report_verified": report_verified,
            "clocks": clocks, "BR_bits_per_s_per_channel": float(br["BR"]), "lossless": True,
            "symbols_per_step": nsym_all, "numa_node_rank0": numa_node}))
    if peer is not None:
        peer.close()
    if world > 1:
        dist.destroy_process_group()


def run_e2e(a, W, rec, cb, dev, world, rank):
    """Same metric through the public API with HOST buffers.  The host owns this rank's WHOLE shard in pinned memory (the bins
    the path reads: [0, H + T//2) of every row).  Every step moves it to the device in channel blocks over `nstream` CUDA
    streams, runs calibrate / encode / decode on each block, packs the used part of the streams densely on the device
    (mua_pack_streams) and reads the dense streams, their offsets and the per-channel report back to pinned host memory.
    The dense size is only known on the device, so the read-back of block i is enqueued once the (tiny) offsets copy of that
    block has completed -- two blocks later in the loop, so the host never waits on the block it has just enqueued."""
    import torch
    import torch.distributed as dist
    from mua_b200 import pipeline as P
    C, T, S_ = a.channels, W["T"], W["S"]
    HS, h_enc = W["H"], W["h_enc"]
    H_ = HS[h_enc]
    nblk = max(1, min(a.e2e_blocks, C // 1000)) if C >= 1000 else 1
    nb = (C + nblk - 1) // nblk                             # channels per block (the last block may be shorter)
    nblk = (C + nb - 1) // nb
    nstream = 4
    slot = cb.worst_case_slot_bytes(T // 2 + 16)
    want = W["want"]
    max_end = H_ + T // 2
    need = min(rec.stride, (max_end + 15) // 16 * 16)      # bytes of every row the path reads
    sel = lambda cal, k: cal[k][:, h_enc].contiguous() if len(HS) > 1 else cal[k][:, 0]
    # ---- setup (untimed): the host's copy of the shard, per-stream device and host buffers ----
    h_in = torch.empty((C, need), dtype=torch.uint8, pin_memory=True)
    for i in range(nblk):
        h_in[i * nb:(i + 1) * nb].copy_(rec.sym[i * nb:(i + 1) * nb, :need])
    streams = [torch.cuda.Stream(device=dev) for _ in range(nstream)]
    bufs = []
    for s in streams:
        with torch.cuda.stream(s):
            r = P.Recording(sym=torch.zeros((nb, rec.stride), dtype=torch.uint8, device=dev), C=nb, T=T, stride=rec.stride)
            r.upload_rows(h_in[:nb], need)
            cal = P.calibrate(r, cb, HS, use_sort=True, window=W["window"], want=want)
            es = P.encode(r, cb, sel(cal, "cutoff"), sel(cal, "end"), sel(cal, "peak"), sel(cal, "enc"), slot_bytes=slot)
            bufs.append({"rec": r, "cal": cal, "es": es, "dec": torch.zeros_like(r.sym),
                         "dense": torch.empty(nb * slot, dtype=torch.uint8, device=dev),
                         "uoff": torch.empty(nb + 1, dtype=torch.int64, device=dev),
                         "rep": torch.empty((nb, 2), dtype=torch.int64, device=dev),
                         "status": torch.zeros(1, dtype=torch.int32, device=dev),
                         "h_dense": torch.empty(nb * slot, dtype=torch.uint8, pin_memory=True),
                         "h_uoff": torch.empty(nb + 1, dtype=torch.int64, pin_memory=True),
                         "h_rep": torch.empty((nb, 2), dtype=torch.int64, pin_memory=True),
                         "ev": torch.cuda.Event(), "done": torch.cuda.Event()})
    torch.cuda.synchronize()
    acct = {"d2h": 0, "nsym": 0, "bits": 0}

    def finish(i):
        """block i's offsets have been copied: enqueue the read-back of exactly its dense bytes"""
        b, s = bufs[i % nstream], streams[i % nstream]
        b["ev"].synchronize()
        nbytes = int(b["h_uoff"][-1 if (i + 1) * nb <= C else C - i * nb]) * 16
        with torch.cuda.stream(s):
            b["h_dense"][:nbytes].copy_(b["dense"][:nbytes], non_blocking=True)
            b["done"].record(s)
        acct["d2h"] += nbytes + b["h_uoff"].numel() * 8 + b["h_rep"].numel() * 8
        n_i = min(nb, C - i * nb)
        acct["nsym"] += int(b["h_rep"][:n_i, 1].sum())
        acct["bits"] += int(b["h_rep"][:n_i, 0].sum())

    def one_step():
        acct["d2h"] = acct["nsym"] = acct["bits"] = 0
        lag = nstream - 2
        for i in range(nblk):
            if i >= nstream:
                bufs[i % nstream]["done"].synchronize()          # the host buffers of this slot have been read back
            b, s = bufs[i % nstream], streams[i % nstream]
            n_i = min(nb, C - i * nb)
            with torch.cuda.stream(s):
                r, cal, es = b["rec"], b["cal"], b["es"]
                r.upload_rows(h_in[i * nb:i * nb + n_i], need)   # only the bins the path reads: [0, H + T//2)
                P.calibrate(r, cb, HS, use_sort=True, window=W["window"], want=want, out=cal)
                st, en, pk, ec = sel(cal, "cutoff"), sel(cal, "end"), sel(cal, "peak"), sel(cal, "enc")
                P.encode(r, cb, st, en, pk, ec, out=es)
                P.decode(es, r, cb, st, en, pk, ec, out=b["dec"], max_end=max_end, status=b["status"])
                P.pack_streams(es, b["dense"], b["uoff"])
                b["rep"][:, 0] = es.total_bits
                b["rep"][:, 1] = (en - st).clamp(min=0)
                b["h_uoff"].copy_(b["uoff"], non_blocking=True)
                b["h_rep"].copy_(b["rep"], non_blocking=True)
                b["ev"].record(s)
            if i >= lag:
                finish(i - lag)
        for i in range(max(nblk - lag, 0), nblk):
            finish(i)
        for s in streams:
            s.synchronize()

    one_step()
    torch.cuda.synchronize()
    assert all(int(b["status"].item()) == 0 and int(b["es"].overflow.item()) == 0 for b in bufs)
    # spot check of the read-back: block 0's first stream equals the slot it was packed from
    # ---- concurrent host->device probe: what the host side gives each rank when all ranks copy at once ----
    probe = None
    if world > 1 or True:
        tgt = torch.empty((nb, need), dtype=torch.uint8, device=dev)
        def h2d_all():
            for i in range(nblk):
                n_i = min(nb, C - i * nb)
                tgt[:n_i].copy_(h_in[i * nb:i * nb + n_i], non_blocking=True)
        h2d_all(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms_p = _timed(h2d_all, 1)
        mine = h_in.numel() / ms_p / 1e6
        if world > 1:
            allv = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
            dist.all_gather(allv, torch.tensor([mine], dtype=torch.float64, device=dev))
            per_rank = [float(v.item()) for v in allv]
        else:
            per_rank = [mine]
        probe = {"per_rank_gbs": [round(v, 2) for v in per_rank], "aggregate_gbs": round(sum(per_rank), 2),
                 "what": "every rank copies its whole pinned shard (%d x %d B) host->device at the same time, plain cudaMemcpyAsync, "
                         "nothing else running" % (C, need)}
        del tgt
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.e2e_steps):
        one_step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.e2e_steps
    wall_ms = (time.perf_counter() - t0) * 1e3 / a.e2e_steps
    ms_own = max(ms, wall_ms)
    ms = ms_own
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    tot = torch.tensor([acct["nsym"], acct["d2h"], int(h_in.numel())], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    nsym_all = int(tot[0].item())
    h2d_rate = h_in.numel() / ms_own / 1e6
    ceiling = None
    if probe:
        # if the host->device copies were the only thing in a step: symbols / (bytes / the slowest rank's concurrent copy rate)
        ceiling = nsym_all / (h_in.numel() / (min(probe["per_rank_gbs"]) * 1e9))
    return {"value": nsym_all / (ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(tot[2].item()),
            "d2h_bytes_per_step": int(tot[1].item()), "ms_per_step": ms, "steps": a.e2e_steps,
            "h2d_gbs_this_rank": round(h2d_rate, 2), "h2d_probe": probe, "h2d_only_ceiling": ceiling,
            "frac_of_h2d_ceiling": (nsym_all / (ms * 1e-3)) / ceiling if ceiling else None,
            "pinned_host_bytes_per_rank": int(h_in.numel()),
            "api": "mua_b200.pipeline.calibrate/encode/decode/pack_streams (C ABI) on pinned host buffers holding the rank's whole shard, "
                   "%d channel blocks of %d channels over %d CUDA streams; only bins [0, H + T//2) of every row are uploaded (all the "
                   "path reads); read back: the dense streams (ceil(bits/128)*16 B per channel), their offsets and the per-channel report"
                   % (nblk, nb, nstream)}


def _protect_stdout():
    """Everything libraries print to fd 1 (e.g. NCCL's version banner) goes to stderr; the JSON line is the only
    thing written to the real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


if __name__ == "__main__":
    _protect_stdout()
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
