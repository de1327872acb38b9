"""Channel-sharded multi-GPU plumbing (one process per GPU, torch.distributed).

Channels are independent units (per-channel state is cutoff, peak/rank map, chosen SCLV), so every
rank runs the whole pipeline on a contiguous block of channels with no data-path collective; only the
per-channel report (bit count, symbol count, chosen SCLV, peak) is gathered -- one all_gather of a
packed int64 [n, 4] tensor over NCCL/NVLink (gloo on CPU in the tests).  BR float math is done on the
host over the gathered array in global channel order (np.mean is order sensitive)."""
import os

import numpy as np
import torch
import torch.distributed as dist


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs: the PCI device's numa_node and
    the node's cpulist), so that the pinned host buffers it allocates afterwards are node-local (first touch) and
    the host->device copies of one rank do not cross the socket interconnect.  One process per GPU makes this
    safe; returns the node id, or None when the topology is not exposed (containers often report -1)."""
    try:
        prop = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus or cpus == allowed:
            return node if cpus else None
        os.sched_setaffinity(0, cpus)
        return node
    except (OSError, ValueError, AttributeError, RuntimeError, AssertionError):
        return None


def shard_range(C, rank, world):
    """Contiguous channel block [lo, hi) of `rank`; blocks differ by at most one channel."""
    return (C * rank) // world, (C * (rank + 1)) // world


def gather_channel_report(bits, nsym, enc, peak, C_total, group=None, out=None, dtype=torch.int64):
    """All ranks receive `dtype` [C_total, 4] = (bits, nsym, enc, peak) in global channel order.
    One collective per call; with equal shards (C_total divisible by the world size) the gathered
    buffer is returned as is -- pass `out` (`dtype` [C_total, 4]) to reuse it across calls.
    dtype=torch.int32 halves the bytes on the wire; the caller must know that a channel's bit count fits
    (it does whenever bins x longest codeword < 2^31)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    local = torch.stack([bits.to(dtype), nsym.to(dtype), enc.to(dtype), peak.to(dtype)], dim=1)
    if world == 1:
        assert local.shape[0] == C_total
        return local
    rank = dist.get_rank(group)
    lo, hi = shard_range(C_total, rank, world)
    assert local.shape[0] == hi - lo, "local shard does not match shard_range()"
    if C_total % world == 0:                                # equal shards: gather straight into the result
        if out is None:
            out = torch.empty((C_total, 4), dtype=dtype, device=local.device)
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    sizes = [shard_range(C_total, r, world)[1] - shard_range(C_total, r, world)[0] for r in range(world)]
    m = max(sizes)
    padded = torch.zeros((m, 4), dtype=dtype, device=local.device)
    padded[: local.shape[0]] = local
    buf = torch.empty((world * m, 4), dtype=dtype, device=local.device)
    dist.all_gather_into_tensor(buf, padded, group=group)
    return torch.cat([buf[r * m: r * m + sizes[r]] for r in range(world)], dim=0)


def br_report(report, BP):
    """Host-side BR report from a gathered report tensor: per-channel average bits/symbol and the
    chosen-system mean BR = np.mean(avg) / (BP/1000) (test_chosen_system.py:121-125)."""
    r = report.cpu().numpy()
    with np.errstate(all="ignore"):
        avg = r[:, 0].astype(np.float64) / r[:, 1].astype(np.float64)
        return {"avg_bits_per_symbol": avg, "BR": np.mean(avg) / (BP / 1000),
                "total_bits": int(r[:, 0].sum()), "total_symbols": int(r[:, 1].sum()),
                "enc": r[:, 2].astype(np.uint8), "peak": r[:, 3].astype(np.uint8)}


class _DevMem:
    """raw device allocation -> torch tensor (zero-copy) through __cuda_array_interface__"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerReport:
    """The multi-GPU report WITHOUT a collective (include/mua_b200.h "multi-GPU report sink"): every rank owns two report
    buffers int32 [C_total, 4] (alternated by step parity) and a flag block in one cudaMalloc'd, IPC-exported allocation;
    all ranks map each other's allocation (cudaIpcOpenMemHandle, NVLink peer access).  The encoder (pipeline.encode(sink=...))
    stores each channel's row {bits, symbols, SCLV row, peak} into all ranks' buffers in its channel epilogue;
    the encoder's last block (or a separate `signal(step)` launch) publishes "my rows are written", `wait(step)` -- enqueued where the report is needed,
    e.g. after the round-trip decode -- returns on the stream once every rank's rows of that step have landed.

    One process per GPU of one node; the handles travel through torch.distributed (all_gather_object)."""

    def __init__(self, C_local, C_total, group=None):
        import ctypes as C
        from . import _lib
        self._lib, self._C = _lib.load(), C
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        assert self.world <= _lib.MAX_PEERS
        lo, hi = shard_range(C_total, self.rank, self.world)
        assert hi - lo == C_local, "local shard does not match shard_range()"
        self.C_total, self.row0 = int(C_total), int(lo)
        self.rep_bytes = (C_total * 16 + 255) // 256 * 256
        self.nbytes = 2 * self.rep_bytes + 256
        self.dev = torch.device("cuda", torch.cuda.current_device())
        own = C.c_void_p()
        handle = (C.c_uint8 * _lib.IPC_HANDLE_BYTES)()
        _lib.check(self._lib.mua_peer_alloc(self.nbytes, C.byref(own), handle))
        self.own = own.value
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle), group=group)
        self.base = []
        for r in range(self.world):
            if r == self.rank:
                self.base.append(self.own)
                continue
            p = C.c_void_p()
            hb = (C.c_uint8 * _lib.IPC_HANDLE_BYTES).from_buffer_copy(handles[r])
            _lib.check(self._lib.mua_peer_open(hb, C.byref(p)))
            self.base.append(p.value)
        self._mem = _DevMem(self.own, self.nbytes)
        self._own_t = torch.as_tensor(self._mem, device=self.dev)
        self._sinks = []
        for parity in range(2):
            s = _lib.ReportSink()
            s.n_peers, s.rank, s.row0 = self.world, self.rank, self.row0
            for r in range(self.world):
                s.d_report[r] = self.base[r] + parity * self.rep_bytes
                s.d_flags[r] = self.base[r] + 2 * self.rep_bytes
            self._sinks.append(s)
        dist.barrier(group=group)                       # every rank has mapped every buffer before anyone stores into them

    def sink(self, step, signal=True):
        """the sink of step `step` (1, 2, ...): pass to pipeline.encode.  signal=True: the encoder itself publishes the step to
        the peers when its last block retires (no `signal` launch needed)."""
        s = self._sinks[step & 1]
        s.signal_step = int(step) if signal else 0
        return s

    def _stream(self):
        return self._C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def signal(self, step):
        from . import _lib
        _lib.check(self._lib.mua_report_signal(self._C.byref(self._sinks[step & 1]), int(step), self._stream()))

    def wait(self, step):
        from . import _lib
        _lib.check(self._lib.mua_report_wait(self._C.byref(self._sinks[step & 1]), int(step), self._stream()))

    def report(self, step):
        """int32 [C_total, 4] view of the own buffer of that step's parity (valid after wait(step) has completed)"""
        off = (step & 1) * self.rep_bytes
        return self._own_t[off: off + self.C_total * 16].view(torch.int32).view(self.C_total, 4)

    def timed_out(self):
        flags = self._own_t[2 * self.rep_bytes:].view(torch.int32)
        return bool(flags[16].item())

    def close(self):
        from . import _lib
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier()
        for r, p in enumerate(self.base):
            if r != self.rank and p:
                self._lib.mua_peer_close(self._C.c_void_p(p))
        self._own_t = None
        self._lib.mua_peer_free(self._C.c_void_p(self.own))
        self.base = []
