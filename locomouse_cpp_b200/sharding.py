"""Frame-range sharding of the detection path across the GPUs of one box (one process per GPU).

The hot path is per-frame except for a one-frame look-back (checkVelCriterion reads the previous
image, LocoMouse_class.cpp:1256-1267, 1469-1470), so a video shards into contiguous frame ranges with
a ONE-FRAME TEMPORAL HALO and no data-path collective.  The only communication is the
final gather of the (compacted) candidate lists to rank 0, where the sequential host tracker
(match2nd) runs.  torch.distributed is plumbing: backend "nccl" on the GPU box, "gloo" in CPU tests.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .types import CAND_DTYPE, Results


def frame_range(n_frames: int, world: int, rank: int):
    """Contiguous, balanced [f0, f1) for `rank`; the first n % world ranks get one extra frame."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    q, r = divmod(int(n_frames), world)
    f0 = rank * q + min(rank, r)
    return f0, f0 + q + (1 if rank < r else 0)


def weighted_counts(total: int, weights, quantum: int = 1):
    """Frame counts per rank in proportion to `weights` (e.g. each rank's measured host-to-device rate: on a box whose GPUs sit
    behind unequal PCIe / memory paths the slowest rank bounds an evenly split job), summing to `total`.  Counts are multiples
    of `quantum` (the sub-batch size) except that the remainder goes to the heaviest rank; every rank gets at least one
    quantum when total allows."""
    w = np.asarray(list(weights), dtype=np.float64)
    if w.ndim != 1 or w.size == 0 or not np.all(np.isfinite(w)) or np.any(w <= 0):
        raise ValueError("weights must be positive and finite")
    total, quantum = int(total), max(1, int(quantum))
    if total < 0:
        raise ValueError("negative total")
    ideal = total * w / w.sum()
    counts = np.floor(ideal / quantum).astype(np.int64) * quantum
    if total >= quantum * w.size:
        counts = np.maximum(counts, quantum)
    # hand the remaining quanta to the ranks furthest below their ideal share, the last odd frames to the heaviest rank
    while counts.sum() + quantum <= total:
        counts[int(np.argmax(ideal - counts))] += quantum
    while counts.sum() > total:
        counts[int(np.argmax(counts - ideal))] -= min(quantum, int(counts.sum() - total))
    counts[int(np.argmax(w))] += total - int(counts.sum())
    return [int(c) for c in counts]


def weighted_ranges(total: int, weights, quantum: int = 1):
    """Contiguous [f0, f1) per rank with weighted_counts' sizes (rank r's range follows rank r - 1's)."""
    out, f0 = [], 0
    for c in weighted_counts(total, weights, quantum):
        out.append((f0, f0 + c))
        f0 += c
    return out


@dataclass(frozen=True)
class Shard:
    video: int
    f0: int
    f1: int

    @property
    def needs_halo(self) -> bool:
        """True when the frame before f0 must be supplied as prev_frame (video frame 0 has no look-back)."""
        return self.f0 > 0


def plan(n_videos: int, frames_per_video: int, world: int):
    """Per-rank list of Shards.  Whole videos are dealt round-robin when there are at least as many
    videos as ranks (no halo at all); otherwise every video is cut into `world` frame ranges."""
    out = [[] for _ in range(world)]
    if n_videos >= world:
        for v in range(n_videos):
            out[v % world].append(Shard(v, 0, frames_per_video))
    else:
        for v in range(n_videos):
            for r in range(world):
                f0, f1 = frame_range(frames_per_video, world, r)
                if f1 > f0:
                    out[r].append(Shard(v, f0, f1))
    return out


# ---- compact wire format for the gather ------------------------------------------------------------
def pack(res: Results) -> np.ndarray:
    """Results -> 1-D uint8 array holding only the live entries (counts + ragged lists)."""
    n, cap = res.n, res.cand_cap
    nb, ns = res.n_bottom[:n], res.n_side[:n]
    sel_b = np.arange(cap)[None, None, :] < nb[:, :, None]
    sel_s = np.arange(cap)[None, None, :] < ns[:, :, None]
    mn = res.match_n[:n][sel_b]
    per_list = np.where(sel_b, res.match_n[:n], 0).sum(axis=2).astype(np.int64)
    sel_m = np.arange(res.match_cap)[None, None, :] < per_list[:, :, None]
    head = np.array([n, cap, res.match_cap, res.n_tail_points], np.int64)
    parts = [head, nb, ns, res.bottom[:n][sel_b], res.side[:n][sel_s], mn.astype(np.int32),
             res.match_y[:n][sel_m], res.match_s[:n][sel_m], res.tail[:n], res.flags[:n]]
    return np.concatenate([np.ascontiguousarray(p).view(np.uint8).reshape(-1) for p in parts])


def unpack(buf: np.ndarray) -> Results:
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    o = 0

    def take(dtype, count):
        nonlocal o
        nbytes = np.dtype(dtype).itemsize * count
        a = buf[o:o + nbytes].view(dtype)
        o += nbytes
        return a

    n, cap, mcap, ntp = map(int, take(np.int64, 4))
    res = Results(n, cap, mcap, ntp)
    res.bottom[:] = np.array((-1, -1, -1.0), CAND_DTYPE)
    res.side[:] = np.array((-1, -1, -1.0), CAND_DTYPE)
    res.match_y[:] = -1
    res.match_s[:] = -1.0
    if n == 0:
        return res
    res.n_bottom[:n] = take(np.int32, n * 2).reshape(n, 2)
    res.n_side[:n] = take(np.int32, n * 2).reshape(n, 2)
    sel_b = np.arange(cap)[None, None, :] < res.n_bottom[:n][:, :, None]
    sel_s = np.arange(cap)[None, None, :] < res.n_side[:n][:, :, None]
    res.bottom[:n][sel_b] = take(CAND_DTYPE, int(sel_b.sum()))
    res.side[:n][sel_s] = take(CAND_DTYPE, int(sel_s.sum()))
    mn = take(np.int32, int(sel_b.sum()))
    res.match_n[:n][sel_b] = mn
    per_list = res.match_n[:n].sum(axis=2).astype(np.int64)
    sel_m = np.arange(mcap)[None, None, :] < per_list[:, :, None]
    res.match_y[:n][sel_m] = take(np.int32, int(sel_m.sum()))
    res.match_s[:n][sel_m] = take(np.float64, int(sel_m.sum()))
    res.tail[:n] = take(np.int32, n * 3 * ntp).reshape(n, 3, ntp)
    res.flags[:n] = take(np.uint32, n)
    return res


def concat(parts) -> Results:
    """Concatenate per-shard Results in frame order."""
    parts = list(parts)
    n = sum(p.n for p in parts)
    ref = parts[0]
    out = Results(n, ref.cand_cap, ref.match_cap, ref.n_tail_points)
    o = 0
    for p in parts:
        for name in Results.ARRAYS:
            getattr(out, name)[o:o + p.n] = getattr(p, name)[:p.n]
        o += p.n
    return out


_GATHER_CACHE = {}
_SHM_OWN = {}      # path -> (np.memmap, pinned pointer or None): segments this process created
_SHM_PEER = {}     # path -> np.memmap: other ranks' segments rank 0 has mapped


def _shm_path(tag: str, rank: int) -> str:
    import os
    import tempfile

    d = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    return os.path.join(d, f"lmres-{tag}-{rank}")


def shared_results(n: int, cand_cap: int, match_cap: int, n_tail_points: int, rank: int, tag: str, pin: bool = True) -> Results:
    """Result buffers for `n` frames in a named shared-memory file (page-locked when a CUDA device is present and `pin`):
    lm_detect_batch writes into them like into any page-locked arrays, and on a single box rank 0 reads every rank's records in
    place (gather_to_rank0 recognises such Results) -- the fixed-capacity records (11 kB per frame) then never cross PCIe a
    second and third time on their way to the host tracker.  `tag` must be the same on every rank of the job and unique per
    job (e.g. the rendezvous port)."""
    import atexit
    import os

    nbytes = Results.raw_nbytes(n, cand_cap, match_cap, n_tail_points)
    path = _shm_path(tag, rank)
    old = _SHM_OWN.pop(path, None)
    if old is not None:
        _release_own(path, old)
    if os.path.exists(path):   # left over from a run that died
        os.unlink(path)
    buf = np.memmap(path, dtype=np.uint8, mode="w+", shape=(nbytes,))
    buf[:] = 0
    pinned = None
    if pin:
        try:
            import torch

            if torch.cuda.is_available():
                if int(torch.cuda.cudart().cudaHostRegister(buf.ctypes.data, nbytes, 0)) == 0:
                    pinned = buf.ctypes.data
        except Exception:
            pinned = None
    if not _SHM_OWN:
        atexit.register(_release_all)
    _SHM_OWN[path] = (buf, pinned)
    res = Results(n, cand_cap, match_cap, n_tail_points, buffer=buf)
    res._shm_tag = tag
    res._shm_pinned = pinned is not None
    return res


def _release_own(path, entry):
    import os

    _buf, pinned = entry
    try:
        if pinned is not None:
            import torch

            torch.cuda.cudart().cudaHostUnregister(pinned)
    except Exception:
        pass
    try:
        os.unlink(path)
    except OSError:
        pass


def _release_all():
    _SHM_PEER.clear()
    for path, entry in list(_SHM_OWN.items()):
        _release_own(path, entry)
    _SHM_OWN.clear()


def gather_to_rank0(res: Results, device=None):
    """Gather every rank's Results on rank 0 (list in rank order; None elsewhere).

    Results created by shared_results (named shared memory, ranks of one box): rank 0 maps the peers' buffers, nothing moves.
    Otherwise, equal frame counts on all ranks: every rank's contiguous result buffer (Results.raw, fixed size) moves with ONE
    collective and rank 0 adopts the received buffers as Results views -- no per-entry host work.  Unequal counts on GPU ranks:
    the same buffers point to point.  Unequal counts on CPU ranks: compact payloads (pack / unpack), sizes exchanged first."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device(device) if device is not None else torch.device("cpu")
    meta = torch.tensor([res.n, res.cand_cap, res.match_cap, res.n_tail_points], dtype=torch.int64, device=dev)
    metas = [torch.zeros(4, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(metas, meta)
    metas = [tuple(int(v) for v in m.tolist()) for m in metas]
    tag = getattr(res, "_shm_tag", None)
    if tag is not None:
        # Results in named shared memory (shared_results): the all_gather above ordered every rank's writes before this point;
        # one more collective tells rank 0 whether all ranks use the shared path, then it maps the peers' segments.
        flag = torch.tensor([1], dtype=torch.int64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 1:
            if rank != 0:
                return None
            out = []
            for r, m in enumerate(metas):
                if r == 0:
                    out.append(res)
                    continue
                path = _shm_path(tag, r)
                nbytes = Results.raw_nbytes(m[0], m[1], m[2], m[3])
                # mapped anew at every call (microseconds): a rank may have re-created its file since the last one
                mm = _SHM_PEER[path] = np.memmap(path, dtype=np.uint8, mode="r+", shape=(nbytes,))
                out.append(Results(m[0], m[1], m[2], m[3], buffer=mm))
            return out
    if tag is None:
        # keep the collective sequence identical on ranks that do and do not use shared results (mixed use falls back below)
        flag = torch.tensor([0], dtype=torch.int64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if all(m == metas[0] for m in metas):
        n, cap, mcap, ntp = metas[0]
        if dev.type != "cuda":
            send = torch.from_numpy(res.raw)
            recv = [torch.empty_like(send) for _ in range(world)] if rank == 0 else None
            dist.gather(send, recv, dst=0)
            if rank != 0:
                return None
            return [Results(n, cap, mcap, ntp, buffer=r.numpy()) for r in recv]
        # GPU ranks: the receive buffers on the device and the page-locked host buffers they are copied into are allocated
        # once and reused (a fresh pageable .cpu() copy of world x 11 kB/frame per call had made rank 0 the bottleneck of
        # every step); the returned Results alias those host buffers and stay valid until the next call.
        nbytes = int(res.raw.nbytes)
        key = (world, rank, nbytes, str(dev))
        bufs = _GATHER_CACHE.get(key)
        if bufs is None:
            _GATHER_CACHE.clear()
            send = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            recv = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
            host = [torch.empty(nbytes, dtype=torch.uint8, pin_memory=True) for _ in range(world)] if rank == 0 else None
            bufs = _GATHER_CACHE[key] = (send, recv, host)
        send, recv, host = bufs
        send.copy_(torch.from_numpy(res.raw), non_blocking=True)
        dist.gather(send, recv, dst=0)
        if rank != 0:
            return None
        for h, r in zip(host, recv):
            h.copy_(r, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return [Results(n, cap, mcap, ntp, buffer=h.numpy()) for h in host]
    if dev.type == "cuda" and all(m[1:] == metas[0][1:] for m in metas):
        # unequal frame counts (weighted shards) with equal capacities: each rank's contiguous buffer moves as it is, point to
        # point, into reused device / page-locked buffers of the right sizes on rank 0 -- no per-entry host work either
        sizes = [Results.raw_nbytes(m[0], m[1], m[2], m[3]) for m in metas]
        key = ("ragged", world, rank, tuple(sizes), str(dev))
        bufs = _GATHER_CACHE.get(key)
        if bufs is None:
            _GATHER_CACHE.clear()
            send = torch.empty(sizes[rank], dtype=torch.uint8, device=dev)
            recv = [torch.empty(sz, dtype=torch.uint8, device=dev) for sz in sizes] if rank == 0 else None
            host = [torch.empty(sz, dtype=torch.uint8, pin_memory=True) for sz in sizes] if rank == 0 else None
            bufs = _GATHER_CACHE[key] = (send, recv, host)
        send, recv, host = bufs
        send.copy_(torch.from_numpy(res.raw), non_blocking=True)
        if rank == 0:
            ops = [dist.P2POp(dist.irecv, recv[r], r) for r in range(1, world)]
            recv[0].copy_(send, non_blocking=True)
        else:
            ops = [dist.P2POp(dist.isend, send, 0)]
        for w in (dist.batch_isend_irecv(ops) if ops else []):
            w.wait()
        if rank != 0:
            return None
        for h, r in zip(host, recv):
            h.copy_(r, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return [Results(m[0], m[1], m[2], m[3], buffer=h.numpy()) for m, h in zip(metas, host)]
    payload = torch.from_numpy(pack(res).copy())
    size = torch.tensor([payload.numel()], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, size)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    send = torch.zeros(mx, dtype=torch.uint8, device=dev)
    send[: payload.numel()] = payload.to(dev)
    recv = [torch.zeros(mx, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
    dist.gather(send, recv, dst=0)
    if rank != 0:
        return None
    return [unpack(r[:s].cpu().numpy()) for r, s in zip(recv, sizes)]


def bind_to_gpu_numa_node(local_rank: int) -> bool:
    """Pin this process to the CPUs that are local to its GPU (NVML's CPU affinity) so that pinned host buffers are
    first-touched on the GPU's own NUMA node: H2D copies of all ranks then scale instead of crossing the socket
    interconnect.  Returns False (and changes nothing) when NVML or the affinity call is unavailable."""
    import os

    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(local_rank))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False
