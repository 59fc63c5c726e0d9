"""Site sharding across the GPUs of one node.

One process per GPU (`torchrun`); every rank holds a contiguous range of sites.  Every
kernel is independent per site, so the only exchanges are tiny: the per-problem sum of
squared AF changes once per EM iteration (so that all ranks stop at the reference's global
iteration, emMAF.py:21-25), the allele-depth class tallies of the z-score, and the final
per-individual x per-population float64 partial sums.  Partials are all-gathered and
summed in rank order, so a result does not depend on reduction timing.

`torch.distributed` is the plumbing (NCCL over NVLink on GPUs, gloo in CPU tests).
"""
import numpy as np

_cfg = {"enabled": False, "rank": 0, "world": 1, "M_total": None, "offset": 0, "device": None, "ranges": None}


def _parse_cpulist(text):
    """'0-3,8,10-11' -> {0, 1, 2, 3, 8, 10, 11} (the format of sysfs cpulist files)."""
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_near_gpu(device, sysfs="/sys"):
    """Restrict this process to the CPUs of the NUMA node its GPU is attached to, BEFORE it allocates pinned host
    buffers: the pages of a pinned allocation land on the node of the allocating thread, and with one process per GPU
    started by torchrun about half of them would otherwise sit on the other socket, every upload crossing the
    socket interconnect (eight concurrent 4 GB uploads shared 163-180 GB/s in round 2's measurements against 47 GB/s for
    one alone).  Does nothing - and says why in the returned string - when the topology is not visible, the node has
    no CPU this process may use, or WGS_NO_NUMA_BIND is set.  Returns a one-line description of what it did."""
    import os
    if os.environ.get("WGS_NO_NUMA_BIND"):
        return "numa: off (WGS_NO_NUMA_BIND)"
    try:
        import ctypes
        from . import _lib
        buf = ctypes.create_string_buffer(32)
        if _lib.lib().wgs_device_pci_bus_id(int(device), buf, 32) != 0:
            return "numa: no PCI bus id for device %d" % device
        bus = buf.value.decode().lower()
        node = int(open(os.path.join(sysfs, "bus/pci/devices", bus, "numa_node")).read().strip())
        if node < 0:
            return "numa: %s reports no node" % bus
        cpus = _parse_cpulist(open(os.path.join(sysfs, "devices/system/node/node%d/cpulist" % node)).read())
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use:
            return "numa: node %d of %s has no CPU this process may use" % (node, bus)
        if use != allowed:
            os.sched_setaffinity(0, use)
        return "numa: device %d (%s) on node %d, %d of %d allowed CPUs" % (device, bus, node, len(use), len(allowed))
    except Exception as e:                                    # topology not visible (containers without sysfs PCI entries, ...)
        return "numa: not bound (%s: %s)" % (type(e).__name__, e)


def shard_range(M, rank, world):
    """Contiguous site range [lo, hi) of `rank` out of `world`."""
    lo = (M * rank) // world
    hi = (M * (rank + 1)) // world
    return lo, hi


def enable(M_total, offset, device=None, ranges=None):
    """Declare that this process holds sites [offset, offset+M_local) of M_total.  ranges: the [lo, hi) of every rank
    when they are not the equal split of shard_range (byte-range parts of a BGZF file hold unequal numbers of rows)."""
    import torch.distributed as td
    if not td.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    if ranges is not None:
        ranges = [(int(a), int(b)) for a, b in ranges]
        if len(ranges) != td.get_world_size() or ranges[0][0] != 0 or ranges[-1][1] != int(M_total) or \
                any(ranges[i][1] != ranges[i + 1][0] for i in range(len(ranges) - 1)) or ranges[td.get_rank()][0] != int(offset):
            raise ValueError("ranges must be contiguous, cover [0, M_total) and agree with offset")
    _cfg.update(enabled=True, rank=td.get_rank(), world=td.get_world_size(), M_total=int(M_total),
                offset=int(offset), device=device, ranges=ranges)


def disable():
    _cfg.update(enabled=False, rank=0, world=1, M_total=None, offset=0, device=None, ranges=None)


def rank_ranges():
    """[lo, hi) of every rank: the declared ranges, else the equal split."""
    if _cfg["ranges"] is not None:
        return list(_cfg["ranges"])
    return [shard_range(_cfg["M_total"], r, _cfg["world"]) for r in range(_cfg["world"])]


def row_counts_to_ranges(n_local):
    """All-gather every rank's row count (ranks hold consecutive pieces of one file): -> (M_total, offset, ranges)."""
    import torch
    import torch.distributed as td
    world, me = td.get_world_size(), td.get_rank()
    t = torch.zeros(world, dtype=torch.int64)
    t[me] = int(n_local)
    dev = _cfg["device"]
    if td.get_backend() == "nccl":
        t = t.cuda() if dev is None else t.to(dev)
    td.all_reduce(t)
    counts = [int(x) for x in t.cpu().tolist()]
    ranges, lo = [], 0
    for c in counts:
        ranges.append((lo, lo + c))
        lo += c
    return lo, ranges[me][0], ranges


def enabled():
    return _cfg["enabled"] and _cfg["world"] > 1


def rank():
    return _cfg["rank"]


def world():
    return _cfg["world"]


def total_sites(M_local):
    return _cfg["M_total"] if _cfg["enabled"] else M_local


def allreduce_sum(arr):
    """In-place sum of a float64/int64 NumPy array across ranks: all-gather, then a
    rank-ordered sum (deterministic for a given world size)."""
    if not enabled():
        return arr
    import torch
    import torch.distributed as td
    t = torch.from_numpy(np.ascontiguousarray(arr))
    dev = _cfg["device"]
    if dev is not None:
        t = t.to(dev)
    bufs = [torch.empty_like(t) for _ in range(_cfg["world"])]
    td.all_gather(bufs, t)
    acc = bufs[0].clone()
    for b in bufs[1:]:
        acc += b
    arr[...] = acc.cpu().numpy().reshape(arr.shape)
    return arr


def combine(ctx, arr):
    """Cross-rank sum of an operator's per-individual partial sums: nothing to do when the context has already
    combined them on the device (its own NCCL communicator), else the rank-ordered host sum."""
    if ctx is not None and ctx.partials_combined():
        return arr
    return allreduce_sum(arr)


def gather_rows(arr):
    """Concatenate per-rank row blocks (per-site outputs such as the AF matrix) on every rank."""
    if not enabled():
        return arr
    import torch
    import torch.distributed as td
    dev = _cfg["device"]
    counts = rank_ranges()
    out = []
    for r, (lo, hi) in enumerate(counts):
        shape = (hi - lo,) + tuple(arr.shape[1:])
        t = torch.from_numpy(np.ascontiguousarray(arr)) if r == _cfg["rank"] else torch.empty(shape, dtype=torch.from_numpy(arr[:0]).dtype)
        if dev is not None:
            t = t.to(dev)
        td.broadcast(t, src=r)
        out.append(t.cpu().numpy())
    return np.concatenate(out, axis=0)


def attach(ctx, nccl=True):
    """Give a Context the shard geometry, its rank, the all-reduce callback and - on GPUs - its own NCCL
    communicator (the id is created by rank 0 and broadcast through torch.distributed): with it the EM stop rule,
    the z-score table hand-over and every small cross-rank sum stay on the device (NVLink); without it
    (nccl=False, or no loadable libnccl) they pass through the host callback."""
    if enabled():
        ctx.set_shard(_cfg["M_total"], _cfg["offset"], allreduce_sum)
        ctx.set_rank(_cfg["rank"], _cfg["world"])
        import torch.distributed as td
        if nccl and _cfg["device"] is not None and td.get_backend() == "nccl" and not getattr(ctx, "_nccl_ready", False):
            import sys
            from . import _lib
            try:
                box = [_lib.nccl_unique_id() if _cfg["rank"] == 0 else None]
            except _lib.WgsError as e:                       # no loadable libnccl: the stop rule keeps the host callback
                box = [None]
                if _cfg["rank"] == 0:
                    print("wgsassign_b200: NCCL stop rule disabled (%s)" % e, file=sys.stderr)
            td.broadcast_object_list(box, src=0)
            if box[0] is not None:
                ctx.nccl_init(box[0], _cfg["rank"], _cfg["world"])
            ctx._nccl_ready = True
    else:
        ctx.set_shard(-1, 0, None)
