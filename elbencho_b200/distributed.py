"""Multi-process plumbing: one process per GPU (torch.distributed), no data-path collective.

The hot path shards by construction: worker rank r of numDataSetThreads owns a
contiguous global block range (LocalWorker.cpp:3576-3589), a disjoint random sub-range
(:3494-3495) or a private directory namespace (:3064-3068). A process therefore only needs its
rank offset (the reference's --rankoffset, ProgArgs.cpp:3845-3848) and the global thread count.

The only collective is the stats reduce that replaces the for-loops of
Statistics::getLiveOps / generatePhaseResults (Statistics.cpp:1340-1344, 1651-1737): sums of the
live counters and histogram buckets, min/max of the per-rank elapsed times and histogram extrema.
With the nccl backend the payload lives in device memory (NVLink/NVSwitch); with gloo (CPU tests)
the same code runs on host tensors.
"""
import dataclasses
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from ._native import LATHISTO_NUMBUCKETS

# layout of the summed vector
_SUM_FIELDS = [
    ("ops_total", "entries"), ("ops_total", "bytes"), ("ops_total", "iops"),
    ("ops_stonewall_total", "entries"), ("ops_stonewall_total", "bytes"),
    ("ops_stonewall_total", "iops"),
    ("ops_readmix_total", "entries"), ("ops_readmix_total", "bytes"), ("ops_readmix_total", "iops"),
]
_SUM_SCALARS = ["verify_mismatch_bytes", "verified_bytes", "filled_bytes", "num_kernel_launches",
                "h2d_bytes", "d2h_bytes", "dev_kernel_usec", "num_workers_done",
                "num_workers_done_with_error"]
_HISTOS = ["iops_lat_histo", "entries_lat_histo"]
_U64_MAX = (1 << 64) - 1
_I64_MAX = (1 << 63) - 1


def rank_layout(world_size: int, rank: int, threads_per_rank: int):
    """-> (rank_offset, num_dataset_threads) for this process (ProgArgs.cpp:3845-3848)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    return rank * threads_per_rank, world_size * threads_per_rank


def _to_i64(value: int) -> int:
    """u64 -> i64 bit pattern clamp for tensors (counters never reach 2^63 in practice; the
    histogram's 'empty' min of ~0 is mapped to i64 max)."""
    return min(int(value), _I64_MAX)


def pack_phase_results(res: Dict) -> Dict[str, List[int]]:
    sums = [res[a][b] for a, b in _SUM_FIELDS] + [res[k] for k in _SUM_SCALARS]
    for name in _HISTOS:
        sums += list(res[name]["buckets"]) + [res[name]["num"], res[name]["sum_usec"]]
    mins = [res["first_finish_usec"] or _I64_MAX] + [res[n]["min_usec"] for n in _HISTOS]
    maxs = [res["last_finish_usec"]] + [res[n]["max_usec"] for n in _HISTOS]
    return {"sum": [_to_i64(v) for v in sums], "min": [_to_i64(v) for v in mins],
            "max": [_to_i64(v) for v in maxs]}


def unpack_phase_results(packed: Dict[str, List[int]]) -> Dict:
    sums = list(packed["sum"])
    res: Dict = {}
    pos = 0
    for a, b in _SUM_FIELDS:
        res.setdefault(a, {})[b] = sums[pos]
        pos += 1
    for key in _SUM_SCALARS:
        res[key] = sums[pos]
        pos += 1
    for i, name in enumerate(_HISTOS):
        buckets = sums[pos:pos + LATHISTO_NUMBUCKETS]
        pos += LATHISTO_NUMBUCKETS
        res[name] = {"buckets": buckets, "num": sums[pos], "sum_usec": sums[pos + 1],
                     "min_usec": packed["min"][1 + i], "max_usec": packed["max"][1 + i]}
        pos += 2
    first = packed["min"][0]
    res["first_finish_usec"] = 0 if first == _I64_MAX else first
    res["last_finish_usec"] = packed["max"][0]
    return res


def per_sec_from_usec(total: int, elapsed_usec: int) -> int:
    """UnitTk::getPerSecFromUSec (toolkits/UnitTk.h:48-56)"""
    return int(total * (1000000.0 / elapsed_usec)) if elapsed_usec else 0


def finish_results(res: Dict) -> Dict:
    """per-second values of the reduced totals (Statistics.cpp:1726-1730)"""
    last = res["last_finish_usec"]
    first = res["first_finish_usec"]
    res["ops_per_sec"] = {k: per_sec_from_usec(v, last) for k, v in res["ops_total"].items()}
    res["ops_stonewall_per_sec"] = {k: per_sec_from_usec(v, first)
                                    for k, v in res["ops_stonewall_total"].items()}
    return res


def reduce_phase_results(res: Dict, device: Optional[torch.device] = None, group=None) -> Dict:
    """All ranks call this with their local phase results; every rank gets the job-wide result
    (sum / min / max reduce). Without an initialised process group it is the identity."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return finish_results(unpack_phase_results(pack_phase_results(res)))
    packed = pack_phase_results(res)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) \
            if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t_sum = torch.tensor(packed["sum"], dtype=torch.int64, device=device)
    t_min = torch.tensor(packed["min"], dtype=torch.int64, device=device)
    t_max = torch.tensor(packed["max"], dtype=torch.int64, device=device)
    dist.all_reduce(t_sum, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(t_min, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(t_max, op=dist.ReduceOp.MAX, group=group)
    return finish_results(unpack_phase_results({"sum": t_sum.tolist(), "min": t_min.tolist(),
                                                "max": t_max.tolist()}))


def reduce_device_counters(counters: torch.Tensor, group=None) -> torch.Tensor:
    """Sum the device-resident counter blocks of all ranks in place (the K2 mismatch counter lands
    in this block, so the reduce needs no host copy)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
    return counters


def reduce_max_float(value: float, device: Optional[torch.device] = None, group=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return value
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) \
            if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
