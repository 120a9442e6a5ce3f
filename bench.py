#!/usr/bin/env python
"""bench.py — headline benchmark of the GPU I/O worker (BASELINE.json configs[1]).

Workload (N=1): single file of --file-gib GiB (default 64), 1 MiB blocks, sequential write phase
then read phase with --verify, --gpuids <rank's GPU>, pinned-ring + cudaMemcpyAsync staging.

What is measured
  value   : GiB/s of the on-GPU work alone (K1 fill_pattern + K2 verify_pattern over a window of
            1 MiB blocks that is resident in HBM when the timed region starts). A "step" is one
            write pass (fill) + one read pass (verify) over the window = 2 x window bytes.
  e2e     : the same metric through the worker's public C ABI with host buffers and real files:
            write phase + read phase over the whole file, host<->device copies and storage I/O
            inside the timed region. (bytes written + bytes read) / (write time + read time).
  roofline: the dominant kernel's algorithmic bytes per launch / its CUDA-event duration, against
            the measured HBM peak of MEASURED_PEAKS.json.
  cpu_baseline: the CPU LocalWorker (oracle port of the reference loop) on a bounded sample of the
            same workload on this box's host cores (rank 0, N=1 only).

--impl reference times the reference's CPU implementation of the path (the oracle port; the
reference binary cannot be built here) on the host cores for the same metric.

Multi-GPU: one process per GPU under torchrun; rank r owns file r and worker ranks
[r*T, (r+1)*T) of N*T dataset threads (the reference's --rankoffset sharding); no data-path
collective; NCCL only reduces the stats counters. scaling = weak (per-GPU work fixed).
"""
import argparse
import ctypes
import json
import os
import shutil
import statistics
import sys
import threading
import time

REPO_ROOT = os.path.dirname(os.path.abspath(__file__))
if REPO_ROOT not in sys.path:
    sys.path.insert(0, REPO_ROOT)

MiB = 1 << 20
GiB = 1 << 30
METRIC = "seq_write_read_verify_throughput"
UNIT = "GiB/s"
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=100)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--file-gib", type=float, default=64.0, help="file size per GPU (GiB)")
    p.add_argument("--block-mib", type=float, default=1.0)
    p.add_argument("--threads", type=int, default=int(os.environ.get("ELB_BENCH_THREADS", "16")),
                   help="worker threads per GPU (-t)")
    p.add_argument("--window-gib", type=float, default=4.0,
                   help="HBM-resident window of the kernel-level measurement")
    p.add_argument("--dir", default=os.environ.get("ELB_BENCH_DIR", "/dev/shm"))
    p.add_argument("--salt", type=int, default=1)
    p.add_argument("--cpu-threads", type=int,
                   default=int(os.environ.get("ELB_BENCH_CPU_THREADS", "0")),
                   help="threads of the CPU LocalWorker arm (0 = calibrate: best of 1/4/8/16/32/nproc "
                        "on a small sample, i.e. all the host threads it can use productively)")
    p.add_argument("--cpu-sample-gib", type=float, default=8.0,
                   help="file size of the bounded CPU sample")
    p.add_argument("--ref-step-gib", type=float, default=2.0,
                   help="--impl reference: file size written+read per step")
    p.add_argument("--direct", action="store_true", help="O_DIRECT (--direct)")
    p.add_argument("--skip-e2e", action="store_true")
    p.add_argument("--skip-cpu", action="store_true")
    p.add_argument("--batch-blocks", type=int, default=0)
    p.add_argument("--num-batches", type=int, default=0)
    p.add_argument("--no-write-gate", action="store_true",
                   help="do not queue buffered writers of one file in user space")
    p.add_argument("--single-thread-sample-gib", type=float, default=4.0,
                   help="size of the extra -t 1 comparison (GPU worker vs CPU LocalWorker); 0 = skip")
    return p.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU while running (NVML, 10 ms period)."""

    REASONS = {
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
        0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks_setting",
    }

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None
            return self
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def _run(self):
        nv = self._nvml
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self._handle, nv.NVML_CLOCK_SM)
                util = nv.nvmlDeviceGetUtilizationRates(self._handle).gpu
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._handle)
                self.samples.append((mhz, util))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.01)

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        loaded = [m for m, u in self.samples if u > 0] or [m for m, _ in self.samples]
        return {"sm_mhz": statistics.median(loaded), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------

def load_hbm_peak():
    path = os.path.join(REPO_ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_ncu_traffic(window_bytes, kernel_key):
    """DRAM bytes (read + write) of one launch of the given kernel from the committed
    `ncu --set full` capture, if that capture used the same window size; else None."""
    path = os.path.join(REPO_ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            data = json.load(f)
        if int(data["window_bytes"]) != int(window_bytes):
            return None, None
        entry = data["kernels"][kernel_key]
        return entry["dram_bytes_read"] + entry["dram_bytes_write"], data.get("source")
    except Exception:
        return None, None


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return rank, local_rank, world


def bench_dir(args, rank):
    path = os.path.join(args.dir, "elb_bench_%d" % os.getuid())
    os.makedirs(path, exist_ok=True)
    return path


_CPU_THREADS_CACHE = {}


def cpu_threads_default(args, nfiles=1):
    """Thread count of the CPU LocalWorker arm. The reference's throughput on a single shared file
    peaks at a moderate thread count (buffered writes serialise on the inode lock) and falls
    beyond it, so 'all the host threads it can use' is found by a short calibration."""
    if args.cpu_threads:
        return args.cpu_threads
    if "best" in _CPU_THREADS_CACHE:
        return _CPU_THREADS_CACHE["best"]
    nproc = os.cpu_count() or 1
    candidates = sorted({t for t in (1, 4, 8, 16, 32, 64, nproc) if nfiles <= t <= nproc})
    block = int(args.block_mib * MiB)
    size = max(block * nproc, (4 * GiB) // nfiles)
    size -= size % block
    paths = [os.path.join(bench_dir(args, 0), "cpu_calibrate_%d.bin" % i) for i in range(nfiles)]
    best, best_val, table = candidates[0], 0.0, {}
    try:
        for threads in candidates:
            for path in paths:
                if os.path.exists(path):
                    os.unlink(path)
            res = run_cpu_localworker(paths, threads, size, block, args.salt, args.direct)
            table[threads] = round(res["gib_s"], 2)
            if res["gib_s"] > best_val:
                best, best_val = threads, res["gib_s"]
    finally:
        for path in paths:
            if os.path.exists(path):
                os.unlink(path)
    _CPU_THREADS_CACHE["best"] = best
    _CPU_THREADS_CACHE["table"] = table
    return best


# ------------------------------------------------------------------------------------------------
# the CPU LocalWorker arm (oracle port of the reference loop)
# ------------------------------------------------------------------------------------------------

def run_cpu_localworker(paths, threads, file_size, block_size, salt, direct, rank_offset=0,
                        dataset_threads=0):
    """write phase + read phase with --verify on the CPU. -> dict(bytes, usec, gib_s, iops)"""
    from elbencho_b200 import BenchPhase, WorkerConfig
    from tests import oracle_lib  # oracle: only used as the CPU baseline / reference arm here
    cfg = WorkerConfig(paths=paths, num_threads=threads, block_size=block_size,
                       file_size=file_size, integrity_check_salt=salt, use_direct_io=direct,
                       rank_offset=rank_offset, num_dataset_threads=dataset_threads)
    total_bytes = 0
    total_usec = 0
    total_iops = 0
    phases = {}
    for phase in (BenchPhase.CREATEFILES, BenchPhase.READFILES):
        rc, workers, pres = oracle_lib.run_oracle_phase(cfg, phase)
        if rc != 0:
            raise RuntimeError("CPU LocalWorker failed: " +
                               "; ".join(w.errorMsg.decode() for w in workers if w.hadError))
        total_bytes += pres.opsTotal.numBytesDone
        total_iops += pres.opsTotal.numIOPSDone
        total_usec += pres.lastFinishUSec
        phases[phase.name] = {"bytes": pres.opsTotal.numBytesDone, "usec": pres.lastFinishUSec}
    return {"bytes": total_bytes, "usec": total_usec, "iops": total_iops, "phases": phases,
            "gib_s": (total_bytes / GiB) / (total_usec / 1e6) if total_usec else 0.0}


def cpu_kernels_per_core(block_size, salt, total_bytes=256 * MiB):
    """GB/s of ONE host core for the reference's per-block operators (the oracle restatement of
    preWriteIntegrityCheckFillBuf / postReadIntegrityCheckVerifyBuf, LocalWorker.cpp:2091-2179) on
    a block-sized buffer, as a CPU counterpart of the K1 / K2 numbers."""
    import ctypes
    from tests import oracle_lib  # oracle: CPU baseline leg only
    lib = oracle_lib.load_oracle()
    buf = ctypes.create_string_buffer(block_size)
    nblocks = max(1, total_bytes // block_size)
    num, first = ctypes.c_uint64(), ctypes.c_uint64()
    exp, act = ctypes.c_uint(), ctypes.c_uint()
    msg = ctypes.create_string_buffer(256)
    out = {}
    t0 = time.perf_counter()
    for i in range(nblocks):
        lib.orc_fill_pattern(buf, block_size, i * block_size, salt)
    out["fill_pattern_gb_s"] = round(nblocks * block_size / (time.perf_counter() - t0) / 1e9, 2)
    t0 = time.perf_counter()
    bad = 0
    for i in range(nblocks):
        bad += lib.orc_verify_pattern(buf, block_size, (nblocks - 1) * block_size, salt,
                                      ctypes.byref(num), ctypes.byref(first), ctypes.byref(exp),
                                      ctypes.byref(act), msg, len(msg))
    out["verify_pattern_gb_s"] = round(nblocks * block_size / (time.perf_counter() - t0) / 1e9, 2)
    if bad:
        raise RuntimeError("CPU verify of a CPU-filled block failed")
    out["note"] = "one core, %d x %d KiB blocks, cache-resident buffer" % (nblocks,
                                                                            block_size // 1024)
    return out


def storage_roofline(args, threads, sample_size):
    """Raw pread/pwrite pass (no fill, no verify, no GPU) over a bounded sample on the same
    storage with the same thread count: the storage-bandwidth roofline of the e2e number."""
    block = int(args.block_mib * MiB)
    path = os.path.join(bench_dir(args, 0), "storage_roofline.bin")
    try:
        res = run_cpu_localworker([path], threads, sample_size, block, 0, args.direct)
    finally:
        if os.path.exists(path):
            os.unlink(path)
    return {"gib_s": round(res["gib_s"], 3),
            "write_gib_s": round(res["phases"]["CREATEFILES"]["bytes"] / GiB /
                                 (res["phases"]["CREATEFILES"]["usec"] / 1e6), 3),
            "read_gib_s": round(res["phases"]["READFILES"]["bytes"] / GiB /
                                (res["phases"]["READFILES"]["usec"] / 1e6), 3),
            "threads": threads, "sample_gib": sample_size / GiB}


def reference_arm(args):
    """bench.py --impl reference: the reference's CPU path on the host cores, same metric."""
    rank, local_rank, world = dist_env()
    if rank != 0:
        return 0  # other ranks exit without work
    # the GPU arm's config at N GPUs is N files (rank r <-> file r): same file set here
    nfiles = max(1, args.gpus)
    threads = cpu_threads_default(args, nfiles)
    block = int(args.block_mib * MiB)
    step_size = int(args.ref_step_gib * GiB)
    step_size -= step_size % block
    workdir = bench_dir(args, 0)
    paths = [os.path.join(workdir, "ref_arm_%d.bin" % i) for i in range(nfiles)]
    times = []
    nbytes = 0
    try:
        for step in range(args.warmup + args.steps):
            for path in paths:
                if os.path.exists(path):
                    os.unlink(path)
            res = run_cpu_localworker(paths, threads, step_size, block, args.salt, args.direct)
            if step >= args.warmup:
                times.append(res["usec"] / 1e6)
                nbytes += res["bytes"]
    finally:
        for path in paths:
            if os.path.exists(path):
                os.unlink(path)
    total = sum(times)
    value = (nbytes / GiB) / total if total else 0.0
    sample = "%d steps x (write+read --verify of %d x %.1f GiB file(s), %d MiB blocks, -t %d) in %s" % (
        args.steps, nfiles, step_size / GiB, block // MiB, threads, args.dir)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1000 * total / max(1, args.steps), 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "seq 1 MiB write+read --verify, CPU LocalWorker (oracle port of "
                               "LocalWorker.cpp:1669-1781, 2091-2179)",
                   "file_gib": step_size / GiB, "num_files": nfiles, "block_mib": args.block_mib,
                   "threads": threads,
                   "thread_calibration_gib_s": _CPU_THREADS_CACHE.get("table"),
                   "dir": args.dir, "direct": args.direct},
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample,
                         "operators_per_core": cpu_kernels_per_core(block, args.salt)},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# kernel-level measurement (HBM-resident window)
# ------------------------------------------------------------------------------------------------

def kernel_level(args, torch, device, rank):
    from elbencho_b200 import kernels
    block = int(args.block_mib * MiB)
    nblocks = max(1, int(args.window_gib * GiB) // block)
    window = nblocks * block
    file_size = int(args.file_gib * GiB)
    arena = torch.empty(window, dtype=torch.uint8, device=device)
    counters = torch.zeros(kernels.DEVCTR_NUM, dtype=torch.int64, device=device)
    results = torch.zeros(2 * nblocks, dtype=torch.int64, device=device)
    stream = torch.cuda.current_stream(device)
    handle = stream.cuda_stream
    total_steps = args.warmup + args.steps

    # per step: the window walks through the file (block i of step s at file offset ...)
    desc_tensors = []
    for step in range(total_steps):
        base = (step * window) % max(window, file_size - file_size % window)
        blocks = [(arena.data_ptr() + i * block, block, base + i * block, 0) for i in range(nblocks)]
        raw = kernels.pack_block_descs(blocks)
        desc_tensors.append(torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(device))
    rand_descs = desc_tensors[0]

    def ev():
        return torch.cuda.Event(enable_timing=True)

    fill_events = [(ev(), ev()) for _ in range(args.steps)]
    verify_events = [(ev(), ev()) for _ in range(args.steps)]

    def one_step(step, timed_idx=None):
        descs = desc_tensors[step]
        if timed_idx is not None:
            fill_events[timed_idx][0].record(stream)
        kernels.fill_pattern_batch(descs.data_ptr(), nblocks, args.salt, counters.data_ptr(),
                                   handle, total_bytes=window, max_block_len=block)
        if timed_idx is not None:
            fill_events[timed_idx][1].record(stream)
            verify_events[timed_idx][0].record(stream)
        kernels.verify_pattern_batch(descs.data_ptr(), nblocks, args.salt, results.data_ptr(),
                                     counters.data_ptr(), handle, total_bytes=window,
                                     max_block_len=block)
        if timed_idx is not None:
            verify_events[timed_idx][1].record(stream)

    for step in range(args.warmup):
        one_step(step)

    barrier(torch, device)
    launches_before = kernels.num_kernel_launches()
    start, end = ev(), ev()
    t0 = time.perf_counter()
    start.record(stream)
    for i in range(args.steps):
        one_step(args.warmup + i, i)
    end.record(stream)
    torch.cuda.synchronize(device)
    wall = time.perf_counter() - t0
    launches = kernels.num_kernel_launches() - launches_before
    elapsed_ms = start.elapsed_time(end)

    fill_ms = [a.elapsed_time(b) for a, b in fill_events]
    verify_ms = [a.elapsed_time(b) for a, b in verify_events]

    ctr = counters.cpu().tolist()
    mismatches = ctr[kernels.DEVCTR_VERIFY_MISMATCH_BYTES]
    if mismatches:
        raise RuntimeError("verify kernel reported %d mismatching bytes on freshly filled data"
                           % mismatches)
    if int(results.view(-1, 2)[:, 0].sum()) != 0:
        raise RuntimeError("per-block verify results are not clean")

    # K3 (random fill) outside the headline timed region, for the roofline table
    rnd_events = []
    for i in range(3 + 10):
        a, b = ev(), ev()
        a.record(stream)
        kernels.fill_random_batch(rand_descs.data_ptr(), nblocks, 100, 12345, 0, handle,
                                  total_bytes=window, max_block_len=block)
        b.record(stream)
        if i >= 3:
            rnd_events.append((a, b))
    torch.cuda.synchronize(device)
    rnd_ms = [a.elapsed_time(b) for a, b in rnd_events]

    return {
        "window_bytes": window, "nblocks": nblocks, "elapsed_ms": elapsed_ms, "wall_s": wall,
        "fill_ms_avg": statistics.mean(fill_ms), "verify_ms_avg": statistics.mean(verify_ms),
        "fill_ms_min": min(fill_ms), "verify_ms_min": min(verify_ms),
        "rand_ms_avg": statistics.mean(rnd_ms), "launches": launches,
    }


def barrier(torch, device):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier(device_ids=[device.index] if device.type == "cuda" else None)
    torch.cuda.synchronize(device)


# ------------------------------------------------------------------------------------------------
# end-to-end measurement through the worker ABI
# ------------------------------------------------------------------------------------------------

def fit_file_size_to_storage(args, torch, device, rank, world):
    """world files of --file-gib must fit into --dir (tmpfs pages are RAM): if they do not, all
    ranks agree on a smaller per-GPU file and the JSON line says so."""
    args.file_gib_requested = args.file_gib
    if args.skip_e2e:
        return
    os.makedirs(args.dir, exist_ok=True)
    stat = os.statvfs(args.dir)
    free_gib = stat.f_bavail * stat.f_frsize / GiB
    fit = torch.tensor([free_gib], dtype=torch.float64, device=device)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(fit, op=dist.ReduceOp.MIN)
    usable_gib = float(fit.item()) * 0.8 / world  # leave room for warm-up and baseline files
    if args.file_gib > usable_gib:
        args.file_gib = max(1.0, float(int(usable_gib)))


def e2e_level(args, torch, device, rank, world):
    from elbencho_b200 import BenchPhase, WorkerConfig, WorkerManager
    from elbencho_b200 import distributed as elbdist
    block = int(args.block_mib * MiB)
    file_size = int(args.file_gib * GiB)
    file_size -= file_size % block
    workdir = bench_dir(args, rank)
    paths = [os.path.join(workdir, "bench_file_%d.bin" % r) for r in range(world)]
    rank_offset, dataset_threads = elbdist.rank_layout(world, rank, args.threads)

    # warm-up pass on a small private file: CUDA context, pinned rings, kernels, page cache code
    warm_path = os.path.join(workdir, "warm_%d.bin" % rank)
    with WorkerManager(WorkerConfig(paths=[warm_path], num_threads=args.threads,
                                    block_size=block, file_size=max(block * args.threads * 32,
                                                                    256 * MiB),
                                    integrity_check_salt=args.salt, gpu_ids=[device.index],
                                    use_direct_io=args.direct,
                                    pipeline_batch_blocks=args.batch_blocks,
                                    pipeline_num_batches=args.num_batches)) as mgr:
        for _ in range(3):
            mgr.run_phase(BenchPhase.CREATEFILES)
            mgr.run_phase(BenchPhase.READFILES)
        mgr.run_phase(BenchPhase.DELETEFILES)

    for path in paths[rank:rank + 1]:
        if os.path.exists(path):
            os.unlink(path)
    barrier(torch, device)

    cfg = WorkerConfig(paths=paths, num_threads=args.threads, rank_offset=rank_offset,
                       num_dataset_threads=dataset_threads, block_size=block,
                       file_size=file_size, integrity_check_salt=args.salt,
                       gpu_ids=[device.index], use_direct_io=args.direct,
                       pipeline_batch_blocks=args.batch_blocks,
                       pipeline_num_batches=args.num_batches,
                       serialize_buffered_writes=not args.no_write_gate)
    out = {}
    with WorkerManager(cfg) as mgr:
        total_usec = 0
        total_bytes = 0
        for phase in (BenchPhase.CREATEFILES, BenchPhase.READFILES):
            barrier(torch, device)
            local = mgr.run_phase(phase)
            barrier(torch, device)
            res = elbdist.reduce_phase_results(local, device)  # NCCL: stats only
            out[phase.name] = res
            total_usec += res["last_finish_usec"]  # max over all ranks' workers
            total_bytes += res["ops_total"]["bytes"]
        if out["READFILES"]["verify_mismatch_bytes"]:
            raise RuntimeError("e2e read phase found integrity mismatches")
        expected = file_size * world
        for name in ("CREATEFILES", "READFILES"):
            if out[name]["ops_total"]["bytes"] != expected:
                raise RuntimeError("e2e %s moved %d bytes, expected %d" % (
                    name, out[name]["ops_total"]["bytes"], expected))
    barrier(torch, device)
    if os.path.exists(paths[rank]):
        os.unlink(paths[rank])

    w, r = out["CREATEFILES"], out["READFILES"]
    return {
        "gib_s": (total_bytes / GiB) / (total_usec / 1e6),
        "write_gib_s": (w["ops_total"]["bytes"] / GiB) / (w["last_finish_usec"] / 1e6),
        "read_gib_s": (r["ops_total"]["bytes"] / GiB) / (r["last_finish_usec"] / 1e6),
        "write_iops": w["ops_per_sec"]["iops"], "read_iops": r["ops_per_sec"]["iops"],
        "write_first_done_gib_s": (w["ops_stonewall_total"]["bytes"] / GiB) /
                                  (w["first_finish_usec"] / 1e6),
        "read_first_done_gib_s": (r["ops_stonewall_total"]["bytes"] / GiB) /
                                 (r["first_finish_usec"] / 1e6),
        "h2d_bytes": r["h2d_bytes"] + w["h2d_bytes"], "d2h_bytes": w["d2h_bytes"] + r["d2h_bytes"],
        "launches": w["num_kernel_launches"] + r["num_kernel_launches"],
        "dev_kernel_usec": w["dev_kernel_usec"] + r["dev_kernel_usec"],
        "file_bytes": file_size, "total_usec": total_usec,
        "lat_write_avg_usec": w["iops_lat_histo"]["sum_usec"] / max(1, w["iops_lat_histo"]["num"]),
        "lat_read_avg_usec": r["iops_lat_histo"]["sum_usec"] / max(1, r["iops_lat_histo"]["num"]),
    }


def single_thread_compare(args, device):
    """The literal '-t 1' form of the config: one worker thread on either arm,
    bounded sample. With one thread the reference serialises fill -> write and read -> verify on
    the CPU, while the GPU worker overlaps its on-GPU work with the storage call."""
    from elbencho_b200 import BenchPhase, WorkerConfig, WorkerManager
    block = int(args.block_mib * MiB)
    size = int(args.single_thread_sample_gib * GiB)
    size -= size % block
    workdir = bench_dir(args, 0)
    gpu_path = os.path.join(workdir, "single_gpu.bin")
    cpu_path = os.path.join(workdir, "single_cpu.bin")
    out = {"sample": "write+read --verify of a %.1f GiB file, 1 MiB blocks, -t 1" % (size / GiB)}
    try:
        total_usec = 0
        with WorkerManager(WorkerConfig(paths=[gpu_path], num_threads=1, block_size=block,
                                        file_size=size, integrity_check_salt=args.salt,
                                        gpu_ids=[device.index], use_direct_io=args.direct)) as mgr:
            for phase in (BenchPhase.CREATEFILES, BenchPhase.READFILES):
                total_usec += mgr.run_phase(phase)["last_finish_usec"]
        out["gpu_worker_gib_s"] = round((2 * size / GiB) / (total_usec / 1e6), 3)
        res = run_cpu_localworker([cpu_path], 1, size, block, args.salt, args.direct)
        out["cpu_localworker_gib_s"] = round(res["gib_s"], 3)
        out["ratio"] = round(out["gpu_worker_gib_s"] / out["cpu_localworker_gib_s"], 2)
    finally:
        for path in (gpu_path, cpu_path):
            if os.path.exists(path):
                os.unlink(path)
    return out


def same_sample_compare(args, device, cpu):
    """The GPU worker on exactly the bounded sample the cpu_baseline was timed on (same file size,
    same thread count, fresh file): an apples-to-apples pair, because on this tmpfs the write rate
    depends on the file size. Never fails the bench: errors are reported in the result."""
    out = {}
    try:
        from elbencho_b200 import BenchPhase, WorkerConfig, WorkerManager
        block = int(args.block_mib * MiB)
        size = int(args.cpu_sample_gib * GiB)
        size -= size % block
        threads = int(cpu["cores"])
        path = os.path.join(bench_dir(args, 0), "same_sample_gpu.bin")
        out["sample"] = "write+read --verify of a %.1f GiB file, 1 MiB blocks, -t %d" % (
            size / GiB, threads)
        try:
            total_usec = 0
            with WorkerManager(WorkerConfig(paths=[path], num_threads=threads, block_size=block,
                                            file_size=size, integrity_check_salt=args.salt,
                                            gpu_ids=[device.index], use_direct_io=args.direct,
                                            serialize_buffered_writes=not args.no_write_gate)) as mgr:
                for phase in (BenchPhase.CREATEFILES, BenchPhase.READFILES):
                    total_usec += mgr.run_phase(phase)["last_finish_usec"]
            out["gpu_worker_gib_s"] = round((2 * size / GiB) / (total_usec / 1e6), 3)
            out["cpu_localworker_gib_s"] = cpu["value"]
            out["ratio"] = round(out["gpu_worker_gib_s"] / cpu["value"], 2) if cpu["value"] else None
        finally:
            if os.path.exists(path):
                os.unlink(path)
    except Exception as err:  # noqa: BLE001 (an extra, must not break the bench line)
        out["error"] = str(err)
    return out


# ------------------------------------------------------------------------------------------------
# main
# ------------------------------------------------------------------------------------------------

_REAL_STDOUT = None


def protect_stdout():
    """The driver parses ONE JSON line from stdout; libraries (NCCL's version banner) also write
    there. Keep the real stdout aside and point fd 1 at stderr for everything else."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse_args()
    protect_stdout()

    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from elbencho_b200 import _native

    _native.load()  # fails loudly if the CUDA library is missing

    rank, local_rank, world = dist_env()
    if world != args.gpus and world > 1:
        raise SystemExit("WORLD_SIZE=%d but --gpus %d" % (world, args.gpus))
    if args.gpus > 1 and world == 1:
        raise SystemExit("--gpus %d needs one process per GPU: launch with "
                         "python -m torch.distributed.run --nproc-per-node %d ..." % (
                             args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA GPU (there is no CPU fallback for the product path)")

    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)

    fit_file_size_to_storage(args, torch, device, rank, world)

    sampler = ClockSampler(local_rank).start() if rank == 0 else None

    kern = kernel_level(args, torch, device, rank)

    # whole-job kernel-level throughput: all ranks' bytes / max-over-ranks device time
    from elbencho_b200 import distributed as elbdist
    elapsed_ms = elbdist.reduce_max_float(kern["elapsed_ms"], device)
    job_bytes = 2 * kern["window_bytes"] * args.steps * world
    value = (job_bytes / GiB) / (elapsed_ms / 1e3)

    e2e = None
    if not args.skip_e2e:
        e2e = e2e_level(args, torch, device, rank, world)

    clocks = sampler.stop() if sampler else None

    cpu = None
    storage = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        threads = cpu_threads_default(args)
        block = int(args.block_mib * MiB)
        sample_size = int(args.cpu_sample_gib * GiB)
        storage = storage_roofline(args, args.threads, sample_size)
        path = os.path.join(bench_dir(args, 0), "cpu_baseline.bin")
        try:
            res = run_cpu_localworker([path], threads, sample_size, block, args.salt, args.direct)
        finally:
            if os.path.exists(path):
                os.unlink(path)
        cpu = {"value": round(res["gib_s"], 3), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "write+read --verify of a %.1f GiB file, 1 MiB blocks, -t %d, in %s "
                         "(oracle port of LocalWorker.cpp:1669-1781 + 2091-2179)" % (
                             sample_size / GiB, threads, args.dir),
               "thread_calibration_gib_s": _CPU_THREADS_CACHE.get("table"),
               "operators_per_core": cpu_kernels_per_core(block, args.salt),
               "write_gib_s": round(res["phases"]["CREATEFILES"]["bytes"] / GiB /
                                    (res["phases"]["CREATEFILES"]["usec"] / 1e6), 3),
               "read_gib_s": round(res["phases"]["READFILES"]["bytes"] / GiB /
                                   (res["phases"]["READFILES"]["usec"] / 1e6), 3)}

    single = None
    same_sample = None
    if rank == 0 and world == 1 and not args.skip_cpu and not args.skip_e2e and \
            args.single_thread_sample_gib > 0:
        single = single_thread_compare(args, device)
    if rank == 0 and world == 1 and cpu is not None and not args.skip_e2e:
        same_sample = same_sample_compare(args, device, cpu)

    if world > 1:
        dist.barrier(device_ids=[device.index])
        dist.destroy_process_group()

    if rank != 0:
        return 0

    peak, peak_src = load_hbm_peak()
    window = kern["window_bytes"]
    verify_gbs = window / (kern["verify_ms_avg"] * 1e-3) / 1e9
    fill_gbs = window / (kern["fill_ms_avg"] * 1e-3) / 1e9
    rand_gbs = window / (kern["rand_ms_avg"] * 1e-3) / 1e9
    # dominant kernel = the one that takes the larger share of a step
    if kern["verify_ms_avg"] >= kern["fill_ms_avg"]:
        dom_name, dom_gbs = "elb_blocks_tiled_kernel<VERIFY_PATTERN> (K2)", verify_gbs
        dom_key = "K2_verify_pattern"
    else:
        dom_name, dom_gbs = "elb_blocks_tiled_kernel<FILL_PATTERN> (K1)", fill_gbs
        dom_key = "K1_fill_pattern"
    traffic, traffic_src = load_ncu_traffic(window, dom_key)

    line = {
        "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(elapsed_ms / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {
            "workload": "BASELINE configs[1]: single %.0f GiB file per GPU, %g MiB blocks, seq "
                        "write+read, --gpuids, cudaMemcpyAsync staging, --verify %d" % (
                            args.file_gib, args.block_mib, args.salt),
            "file_gib": args.file_gib, "file_gib_requested": args.file_gib_requested,
            "block_mib": args.block_mib, "threads_per_gpu": args.threads,
            "window_gib": window / GiB, "step": "K1 fill + K2 verify over the HBM-resident window "
                                               "(2 x window bytes)",
            "l2": "inputs_larger_than_l2 (window %.1f GiB >> 126 MB L2)" % (window / GiB),
            "storage_dir": args.dir, "direct": args.direct,
            "serialize_buffered_writes": not args.no_write_gate,
            "parallelism": "%d process(es) x %d worker threads, rank r <-> file r <-> GPU r" % (
                world, args.threads),
        },
        "gpu_launches": kern["launches"] * world,
        "roofline": {
            "bound": "hbm", "kernel": dom_name, "achieved": round(dom_gbs, 1), "peak": peak,
            "unit": "GB/s", "frac": round(dom_gbs / peak, 4), "traffic": traffic,
            "traffic_source": traffic_src,
            "peak_source": peak_src,
            "note": "peak is the driver-measured COPY bandwidth; one-directional streams exceed "
                    "it (copy-engine cudaMemset writes 7.35 TB/s), hence frac > 1",
            "algorithmic_bytes_per_launch": window,
            "all_kernels": {
                "K1_fill_pattern": {"achieved": round(fill_gbs, 1), "frac": round(fill_gbs / peak, 4),
                                    "ms_per_launch": round(kern["fill_ms_avg"], 4)},
                "K2_verify_pattern": {"achieved": round(verify_gbs, 1),
                                      "frac": round(verify_gbs / peak, 4),
                                      "ms_per_launch": round(kern["verify_ms_avg"], 4)},
                "K3_fill_random_pct100": {"achieved": round(rand_gbs, 1),
                                          "frac": round(rand_gbs / peak, 4),
                                          "ms_per_launch": round(kern["rand_ms_avg"], 4),
                                          "note": "outside the headline timed region"},
            },
        },
        "clocks": clocks,
    }
    if e2e:
        line["e2e"] = {
            "value": round(e2e["gib_s"], 3), "unit": UNIT,
            "h2d_bytes_per_step": e2e["h2d_bytes"], "d2h_bytes_per_step": e2e["d2h_bytes"],
            "step": "one write phase + one read phase (--verify) over the whole file(s)",
            "write_gib_s": round(e2e["write_gib_s"], 3), "read_gib_s": round(e2e["read_gib_s"], 3),
            "write_first_done_gib_s": round(e2e["write_first_done_gib_s"], 3),
            "read_first_done_gib_s": round(e2e["read_first_done_gib_s"], 3),
            "write_iops": e2e["write_iops"], "read_iops": e2e["read_iops"],
            "gpu_launches": e2e["launches"], "dev_kernel_usec": e2e["dev_kernel_usec"],
            "lat_write_avg_usec": round(e2e["lat_write_avg_usec"], 1),
            "lat_read_avg_usec": round(e2e["lat_read_avg_usec"], 1),
            "total_usec": e2e["total_usec"],
        }
    if cpu:
        line["cpu_baseline"] = cpu
    if storage and e2e:
        storage["e2e_frac"] = round(e2e["gib_s"] / storage["gib_s"], 3) if storage["gib_s"] else None
        storage["note"] = ("raw pread/pwrite of the CPU loop without fill/verify on the same "
                           "storage and thread count (bounded sample); buffered writes to ONE "
                           "file serialise on the inode lock, so the write phase does not scale "
                           "with threads on either arm")
        line["storage_roofline"] = storage
    if single:
        line["single_thread"] = single
    if same_sample:
        line["same_sample"] = same_sample
    emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
