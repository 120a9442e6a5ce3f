"""Live statistics across GPUs: the snapshot that sums LiveOps / LiveLatency and
the device-resident kernel counters over all workers must equal the exact per-worker values once a
phase is done, on one GPU (gather kernel only) and on two or more GPUs (gather kernel + one
grouped ncclReduce to the first GPU). Reference: host-side sum in Statistics.cpp:414-470."""
import os
import shutil
import tempfile
import threading

import pytest

from elbencho_b200 import BenchPhase, WorkerConfig, WorkerManager
from elbencho_b200._native import DEVCTR_NUM

pytestmark = pytest.mark.gpu

MiB = 1 << 20
KiB = 1 << 10

DEVCTR_MISMATCH, DEVCTR_VERIFIED, DEVCTR_FILLED = 0, 1, 2


@pytest.fixture()
def workdir(cuda_device):
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    path = tempfile.mkdtemp(prefix="elb_live_", dir=base)
    yield path
    shutil.rmtree(path, ignore_errors=True)


def num_gpus():
    import torch
    return torch.cuda.device_count()


def check_snapshot_equals_workers(mgr, snap):
    ops = {"entries": 0, "bytes": 0, "iops": 0}
    mix = dict(ops)
    ctrs = [0] * DEVCTR_NUM
    for worker in mgr.workers():
        wops, wmix = worker.live_ops()
        for key in ops:
            ops[key] += wops[key]
            mix[key] += wmix[key]
        for i, val in enumerate(worker.dev_counters()):
            ctrs[i] += val
    assert snap["ops"] == ops
    assert snap["ops_readmix"] == mix
    assert snap["dev_counters"] == ctrs
    assert snap["num_workers_total"] == len(mgr.workers())
    assert snap["num_workers_done"] == len(mgr.workers())


def merged_worker_histogram(mgr, kind):
    """LatencyHistogram::operator+= over the workers (LatencyHistogram.h:187-202)"""
    out = {"buckets": None, "num": 0, "sum_usec": 0, "min_usec": (1 << 64) - 1, "max_usec": 0}
    for worker in mgr.workers():
        histo = worker.histogram(kind)
        out["buckets"] = histo["buckets"] if out["buckets"] is None else \
            [a + b for a, b in zip(out["buckets"], histo["buckets"])]
        out["num"] += histo["num"]
        out["sum_usec"] += histo["sum_usec"]
        out["min_usec"] = min(out["min_usec"], histo["min_usec"])
        out["max_usec"] = max(out["max_usec"], histo["max_usec"])
    return out


def check_phase_results_reduced_with_nccl(mgr, res, num_bytes_key, num_bytes):
    assert res["stats_reduced_with_nccl"]
    assert res["iops_lat_histo"] == merged_worker_histogram(mgr, 0)
    assert res["iops_lat_histo_readmix"] == merged_worker_histogram(mgr, 1)
    assert res["entries_lat_histo"] == merged_worker_histogram(mgr, 2)
    assert res["entries_lat_histo_readmix"] == merged_worker_histogram(mgr, 3)
    assert res["iops_lat_histo"]["num"] == res["ops_total"]["iops"] > 0
    assert res["iops_lat_histo"]["min_usec"] <= res["iops_lat_histo"]["max_usec"]
    assert res[num_bytes_key] == num_bytes
    assert res["verify_mismatch_bytes"] == 0


def run_with_polling(mgr, phase):
    """start the phase, poll the snapshot from a second thread until done"""
    seen = []
    stop = threading.Event()

    def poll():
        while not stop.is_set():
            seen.append(mgr.live_snapshot())
            stop.wait(0.002)

    mgr.start_phase(phase)
    poller = threading.Thread(target=poll)
    poller.start()
    try:
        mgr.wait_done(-1)
    finally:
        stop.set()
        poller.join()
    return seen


def test_single_gpu_snapshot_gathers_on_device(workdir):
    size, block, threads, salt = 64 * MiB, 256 * KiB, 4, 41
    path = os.path.join(workdir, "f")
    cfg = WorkerConfig(paths=[path], num_threads=threads, block_size=block, file_size=size,
                       integrity_check_salt=salt)
    with WorkerManager(cfg) as mgr:
        seen = run_with_polling(mgr, BenchPhase.CREATEFILES)
        res = mgr.phase_results()
        snap = mgr.live_snapshot()
        assert snap["num_gpus"] == 1
        assert not snap["reduced_with_nccl"]
        assert snap["gathered_on_device"]
        check_snapshot_equals_workers(mgr, snap)
        assert snap["ops"]["bytes"] == size == res["ops_total"]["bytes"]
        assert snap["dev_counters"][DEVCTR_FILLED] == size
        # live values only grow within a phase
        prev = 0
        for item in seen:
            assert item["ops"]["bytes"] >= prev
            assert item["ops"]["bytes"] <= size
            prev = item["ops"]["bytes"]
        # the live latency counters are consumed: sum over all snapshots == number of I/Os
        num_lat = sum(item["lat"]["numAvgIOLatValues"] for item in seen) + \
            snap["lat"]["numAvgIOLatValues"]
        assert num_lat == size // block

        seen = run_with_polling(mgr, BenchPhase.READFILES)
        snap = mgr.live_snapshot()
        check_snapshot_equals_workers(mgr, snap)
        assert snap["dev_counters"][DEVCTR_VERIFIED] == size
        assert snap["dev_counters"][DEVCTR_MISMATCH] == 0


def test_single_gpu_snapshot_carries_mismatch_count(workdir):
    size, block, salt = 8 * MiB, 64 * KiB, 5
    path = os.path.join(workdir, "f")
    cfg = WorkerConfig(paths=[path], num_threads=2, block_size=block, file_size=size,
                       integrity_check_salt=salt, verify_collect_all=True)
    with WorkerManager(cfg) as mgr:
        mgr.run_phase(BenchPhase.CREATEFILES)
        with open(path, "r+b") as f:  # flip 3 bytes in different blocks
            for off in (17, 3 * MiB + 5, 7 * MiB + 4095):
                f.seek(off)
                byte = f.read(1)
                f.seek(off)
                f.write(bytes([byte[0] ^ 0xFF]))
        mgr.start_phase(BenchPhase.READFILES)
        try:
            mgr.wait_done(-1)
        except Exception:
            pass  # the verify error is the expected phase outcome
        snap = mgr.live_snapshot()
        assert snap["dev_counters"][DEVCTR_MISMATCH] == 3


@pytest.mark.parametrize("threads_per_gpu", [1, 3])
def test_multi_gpu_snapshot_reduced_with_nccl(workdir, threads_per_gpu):
    ngpus = num_gpus()
    if ngpus < 2:
        pytest.skip("needs >= 2 GPUs")
    size, block, salt = 256 * MiB, 1 * MiB, 77
    threads = threads_per_gpu * ngpus
    path = os.path.join(workdir, "f")
    cfg = WorkerConfig(paths=[path], num_threads=threads, block_size=block, file_size=size,
                       integrity_check_salt=salt, gpu_ids=tuple(range(ngpus)))
    with WorkerManager(cfg) as mgr:
        assert sorted({w.gpu_id for w in mgr.workers()}) == list(range(ngpus))
        info = mgr.live_reduce_info()
        assert "NCCL" in info and "root GPU 0" in info, info
        seen = run_with_polling(mgr, BenchPhase.CREATEFILES)
        snap = mgr.live_snapshot()
        assert snap["num_gpus"] == ngpus
        assert snap["reduced_with_nccl"], mgr.live_reduce_info()
        assert snap["gathered_on_device"]
        check_snapshot_equals_workers(mgr, snap)
        expected_bytes = size  # (the last rank takes the remainder blocks, LocalWorker.cpp:3022-3060)
        assert snap["ops"]["bytes"] == expected_bytes
        assert snap["dev_counters"][DEVCTR_FILLED] == expected_bytes
        assert all(item["reduced_with_nccl"] for item in seen)
        check_phase_results_reduced_with_nccl(mgr, mgr.phase_results(), "filled_bytes",
                                              expected_bytes)
        prev = 0
        for item in seen:
            assert prev <= item["ops"]["bytes"] <= expected_bytes
            prev = item["ops"]["bytes"]

        seen = run_with_polling(mgr, BenchPhase.READFILES)
        snap = mgr.live_snapshot()
        check_snapshot_equals_workers(mgr, snap)
        assert snap["dev_counters"][DEVCTR_VERIFIED] == expected_bytes
        assert snap["dev_counters"][DEVCTR_MISMATCH] == 0
        assert snap["reduced_with_nccl"]
        check_phase_results_reduced_with_nccl(mgr, mgr.phase_results(), "verified_bytes",
                                              expected_bytes)


def test_cli_live_line_over_all_gpus(workdir):
    """the console live line of the front end (Statistics.cpp:180-285) fed from the reducer"""
    import subprocess
    from elbencho_b200.build import CLI_PATH
    ngpus = num_gpus()
    path = os.path.join(workdir, "cli.bin")
    env = dict(os.environ, ELB_FORCE_LIVESTATS="1")
    gpuids = ",".join(str(i) for i in range(min(ngpus, 2)))
    proc = subprocess.run([CLI_PATH, "-w", "-r", "-t", "4", "-b", "1M", "-s", "2G", "--verify", "3",
                           "--gpuids", gpuids, "--liveint", "50", path],
                          capture_output=True, text=True, timeout=300, env=env)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    live_lines = [seg for seg in proc.stdout.split("\r") if "MiB/s;" in seg and "IOPS;" in seg]
    assert live_lines, proc.stdout
    assert "WRITE" in proc.stdout and "READ" in proc.stdout
    assert "NOTE:" not in proc.stderr, proc.stderr  # NCCL (or single GPU) path without fallback
