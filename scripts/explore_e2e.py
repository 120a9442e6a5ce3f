"""Exploration (not a bench line): thread scaling of the CPU LocalWorker oracle vs the GPU worker
on this box's storage, with and without --preallocfile, plus the raw pread/pwrite ceiling."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elbencho_b200 import BenchPhase, WorkerConfig, WorkerManager  # noqa: E402
from tests import oracle_lib  # noqa: E402

GiB = 1 << 30
MiB = 1 << 20


def cpu_run(path, threads, size, salt, prealloc, direct=False):
    if os.path.exists(path):
        os.unlink(path)
    cfg = WorkerConfig(paths=[path], num_threads=threads, block_size=MiB, file_size=size,
                       integrity_check_salt=salt, do_prealloc_file=prealloc, use_direct_io=direct)
    out = {}
    for phase in (BenchPhase.CREATEFILES, BenchPhase.READFILES):
        rc, workers, pres = oracle_lib.run_oracle_phase(cfg, phase)
        assert rc == 0, [w.errorMsg for w in workers]
        out[phase.name] = round(pres.opsTotal.numBytesDone / GiB / (pres.lastFinishUSec / 1e6), 2)
    os.unlink(path)
    return out


def gpu_run(path, threads, size, salt, prealloc, batch=0, nbatches=0, direct=False):
    if os.path.exists(path):
        os.unlink(path)
    cfg = WorkerConfig(paths=[path], num_threads=threads, block_size=MiB, file_size=size,
                       integrity_check_salt=salt, do_prealloc_file=prealloc,
                       pipeline_batch_blocks=batch, pipeline_num_batches=nbatches,
                       use_direct_io=direct)
    out = {}
    with WorkerManager(cfg) as mgr:
        for phase in (BenchPhase.CREATEFILES, BenchPhase.READFILES):
            res = mgr.run_phase(phase)
            out[phase.name] = round(
                res["ops_total"]["bytes"] / GiB / (res["last_finish_usec"] / 1e6), 2)
    os.unlink(path)
    return out


def main():
    base = sys.argv[1] if len(sys.argv) > 1 else "/dev/shm"
    size = int(float(sys.argv[2]) * GiB) if len(sys.argv) > 2 else 16 * GiB
    path = os.path.join(base, "elb_explore.bin")
    # warm the GPU path
    gpu_run(path, 2, 1 * GiB, 1, False)
    for prealloc in (False, True):
        for threads in (1, 4, 16, 64):
            t0 = time.time()
            raw = cpu_run(path, threads, size, 0, prealloc)
            cpu = cpu_run(path, threads, size, 1, prealloc)
            row = {"dir": base, "prealloc": prealloc, "threads": threads, "raw": raw, "cpu": cpu}
            if threads <= 32:
                row["gpu"] = gpu_run(path, threads, size, 1, prealloc)
            row["secs"] = round(time.time() - t0, 1)
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
