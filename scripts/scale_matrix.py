"""Scaling matrix on ONE box with several GPUs: the in-process worker pool (--gpuids 0..N-1, one file
per GPU, 16 threads per GPU, live/phase-end statistics through the NCCL reduce) at N = 1, 2, 4, 8
for sequential write, sequential read --verify, 4 KiB random read --verify (sync and kernel AIO
iodepth 64) and random 4 KiB write, next to the CPU LocalWorker (oracle port) with the same thread
count and the raw pread/pwrite storage roofline. Evidence for profiles/, not the bench line.

    python scripts/scale_matrix.py [dir] [file_gib_per_gpu] [rand_amount_gib_per_gpu]
"""
import json
import os
import shutil
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elbencho_b200 import BenchPhase, WorkerConfig, WorkerManager  # noqa: E402
from tests import oracle_lib  # noqa: E402

GiB = 1 << 30
MiB = 1 << 20
KiB = 1 << 10
THREADS_PER_GPU = 16


def gpu_phase(mgr, phase):
    res = mgr.run_phase(phase)
    secs = res["last_finish_usec"] / 1e6
    return {"gib_s": round(res["ops_total"]["bytes"] / GiB / secs, 2),
            "iops": res["ops_per_sec"]["iops"], "secs": round(secs, 3),
            "nccl_stats": res["stats_reduced_with_nccl"],
            "mismatch_bytes": res["verify_mismatch_bytes"]}


def cpu_phase(cfg, phase):
    rc, workers, pres = oracle_lib.run_oracle_phase(cfg, phase)
    assert rc == 0, [w.errorMsg for w in workers if w.hadError]
    secs = pres.lastFinishUSec / 1e6
    return {"gib_s": round(pres.opsTotal.numBytesDone / GiB / secs, 2),
            "iops": pres.opsPerSec.numIOPSDone, "secs": round(secs, 3)}


def raw_storage(paths, threads_per_file, size, block, write):
    """raw pread/pwrite of the same block stream, no fill/verify/GPU: the storage roofline"""
    buf = bytearray(os.urandom(block))
    total_threads = threads_per_file * len(paths)
    share = size // threads_per_file
    start = threading.Barrier(total_threads + 1)
    done = []

    def work(path, rank):
        fd = os.open(path, os.O_RDWR | os.O_CREAT, 0o644)
        view = memoryview(buf)
        start.wait()
        off, end = rank * share, (rank + 1) * share
        while off < end:
            if write:
                os.pwrite(fd, view, off)
            else:
                os.preadv(fd, [view], off)
            off += block
        os.close(fd)
        done.append(time.perf_counter())

    workers = [threading.Thread(target=work, args=(path, rank))
               for path in paths for rank in range(threads_per_file)]
    for worker in workers:
        worker.start()
    start.wait()
    t0 = time.perf_counter()
    for worker in workers:
        worker.join()
    secs = max(done) - t0
    return round(len(paths) * share * threads_per_file / GiB / secs, 2)


def main():
    import torch
    base = sys.argv[1] if len(sys.argv) > 1 else "/dev/shm"
    file_gib = float(sys.argv[2]) if len(sys.argv) > 2 else 16.0
    rand_gib = float(sys.argv[3]) if len(sys.argv) > 3 else 2.0
    ngpus_avail = torch.cuda.device_count()
    workdir = os.path.join(base, "elb_scale_%d" % os.getuid())
    shutil.rmtree(workdir, ignore_errors=True)
    os.makedirs(workdir)
    size = int(file_gib * GiB)
    try:
        for ngpus in [n for n in (1, 2, 4, 8) if n <= ngpus_avail]:
            threads = THREADS_PER_GPU * ngpus
            gpu_ids = tuple(range(ngpus))
            paths = [os.path.join(workdir, "f%d.bin" % i) for i in range(ngpus)]
            cpaths = [os.path.join(workdir, "c%d.bin" % i) for i in range(ngpus)]
            row = {"n_gpus": ngpus, "threads": threads, "file_gib_per_gpu": file_gib}
            seq = dict(num_threads=threads, block_size=MiB, file_size=size, integrity_check_salt=1)
            with WorkerManager(WorkerConfig(paths=paths, gpu_ids=gpu_ids,
                                            serialize_buffered_writes=True, **seq)) as mgr:
                row["reduce"] = mgr.live_reduce_info()
                row["gpu_seq_write"] = gpu_phase(mgr, BenchPhase.CREATEFILES)
                row["gpu_seq_read_verify"] = gpu_phase(mgr, BenchPhase.READFILES)
            amount = int(rand_gib * GiB) * ngpus
            rnd = dict(num_threads=threads, block_size=4 * KiB, file_size=size,
                       integrity_check_salt=1, use_random_offsets=True, random_amount=amount,
                       rand_offset_seed=42)
            with WorkerManager(WorkerConfig(paths=paths, gpu_ids=gpu_ids, io_depth=1,
                                            **rnd)) as mgr:
                row["gpu_rand_read_4k_sync"] = gpu_phase(mgr, BenchPhase.READFILES)
            with WorkerManager(WorkerConfig(paths=paths, gpu_ids=gpu_ids, io_depth=64,
                                            **rnd)) as mgr:
                row["gpu_rand_read_4k_aio64"] = gpu_phase(mgr, BenchPhase.READFILES)
            rndw = dict(rnd, integrity_check_salt=0, block_variance_percent=100,
                        block_variance_seed=5)
            with WorkerManager(WorkerConfig(paths=paths, gpu_ids=gpu_ids, **rndw)) as mgr:
                row["gpu_rand_write_4k_blockvar100"] = gpu_phase(mgr, BenchPhase.CREATEFILES)
            # CPU LocalWorker with the same thread count on files of a quarter of the size
            csize = size // 4
            cseq = dict(seq, file_size=csize)
            row["cpu_sample_gib_per_file"] = csize / GiB
            row["cpu_seq_write"] = cpu_phase(WorkerConfig(paths=cpaths, **cseq),
                                             BenchPhase.CREATEFILES)
            row["cpu_seq_read_verify"] = cpu_phase(WorkerConfig(paths=cpaths, **cseq),
                                                   BenchPhase.READFILES)
            crnd = dict(rnd, file_size=csize, random_amount=amount // 2)
            row["cpu_rand_read_4k_sync"] = cpu_phase(WorkerConfig(paths=cpaths, **crnd),
                                                     BenchPhase.READFILES)
            for path in cpaths:
                os.unlink(path)
            # storage roofline: raw pwrite (new files) / pread of the same shape
            row["raw_seq_write_gib_s"] = raw_storage(cpaths, THREADS_PER_GPU, csize, MiB, True)
            row["raw_seq_read_gib_s"] = raw_storage(cpaths, THREADS_PER_GPU, csize, MiB, False)
            for path in paths + cpaths:
                if os.path.exists(path):
                    os.unlink(path)
            print(json.dumps(row), flush=True)
    finally:
        shutil.rmtree(workdir, ignore_errors=True)


if __name__ == "__main__":
    main()
