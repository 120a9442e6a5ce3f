"""Round 2 exploration: the staged pipeline of the real worker on one GPU.

Sweeps, one factor at a time around the default (kernel staging, 1 MiB batches x 3, write gate
auto, GPU-affine NUMA binding): staging engine, batch shape, write gate, NUMA binding, thread
count; then 4 KiB random reads (sync and AIO 64) by batch size and staging engine; next to the CPU
LocalWorker (oracle port) and raw pread/pwrite on the same files.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elbencho_b200 import BenchPhase, IOEngine, WorkerConfig, WorkerManager  # noqa: E402

GiB = 1 << 30
MiB = 1 << 20
KiB = 1 << 10


def rate(res, key="bytes"):
    return round(res["ops_total"][key] / (res["last_finish_usec"] / 1e6) / (GiB if key == "bytes"
                                                                             else 1), 3)


def gpu_seq(path, size, threads, **kw):
    if os.path.exists(path):
        os.unlink(path)
    cfg = WorkerConfig(paths=[path], num_threads=threads, block_size=MiB, file_size=size,
                       integrity_check_salt=1, **kw)
    out = {}
    with WorkerManager(cfg) as mgr:
        w = mgr.run_phase(BenchPhase.CREATEFILES)
        r = mgr.run_phase(BenchPhase.READFILES)
    out["W"] = rate(w)
    out["R"] = rate(r)
    out["kern_ms_w"] = round(w["dev_kernel_usec"] / 1e3, 1)
    out["kern_ms_r"] = round(r["dev_kernel_usec"] / 1e3, 1)
    out["lat_w"] = round(w["iops_lat_histo"]["sum_usec"] / max(1, w["iops_lat_histo"]["num"]), 1)
    out["lat_r"] = round(r["iops_lat_histo"]["sum_usec"] / max(1, r["iops_lat_histo"]["num"]), 1)
    assert r["verify_mismatch_bytes"] == 0 and r["verified_bytes"] == size
    return out


def gpu_rand_read(path, size, threads, block, iodepth, **kw):
    cfg = WorkerConfig(paths=[path], num_threads=threads, block_size=block, file_size=size,
                       integrity_check_salt=1, use_random_offsets=True, rand_offset_seed=7,
                       io_depth=iodepth, io_engine=IOEngine.AIO if iodepth > 1 else IOEngine.SYNC,
                       **kw)
    with WorkerManager(cfg) as mgr:
        r = mgr.run_phase(BenchPhase.READFILES)
    assert r["verify_mismatch_bytes"] == 0
    return {"iops": int(rate(r, "iops")), "gib_s": rate(r),
            "kern_ms": round(r["dev_kernel_usec"] / 1e3, 1)}


def cpu_run(path, size, threads, block, rand=False):
    from tests import oracle_lib
    out = {}
    phases = (BenchPhase.READFILES,) if rand else (BenchPhase.CREATEFILES, BenchPhase.READFILES)
    if not rand and os.path.exists(path):
        os.unlink(path)
    cfg = WorkerConfig(paths=[path], num_threads=threads, block_size=block, file_size=size,
                       integrity_check_salt=1, use_random_offsets=rand, rand_offset_seed=7)
    for phase in phases:
        rc, workers, pres = oracle_lib.run_oracle_phase(cfg, phase)
        assert rc == 0
        secs = pres.lastFinishUSec / 1e6
        out[phase.name[0]] = round(pres.opsTotal.numBytesDone / GiB / secs, 3)
        out[phase.name[0] + "_iops"] = int(pres.opsTotal.numIOPSDone / secs)
    return out


def emit(**kw):
    print(json.dumps(kw), flush=True)


def main():
    base = sys.argv[1] if len(sys.argv) > 1 else "/dev/shm"
    size = int(float(sys.argv[2]) * GiB) if len(sys.argv) > 2 else 16 * GiB
    path = os.path.join(base, "elb_explore_r2.bin")
    t0 = time.time()
    gpu_seq(path, 1 * GiB, 4)  # warm-up: context, kernels

    default = dict(staging_engine=0, pipeline_batch_blocks=0, pipeline_num_batches=0,
                   serialize_buffered_writes=0, no_gpu_numa_binding=False)

    def seq(label, threads=16, **over):
        kw = dict(default)
        kw.update(over)
        res = gpu_seq(path, size, threads, **kw)
        emit(test="seq", label=label, threads=threads, **over, **res)

    seq("default")
    seq("default_again")
    seq("staging_ce", staging_engine=2)
    for bb, nb in ((1, 2), (1, 4), (1, 6), (2, 2), (2, 3), (4, 2), (16, 2)):
        seq("batch", pipeline_batch_blocks=bb, pipeline_num_batches=nb)
    seq("batch_ce_16x2", staging_engine=2, pipeline_batch_blocks=16, pipeline_num_batches=2)
    seq("gate_off", serialize_buffered_writes=2)
    seq("gate_on", serialize_buffered_writes=1)
    seq("numa_unbound", no_gpu_numa_binding=True)
    seq("numa_remote", numa_zones=[1])
    for threads in (1, 4, 8, 12, 24, 32):
        seq("threads", threads=threads)
    emit(test="cpu_seq", threads=16, **cpu_run(path + ".cpu", size // 2, 16, MiB))
    emit(test="cpu_seq", threads=8, **cpu_run(path + ".cpu", size // 2, 8, MiB))
    if os.path.exists(path + ".cpu"):
        os.unlink(path + ".cpu")

    # random 4 KiB reads on the file of the last sequential run
    seq("default_for_rand")
    for block in (4 * KiB, 64 * KiB):
        for iodepth in (1, 64):
            for label, over in (("default", {}), ("staging_ce", dict(staging_engine=2)),
                                ("bb64", dict(pipeline_batch_blocks=64)),
                                ("bb1024", dict(pipeline_batch_blocks=1024)),
                                ("bb2048x2", dict(pipeline_batch_blocks=2048,
                                                  pipeline_num_batches=2)),
                                ("ce_bb2048x2", dict(staging_engine=2, pipeline_batch_blocks=2048,
                                                     pipeline_num_batches=2))):
                kw = dict(default)
                kw.update(over)
                res = gpu_rand_read(path, size, 16, block, iodepth, **kw)
                emit(test="rand_read", block=block, iodepth=iodepth, label=label, **res)
        emit(test="cpu_rand_read", block=block, threads=16, **cpu_run(path, size, 16, block, True))
    os.unlink(path)
    emit(test="done", secs=round(time.time() - t0, 1))


if __name__ == "__main__":
    main()
