"""Secondary measurements of the other BASELINE.json configs (3, 4, 5) at one GPU, both arms.
Not the bench line (bench.py measures configs[1]); results go to profiles/ as evidence.

  C3: single file, 4 KiB random reads, iodepth 64, --verify          (IOPS)
  C4: one 32 GiB file per GPU written with --blockvarpct 100, 1 MiB sequential read (GiB/s)
  C5: directory tree of 64 KiB files, 16 threads per GPU, write + read --verify (files/s)

GDS cannot be used on the graft boxes (cuFileHandleRegister fails), so C3/C4 run
on the staged path (kernel AIO / pread) here.
"""
import json
import os
import shutil
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elbencho_b200 import BenchPhase, PathType, WorkerConfig, WorkerManager  # noqa: E402
from tests import oracle_lib  # noqa: E402

GiB = 1 << 30
MiB = 1 << 20
KiB = 1 << 10


def gpu_phase(mgr, phase):
    res = mgr.run_phase(phase)
    secs = res["last_finish_usec"] / 1e6
    return {"gib_s": round(res["ops_total"]["bytes"] / GiB / secs, 3),
            "iops": res["ops_per_sec"]["iops"], "entries_s": res["ops_per_sec"]["entries"],
            "secs": round(secs, 3), "launches": res["num_kernel_launches"],
            "kernel_usec": res["dev_kernel_usec"],
            "lat_avg_usec": round(res["iops_lat_histo"]["sum_usec"] /
                                  max(1, res["iops_lat_histo"]["num"]), 1)}


def cpu_phase(cfg, phase):
    rc, workers, pres = oracle_lib.run_oracle_phase(cfg, phase)
    assert rc == 0, [w.errorMsg for w in workers if w.hadError]
    secs = pres.lastFinishUSec / 1e6
    return {"gib_s": round(pres.opsTotal.numBytesDone / GiB / secs, 3),
            "iops": pres.opsPerSec.numIOPSDone, "entries_s": pres.opsPerSec.numEntriesDone,
            "secs": round(secs, 3)}


def main():
    base = sys.argv[1] if len(sys.argv) > 1 else "/dev/shm"
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    workdir = os.path.join(base, "elb_cfgs_%d" % os.getuid())
    shutil.rmtree(workdir, ignore_errors=True)
    os.makedirs(workdir)
    out = []
    try:
        # ---- C3: 4 KiB random reads, iodepth 64 ----
        size = int(16 * GiB * scale)
        path = os.path.join(workdir, "c3.bin")
        with WorkerManager(WorkerConfig(paths=[path], num_threads=16, block_size=MiB,
                                        file_size=size, integrity_check_salt=1,
                                        serialize_buffered_writes=True)) as mgr:
            mgr.run_phase(BenchPhase.CREATEFILES)
        amount = int(4 * GiB * scale)
        for threads in (1, 16):
            rnd = dict(num_threads=threads, block_size=4 * KiB, file_size=size,
                       integrity_check_salt=1, use_random_offsets=True, random_amount=amount,
                       rand_offset_seed=42)
            with WorkerManager(WorkerConfig(paths=[path], io_depth=64, **rnd)) as mgr:
                gpu = gpu_phase(mgr, BenchPhase.READFILES)
            with WorkerManager(WorkerConfig(paths=[path], io_depth=1, **rnd)) as mgr:
                gpu_sync = gpu_phase(mgr, BenchPhase.READFILES)
            cpu = cpu_phase(WorkerConfig(paths=[path], **rnd), BenchPhase.READFILES)
            out.append({"config": "C3 4KiB rand read --verify, %.0f GiB file, %.0f GiB amount" % (
                size / GiB, amount / GiB), "threads": threads, "gpu_aio_iodepth64": gpu,
                "gpu_sync_iodepth1": gpu_sync, "cpu_localworker_sync": cpu})
            print(json.dumps(out[-1]), flush=True)
        os.unlink(path)

        # ---- C4: blockvarpct 100 write (K3), then 1 MiB sequential read ----
        size = int(32 * GiB * scale)
        path = os.path.join(workdir, "c4.bin")
        for threads in (1, 16):
            cfg = dict(num_threads=threads, block_size=MiB, file_size=size)
            with WorkerManager(WorkerConfig(paths=[path], block_variance_percent=100,
                                            block_variance_seed=7, serialize_buffered_writes=True,
                                            **cfg)) as mgr:
                gpu_w = gpu_phase(mgr, BenchPhase.CREATEFILES)
                gpu_r = gpu_phase(mgr, BenchPhase.READFILES)
            cpu_w = cpu_phase(WorkerConfig(paths=[path], block_variance_percent=100, **cfg),
                              BenchPhase.CREATEFILES)
            cpu_r = cpu_phase(WorkerConfig(paths=[path], **cfg), BenchPhase.READFILES)
            out.append({"config": "C4 one %.0f GiB file, --blockvarpct 100 write, 1 MiB seq read" % (
                size / GiB), "threads": threads, "gpu_write": gpu_w, "gpu_read": gpu_r,
                "cpu_write": cpu_w, "cpu_read": cpu_r})
            print(json.dumps(out[-1]), flush=True)
        os.unlink(path)

        # ---- C5: dir tree of 64 KiB files, 16 threads per GPU ----
        ndirs, nfiles = 8, max(1, int(128 * scale))
        for arm in ("gpu", "cpu"):
            tree = os.path.join(workdir, "c5_" + arm)
            os.makedirs(tree)
            cfg = WorkerConfig(paths=[tree], path_type=PathType.DIR, num_threads=16,
                               num_dirs=ndirs, num_files=nfiles, block_size=64 * KiB,
                               file_size=64 * KiB, integrity_check_salt=1)
            row = {}
            if arm == "gpu":
                with WorkerManager(cfg) as mgr:
                    for phase in (BenchPhase.CREATEDIRS, BenchPhase.CREATEFILES,
                                  BenchPhase.READFILES, BenchPhase.DELETEFILES,
                                  BenchPhase.DELETEDIRS):
                        row[phase.name] = gpu_phase(mgr, phase)
            else:
                for phase in (BenchPhase.CREATEDIRS, BenchPhase.CREATEFILES,
                              BenchPhase.READFILES, BenchPhase.DELETEFILES, BenchPhase.DELETEDIRS):
                    row[phase.name] = cpu_phase(cfg, phase)
            out.append({"config": "C5 dir tree 16 threads x %d dirs x %d files x 64 KiB, "
                                  "write+read --verify" % (ndirs, nfiles), "arm": arm, **row})
            print(json.dumps(out[-1]), flush=True)
    finally:
        shutil.rmtree(workdir, ignore_errors=True)


if __name__ == "__main__":
    main()
