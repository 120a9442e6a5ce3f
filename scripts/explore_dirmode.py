"""Exploration: where does the per-file time of directory mode (64 KiB files) go?
files/s of GPU worker vs CPU LocalWorker, with and without fill/verify, at 1 and 16 threads."""
import json
import os
import shutil
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elbencho_b200 import BenchPhase, PathType, WorkerConfig, WorkerManager  # noqa: E402
from tests import oracle_lib  # noqa: E402

KiB = 1 << 10


def main():
    base = sys.argv[1] if len(sys.argv) > 1 else "/dev/shm"
    root = os.path.join(base, "elb_explore_dir")
    for threads in (1, 16):
        for label, extra in (("verify", dict(integrity_check_salt=1)), ("nofill", {}),
                             ("blockvar", dict(block_variance_percent=100, block_variance_seed=3))):
            row = {"threads": threads, "mode": label}
            for arm in ("gpu", "cpu"):
                shutil.rmtree(root, ignore_errors=True)
                os.makedirs(root)
                cfg = WorkerConfig(paths=[root], path_type=PathType.DIR, num_threads=threads,
                                   num_dirs=4, num_files=1024, block_size=64 * KiB,
                                   file_size=64 * KiB, **extra)
                if arm == "gpu":
                    with WorkerManager(cfg) as mgr:
                        mgr.run_phase(BenchPhase.CREATEDIRS)
                        w = mgr.run_phase(BenchPhase.CREATEFILES)
                        r = mgr.run_phase(BenchPhase.READFILES)
                    row["gpu_write_files_s"] = w["ops_per_sec"]["entries"]
                    row["gpu_read_files_s"] = r["ops_per_sec"]["entries"]
                    row["gpu_write_lat_us"] = round(w["entries_lat_histo"]["sum_usec"] /
                                                    max(1, w["entries_lat_histo"]["num"]), 1)
                else:
                    out = {}
                    for phase in (BenchPhase.CREATEDIRS, BenchPhase.CREATEFILES,
                                  BenchPhase.READFILES):
                        rc, workers, pres = oracle_lib.run_oracle_phase(cfg, phase)
                        assert rc == 0
                        out[phase.name] = pres.opsPerSec.numEntriesDone
                    row["cpu_write_files_s"] = out["CREATEFILES"]
                    row["cpu_read_files_s"] = out["READFILES"]
            print(json.dumps(row), flush=True)
    shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
