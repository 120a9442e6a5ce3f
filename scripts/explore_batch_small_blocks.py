"""Exploration: 4 KiB random reads --verify, IOPS vs staging batch size (pipeline_batch_blocks) and
thread count on one GPU. Is a cache-resident staging ring worth more than fewer GPU launches?"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elbencho_b200 import BenchPhase, WorkerConfig, WorkerManager  # noqa: E402
from tests import oracle_lib  # noqa: E402

GiB = 1 << 30
MiB = 1 << 20
KiB = 1 << 10


def main():
    base = sys.argv[1] if len(sys.argv) > 1 else "/dev/shm"
    size = 16 * GiB
    path = os.path.join(base, "elb_explore_batch.bin")
    with WorkerManager(WorkerConfig(paths=[path], num_threads=16, block_size=MiB, file_size=size,
                                    integrity_check_salt=1, serialize_buffered_writes=True)) as mgr:
        mgr.run_phase(BenchPhase.CREATEFILES)
    try:
        for block in (4 * KiB, 64 * KiB):
            for threads in (16, 64, 128):
                amount = (2 * GiB if block == 4 * KiB else 8 * GiB) * threads // 16
                amount = min(amount, 16 * GiB)
                rnd = dict(num_threads=threads, block_size=block, file_size=size,
                           integrity_check_salt=1, use_random_offsets=True, random_amount=amount,
                           rand_offset_seed=42)
                row = {"block": block, "threads": threads}
                for batch_bytes in (0, 8 * MiB, 2 * MiB, 512 * KiB, 128 * KiB):
                    bb = batch_bytes // block
                    for nb in (2, 4):
                        if batch_bytes == 0 and nb != 2:
                            continue
                        with WorkerManager(WorkerConfig(paths=[path], pipeline_batch_blocks=bb,
                                                        pipeline_num_batches=nb if bb else 0,
                                                        **rnd)) as mgr:
                            res = mgr.run_phase(BenchPhase.READFILES)
                        key = "default" if not batch_bytes else "%dK x%d" % (batch_bytes // KiB, nb)
                        row[key] = round(res["ops_per_sec"]["iops"] / 1e6, 2)
                rc, workers, pres = oracle_lib.run_oracle_phase(WorkerConfig(paths=[path], **rnd),
                                                                BenchPhase.READFILES)
                row["cpu"] = round(pres.opsPerSec.numIOPSDone / 1e6, 2)
                print(json.dumps(row), flush=True)
    finally:
        os.unlink(path)


if __name__ == "__main__":
    main()
