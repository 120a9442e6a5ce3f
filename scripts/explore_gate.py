"""Exploration: GPU worker write/read GiB/s vs threads, write gate and batch size (single file)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elbencho_b200 import BenchPhase, WorkerConfig, WorkerManager  # noqa: E402

GiB = 1 << 30
MiB = 1 << 20


def gpu_run(path, threads, size, gate, batch, nbatches, prealloc=False):
    if os.path.exists(path):
        os.unlink(path)
    cfg = WorkerConfig(paths=[path], num_threads=threads, block_size=MiB, file_size=size,
                       integrity_check_salt=1, serialize_buffered_writes=gate,
                       pipeline_batch_blocks=batch, pipeline_num_batches=nbatches,
                       do_prealloc_file=prealloc)
    out = {}
    with WorkerManager(cfg) as mgr:
        for phase in (BenchPhase.CREATEFILES, BenchPhase.READFILES):
            res = mgr.run_phase(phase)
            out[phase.name[:1]] = round(
                res["ops_total"]["bytes"] / GiB / (res["last_finish_usec"] / 1e6), 2)
    os.unlink(path)
    return out


def main():
    base = sys.argv[1] if len(sys.argv) > 1 else "/dev/shm"
    size = int(float(sys.argv[2]) * GiB) if len(sys.argv) > 2 else 16 * GiB
    path = os.path.join(base, "elb_explore_gate.bin")
    gpu_run(path, 2, 1 * GiB, False, 0, 0)
    for threads in (1, 2, 4, 8, 16, 32):
        for gate in (False, True):
            for batch, nbatches in ((16, 2), (4, 2), (2, 3)):
                if threads == 1 and gate:
                    continue
                res = gpu_run(path, threads, size, gate, batch, nbatches)
                print(json.dumps({"t": threads, "gate": gate, "batch": batch, "nb": nbatches,
                                  **res}), flush=True)


if __name__ == "__main__":
    main()
