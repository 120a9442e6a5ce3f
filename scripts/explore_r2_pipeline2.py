"""Round 2 exploration, second pass: just-in-time gated writes and batch shape (one GPU)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.explore_r2_pipeline import GiB, MiB, cpu_run, emit, gpu_seq  # noqa: E402


def main():
    base = sys.argv[1] if len(sys.argv) > 1 else "/dev/shm"
    size = int(float(sys.argv[2]) * GiB) if len(sys.argv) > 2 else 16 * GiB
    path = os.path.join(base, "elb_explore_r2b.bin")
    gpu_seq(path, 1 * GiB, 4)

    def seq(label, threads=16, **over):
        res = gpu_seq(path, size, threads, **over)
        emit(test="seq", label=label, threads=threads, **over, **res)

    for rep in range(2):
        seq("default")
    seq("gate_off", serialize_buffered_writes=2)
    for bb, nb in ((1, 3), (2, 2), (2, 3), (4, 2), (4, 3), (8, 2)):
        seq("batch", pipeline_batch_blocks=bb, pipeline_num_batches=nb)
    seq("ce", staging_engine=2)
    seq("ce_4x2", staging_engine=2, pipeline_batch_blocks=4, pipeline_num_batches=2)
    for threads in (2, 4, 8, 12):
        seq("threads", threads=threads)
    seq("unbound", no_gpu_numa_binding=True)
    for threads in (4, 8, 16):
        emit(test="cpu_seq", threads=threads, **cpu_run(path + ".cpu", size // 2, threads, MiB))
        os.unlink(path + ".cpu")
    os.unlink(path)


if __name__ == "__main__":
    main()
