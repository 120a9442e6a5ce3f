"""N>1 path on CPU: world_size-2 gloo run of the rank layout + stats reduce that bench.py uses
under torchrun (one process per GPU; the data path has no collective)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fake_results(rank):
    """per-rank phase results as WorkerManager.phase_results() returns them"""
    def histo(base):
        buckets = [0] * 112
        buckets[10 + rank] = base
        buckets[111] = rank
        return {"buckets": buckets, "num": base + rank, "sum_usec": 1000 * (rank + 1),
                "min_usec": 5 + rank, "max_usec": 100 * (rank + 1)}
    return {
        "first_finish_usec": 1000 + 10 * rank, "last_finish_usec": 2000 + 100 * rank,
        "ops_total": {"entries": 1 + rank, "bytes": (1 << 30) * (rank + 1), "iops": 1024 * (rank + 1)},
        "ops_stonewall_total": {"entries": rank, "bytes": 1 << 29, "iops": 512},
        "ops_readmix_total": {"entries": 0, "bytes": 0, "iops": 0},
        "iops_lat_histo": histo(7), "entries_lat_histo": histo(3),
        "verify_mismatch_bytes": rank * 3, "verified_bytes": 1 << 30, "filled_bytes": 1 << 30,
        "num_kernel_launches": 10, "h2d_bytes": 5, "d2h_bytes": 6, "dev_kernel_usec": 77,
        "num_workers_done": 2, "num_workers_done_with_error": 0,
    }


def _worker(rank, world_size, port, tmpdir):
    sys.path.insert(0, REPO_ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        from elbencho_b200 import distributed as elbdist
        from elbencho_b200 import BenchPhase, WorkerConfig
        from tests import oracle_lib

        # rank layout (--rankoffset semantics)
        off, total = elbdist.rank_layout(world_size, rank, 2)
        assert (off, total) == (rank * 2, 4)

        # stats reduce: sum / min / max
        res = elbdist.reduce_phase_results(fake_results(rank))
        assert res["ops_total"] == {"entries": 3, "bytes": 3 << 30, "iops": 3072}
        assert res["first_finish_usec"] == 1000 and res["last_finish_usec"] == 2100
        assert res["verify_mismatch_bytes"] == 3
        assert res["iops_lat_histo"]["buckets"][10] == 7 and res["iops_lat_histo"]["buckets"][11] == 7
        assert res["iops_lat_histo"]["buckets"][111] == 1
        assert res["iops_lat_histo"]["num"] == 15
        assert res["iops_lat_histo"]["min_usec"] == 5 and res["iops_lat_histo"]["max_usec"] == 200
        assert res["ops_per_sec"]["bytes"] == elbdist.per_sec_from_usec(3 << 30, 2100)
        assert res["ops_stonewall_per_sec"]["bytes"] == elbdist.per_sec_from_usec(1 << 30, 1000)
        assert elbdist.reduce_max_float(1.5 + rank) == 2.5
        ctr = torch.tensor([rank + 1] * 8, dtype=torch.int64)
        assert elbdist.reduce_device_counters(ctr).tolist() == [3] * 8

        # sharding: 2 processes x 2 threads over 2 shared files = the same bytes on disk as one
        # process with 4 threads (checked with the CPU oracle worker; the GPU worker's twin of
        # this test is tests/test_worker_gpu.py::test_rank_offset_sharding_two_managers)
        paths = [os.path.join(tmpdir, "shared_%d" % i) for i in range(2)]
        cfg = WorkerConfig(paths=paths, num_threads=2, rank_offset=off, num_dataset_threads=total,
                           block_size=4096, file_size=10 * 4096 + 100, integrity_check_salt=9)
        rc, workers, _ = oracle_lib.run_oracle_phase(cfg, BenchPhase.CREATEFILES)
        assert rc == 0
        local_bytes = sum(w.liveOps.numBytesDone for w in workers)
        t = torch.tensor([local_bytes], dtype=torch.int64)
        dist.all_reduce(t)
        assert int(t.item()) == 2 * (10 * 4096 + 100)
        dist.barrier()
        if rank == 0:
            for i, path in enumerate(paths):
                with open(path, "rb") as f:
                    data = f.read()
                assert data == oracle_lib.fill_pattern(10 * 4096 + 100, 0, 9)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_world_size_two_gloo(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)


def test_reduce_is_identity_without_process_group():
    from elbencho_b200 import distributed as elbdist
    res = elbdist.reduce_phase_results(fake_results(1))
    assert res["ops_total"]["bytes"] == 2 << 30
    assert res["iops_lat_histo"]["min_usec"] == 6
    with pytest.raises(ValueError):
        elbdist.rank_layout(2, 2, 4)
