"""CPU tests of bench.py's step plans: the slices of all (rank, step) pairs tile every file exactly
once (the reference's --rankoffset sharding, LocalWorker.cpp:3576-3589), for one and two GPUs'
worth of files; run through the oracle (the reference arm's path)."""
import argparse
import os
import shutil
import tempfile

import pytest

import bench
from elbencho_b200 import BenchPhase, WorkerConfig
from tests import oracle_lib

MiB = 1 << 20


def make_args(workdir, **over):
    args = argparse.Namespace(gpus=1, steps=3, warmup=1, impl="reference", config="c2",
                              file_gib=24 / 1024, threads=2, dir=workdir, salt=5, direct=False)
    for key, val in over.items():
        setattr(args, key, val)
    return args


@pytest.fixture()
def workdir():
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    path = tempfile.mkdtemp(prefix="elb_benchplan_", dir=base)
    yield path
    shutil.rmtree(path, ignore_errors=True)


@pytest.mark.parametrize("world", [1, 2])
def test_c2_slices_tile_every_file_exactly_once(workdir, world):
    args = make_args(workdir, gpus=world)
    workload = bench.C2(args, world)
    assert workload.file_size % (workload.num_slices * workload.threads * MiB) == 0
    covered = [bytearray(workload.file_size // MiB) for _ in range(world)]
    block = workload.block
    for step in range(workload.num_slices):
        for rank in range(world):
            cfg, phases = workload.plan(step, rank)
            assert [p.name for p in phases] == ["CREATEFILES", "READFILES"]
            threads, total = cfg["num_threads"], cfg["num_dataset_threads"]
            blocks_total = world * workload.file_size // block
            per_rank = blocks_total // total
            for t in range(threads):
                g = cfg["rank_offset"] + t
                for blk in range(g * per_rank, (g + 1) * per_rank):
                    file_idx, in_file = divmod(blk, workload.file_size // block)
                    assert file_idx == rank  # a process only touches its own file
                    covered[file_idx][in_file] += 1
    assert all(set(c) == {1} for c in covered)

    # and through the oracle: all steps written -> every file verifies as a whole
    totals = bench.cpu_arm(workload, args.steps, args.warmup)
    assert totals["phase"]["CREATEFILES"]["bytes"] == args.steps * workload.slice_bytes * world
    assert totals["phase"]["READFILES"]["bytes"] == args.steps * workload.slice_bytes * world
    for path in workload.paths:
        assert os.path.getsize(path) == workload.file_size
        whole = WorkerConfig(paths=[path], num_threads=1, block_size=MiB,
                             file_size=workload.file_size, integrity_check_salt=args.salt)
        assert oracle_lib.run_oracle_phase(whole, BenchPhase.READFILES)[0] == 0


def test_config_objects_of_both_arms_are_equal(workdir):
    """what the driver compares for `same_config`"""
    for name in ("c2", "c3", "c4", "c5"):
        args = make_args(workdir, config=name, file_gib=32 / 1024)
        workload = bench.WORKLOADS[name](args, 1)
        assert bench.common_config(args, workload, 1) == bench.common_config(args, workload, 1)
        assert bench.common_config(args, workload, 1)["baseline_config"] == name


def test_c3_issues_the_file_worth_of_ios_over_the_timed_steps(workdir):
    args = make_args(workdir, config="c3", file_gib=16 / 1024, steps=4, warmup=1)
    workload = bench.C3(args, 1)
    assert workload.ios_per_step * args.steps * workload.block == workload.file_size
    totals = bench.cpu_arm(workload, args.steps, args.warmup)
    assert totals["iops"] == workload.ios_per_step * args.steps
    assert workload.value_of(totals) > 0


def test_c5_counts_files_and_reads_them_back(workdir):
    args = make_args(workdir, config="c5", file_gib=64 / 1024, steps=2, warmup=1)
    workload = bench.C5(args, 1)
    totals = bench.cpu_arm(workload, args.steps, args.warmup)
    files_per_step = workload.files_per_dir * workload.DIRS_PER_STEP * workload.threads
    assert totals["phase"]["CREATEFILES"]["entries"] == files_per_step * args.steps
    assert totals["phase"]["READFILES"]["entries"] == files_per_step * args.steps
    assert totals["phase"]["READFILES"]["bytes"] == files_per_step * args.steps * workload.block
