"""CPU tests of the oracle's per-block loops against each other and against closed forms: the aio
restatement (LocalWorker.cpp:1795-2037, kernel AIO syscalls) must write the same bytes and count
the same I/Os as the sync loop (:1669-1781); the rate limiter (toolkits/RateLimiter.h) and the
stonewall snapshot of both counter sets (Worker.h:203-209) behave as the reference describes."""
import hashlib
import os
import shutil
import tempfile
import time

import pytest

from elbencho_b200 import BenchPhase, PathType, WorkerConfig
from tests import oracle_lib

MiB = 1 << 20
KiB = 1 << 10


@pytest.fixture()
def workdir():
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    path = tempfile.mkdtemp(prefix="elb_orc_", dir=base)
    yield path
    shutil.rmtree(path, ignore_errors=True)


def sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


@pytest.mark.parametrize("threads,block,size", [(1, 64 * KiB, 4 * MiB), (3, 256 * KiB, 8 * MiB + 77),
                                                (2, 4 * KiB, 1 * MiB)])
def test_aio_loop_equals_sync_loop(workdir, threads, block, size):
    results = {}
    for depth in (1, 8):
        path = os.path.join(workdir, "f%d" % depth)
        cfg = WorkerConfig(paths=[path], num_threads=threads, block_size=block, file_size=size,
                           integrity_check_salt=21, io_depth=depth)
        out = []
        for phase in (BenchPhase.CREATEFILES, BenchPhase.READFILES):
            rc, workers, pres = oracle_lib.run_oracle_phase(cfg, phase)
            assert rc == 0, [w.errorMsg for w in workers if w.hadError]
            out.append((pres.opsTotal.numBytesDone, pres.opsTotal.numIOPSDone,
                        pres.iopsLatHisto.numStoredValues))
        results[depth] = (out, sha(path), os.path.getsize(path))
    assert results[1] == results[8]
    assert results[1][0][0][0] == size


def test_aio_loop_random_reads_verify_and_find_corruption(workdir):
    path = os.path.join(workdir, "r")
    base = dict(paths=[path], num_threads=2, block_size=4 * KiB, file_size=2 * MiB,
                integrity_check_salt=5)
    assert oracle_lib.run_oracle_phase(WorkerConfig(**base), BenchPhase.CREATEFILES)[0] == 0
    rand = dict(base, io_depth=16, use_random_offsets=True, rand_offset_seed=11)
    rc, _, pres = oracle_lib.run_oracle_phase(WorkerConfig(**rand), BenchPhase.READFILES)
    assert rc == 0 and pres.opsTotal.numBytesDone == 2 * MiB  # full coverage by default
    with open(path, "r+b") as f:
        f.seek(123457)
        f.write(b"\xff")
    rc, workers, _ = oracle_lib.run_oracle_phase(WorkerConfig(**rand), BenchPhase.READFILES)
    assert rc != 0
    msgs = [w.errorMsg.decode() for w in workers if w.hadError]
    assert any(m.startswith("Data verification failed. Offset: 123457;") for m in msgs), msgs


def test_rate_limiter_budget_and_aio_invalidation_rule(workdir):
    """6 x 1 MiB at 2 blocks per second: >= 2 s; in the aio loop every I/O that was pending while
    the limiter slept stays out of the histogram (LocalWorker.cpp:1843-1845, 1966)"""
    path = os.path.join(workdir, "l")
    base = dict(paths=[path], num_threads=1, block_size=MiB, file_size=8 * MiB,
                integrity_check_salt=2)
    assert oracle_lib.run_oracle_phase(WorkerConfig(**base), BenchPhase.CREATEFILES)[0] == 0
    limited = dict(base, use_random_offsets=True, random_amount=6 * MiB, rand_offset_seed=3,
                   limit_read_bps=2 * MiB)
    t0 = time.time()
    rc, _, sync_res = oracle_lib.run_oracle_phase(WorkerConfig(**limited), BenchPhase.READFILES)
    assert rc == 0 and time.time() - t0 >= 2.0
    assert sync_res.opsTotal.numIOPSDone == 6
    assert sync_res.iopsLatHisto.numStoredValues == 6  # sync loop: limiter runs before the stamp
    assert sync_res.iopsLatHisto.maxMicroSecLat < 500000
    rc, _, aio_res = oracle_lib.run_oracle_phase(WorkerConfig(io_depth=4, **limited),
                                                 BenchPhase.READFILES)
    assert rc == 0 and aio_res.opsTotal.numIOPSDone == 6
    assert aio_res.iopsLatHisto.numStoredValues < 6


def test_stonewall_snapshots_both_counter_sets(workdir):
    """the deterministic straggler of tests/test_worker_variants_gpu.py on the oracle alone"""
    path = os.path.join(workdir, "s")
    size = 128 * MiB
    base = dict(paths=[path], num_threads=2, block_size=MiB, file_size=size, integrity_check_salt=3)
    assert oracle_lib.run_oracle_phase(WorkerConfig(**base), BenchPhase.CREATEFILES)[0] == 0
    rc, _, pres = oracle_lib.run_oracle_phase(
        WorkerConfig(num_rwmix_read_threads=1, limit_read_bps=32 * MiB, **base),
        BenchPhase.CREATEFILES)
    assert rc == 0
    assert pres.opsStoneWallTotal.numBytesDone == size // 2          # the writer, when it finished
    assert pres.opsStoneWallReadMixTotal.numBytesDone == 32 * MiB     # the reader's first budget
    assert pres.opsStoneWallReadMixTotal.numIOPSDone == 32
    assert pres.opsReadMixTotal.numBytesDone == size // 2
    assert pres.firstFinishUSec < 1000000 < pres.lastFinishUSec


def test_dir_mode_with_aio_loop(workdir):
    cfg = dict(paths=[workdir], path_type=int(PathType.DIR), num_threads=2, num_dirs=2, num_files=3,
               block_size=16 * KiB, file_size=80 * KiB, integrity_check_salt=8)
    totals = {}
    for depth in (1, 4):
        sub = os.path.join(workdir, "d%d" % depth)
        os.makedirs(sub)
        c = WorkerConfig(**dict(cfg, paths=[sub], io_depth=depth))
        for phase in (BenchPhase.CREATEDIRS, BenchPhase.CREATEFILES, BenchPhase.READFILES):
            rc, workers, pres = oracle_lib.run_oracle_phase(c, phase)
            assert rc == 0, [w.errorMsg for w in workers if w.hadError]
        totals[depth] = (pres.opsTotal.numBytesDone, pres.opsTotal.numIOPSDone,
                         pres.opsTotal.numEntriesDone)
    assert totals[1] == totals[4] == (2 * 2 * 3 * 80 * KiB, 2 * 2 * 3 * 5, 12)
