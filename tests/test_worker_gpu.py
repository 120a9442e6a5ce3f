"""GPU parity tests of the worker pipeline (through the C ABI) against the CPU oracle worker:
identical bytes on disk for --verify writes, identical byte/IOPS/entry counters, identical verify
outcome and exception text, on the same inputs (SURVEY.md §8a "counter identities")."""
import hashlib
import os
import shutil
import tempfile

import pytest

from elbencho_b200 import BenchPhase, PathType, WorkerConfig, WorkerError, WorkerManager
from elbencho_b200.worker import IOEngine
from tests import oracle_lib

pytestmark = pytest.mark.gpu

MiB = 1 << 20
KiB = 1 << 10


@pytest.fixture(autouse=True, params=["kernel", "copyengine"])
def staging_engine(request, monkeypatch):
    """every test of this module runs with both staging engines: the fill / verify kernels moving
    the blocks themselves, and cudaMemcpyAsync (+ CUDA graphs) around the kernels"""
    monkeypatch.setenv("ELB_STAGING", request.param)
    return request.param


@pytest.fixture()
def workdir(cuda_device):
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    path = tempfile.mkdtemp(prefix="elb_test_", dir=base)
    yield path
    shutil.rmtree(path, ignore_errors=True)


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 22), b""):
            h.update(chunk)
    return h.hexdigest()


def gpu_and_cpu_configs(workdir, names, **kwargs):
    """same config twice: GPU worker files and oracle files"""
    gpu_paths = [os.path.join(workdir, "gpu_" + n) for n in names]
    cpu_paths = [os.path.join(workdir, "cpu_" + n) for n in names]
    return WorkerConfig(paths=gpu_paths, **kwargs), WorkerConfig(paths=cpu_paths, **kwargs)


def check_counters(gpu_res, orc_phase_res, orc_worker_res=None, mgr=None):
    assert gpu_res["ops_total"]["bytes"] == orc_phase_res.opsTotal.numBytesDone
    assert gpu_res["ops_total"]["iops"] == orc_phase_res.opsTotal.numIOPSDone
    assert gpu_res["ops_total"]["entries"] == orc_phase_res.opsTotal.numEntriesDone
    assert gpu_res["iops_lat_histo"]["num"] == orc_phase_res.iopsLatHisto.numStoredValues
    assert gpu_res["entries_lat_histo"]["num"] == orc_phase_res.entriesLatHisto.numStoredValues
    assert gpu_res["num_workers_done_with_error"] == 0
    if orc_worker_res is not None and mgr is not None:
        for i, worker in enumerate(mgr.workers()):
            ops, _ = worker.live_ops()
            assert ops["bytes"] == orc_worker_res[i].liveOps.numBytesDone, i
            assert ops["iops"] == orc_worker_res[i].liveOps.numIOPSDone, i
            assert ops["entries"] == orc_worker_res[i].liveOps.numEntriesDone, i
            assert worker.got_work == bool(orc_worker_res[i].gotPhaseWork), i


@pytest.mark.parametrize("threads,nfiles,block,size", [
    (1, 1, MiB, 8 * MiB),            # BASELINE config 1/2 shape, scaled down
    (1, 1, MiB, 5 * MiB + 1000),     # partial last block
    (3, 2, 64 * KiB, 1 * MiB + 77),  # multi-file, ragged partition, last rank takes remainder
    (4, 1, 4 * KiB, 10 * KiB),       # more threads than... 3 blocks: rank 3 takes the remainder
    (5, 1, 8 * KiB, 16 * KiB),       # some workers get no work
    (2, 3, 1000, 10001),             # odd block size
])
def test_file_mode_seq_write_read_verify(workdir, threads, nfiles, block, size):
    names = ["f%d" % i for i in range(nfiles)]
    gcfg, ccfg = gpu_and_cpu_configs(workdir, names, num_threads=threads, block_size=block,
                                     file_size=size, integrity_check_salt=1,
                                     pipeline_batch_blocks=3)
    with WorkerManager(gcfg) as mgr:
        gw = mgr.run_phase(BenchPhase.CREATEFILES)
        rc, ow, opr = oracle_lib.run_oracle_phase(ccfg, BenchPhase.CREATEFILES)
        assert rc == 0
        check_counters(gw, opr, ow, mgr)
        assert gw["filled_bytes"] == gw["ops_total"]["bytes"]
        for g, c in zip(gcfg.paths, ccfg.paths):
            assert os.path.getsize(g) == os.path.getsize(c) == size
            assert sha(g) == sha(c)
        assert mgr.expected_totals(BenchPhase.CREATEFILES)[1] // 1 >= 0

        gr = mgr.run_phase(BenchPhase.READFILES)
        rc, ow, opr = oracle_lib.run_oracle_phase(ccfg, BenchPhase.READFILES)
        assert rc == 0
        check_counters(gr, opr, ow, mgr)
        assert gr["verify_mismatch_bytes"] == 0
        assert gr["verified_bytes"] == gr["ops_total"]["bytes"]
        assert gr["h2d_bytes"] == gr["ops_total"]["bytes"]

    # cross check: the oracle's reader accepts the GPU worker's file and vice versa
    cross = WorkerConfig(paths=gcfg.paths, num_threads=1, block_size=128 * KiB if block >= MiB
                         else block, file_size=size, integrity_check_salt=1)
    rc, _, _ = oracle_lib.run_oracle_phase(cross, BenchPhase.READFILES)
    assert rc == 0
    cross_gpu = WorkerConfig(paths=ccfg.paths, num_threads=2, block_size=128 * KiB if block >= MiB
                             else block, file_size=size, integrity_check_salt=1)
    with WorkerManager(cross_gpu) as mgr:
        res = mgr.run_phase(BenchPhase.READFILES)
        assert res["verify_mismatch_bytes"] == 0


def test_verify_failure_message_matches_oracle(workdir):
    size, block = 4 * MiB, MiB
    gcfg, ccfg = gpu_and_cpu_configs(workdir, ["f"], num_threads=1, block_size=block,
                                     file_size=size, integrity_check_salt=7)
    with WorkerManager(gcfg) as mgr:
        mgr.run_phase(BenchPhase.CREATEFILES)
        oracle_lib.run_oracle_phase(ccfg, BenchPhase.CREATEFILES)
        for path in (gcfg.paths[0], ccfg.paths[0]):
            with open(path, "r+b") as f:
                for pos in (2 * MiB + 12345, 3 * MiB + 5):
                    f.seek(pos)
                    byte = f.read(1)
                    f.seek(pos)
                    f.write(bytes([byte[0] ^ 0x21]))
        rc, ow, _ = oracle_lib.run_oracle_phase(ccfg, BenchPhase.READFILES)
        assert rc != 0 and ow[0].hadError
        oracle_msg = ow[0].errorMsg.decode()
        assert oracle_msg.startswith("Data verification failed. Offset: %d;" % (2 * MiB + 12345))
        with pytest.raises(WorkerError) as excinfo:
            mgr.run_phase(BenchPhase.READFILES)
        assert str(excinfo.value) == oracle_msg
        assert mgr.worker(0).last_error == oracle_msg
        assert mgr.phase_results()["num_workers_done_with_error"] == 1
        # the manager stays usable: rewrite and verify again
        mgr.run_phase(BenchPhase.CREATEFILES)
        assert mgr.run_phase(BenchPhase.READFILES)["verify_mismatch_bytes"] == 0

    # collect-all mode counts every bad byte on the device instead of stopping
    with open(gcfg.paths[0], "r+b") as f:
        for pos in (100, 2 * MiB + 1, 2 * MiB + 2):
            f.seek(pos)
            f.write(b"\xEE")
    count_cfg = WorkerConfig(paths=gcfg.paths, num_threads=2, block_size=block, file_size=size,
                             integrity_check_salt=7, verify_collect_all=True)
    with WorkerManager(count_cfg) as mgr:
        res = mgr.run_phase(BenchPhase.READFILES)
        with open(gcfg.paths[0], "rb") as f:
            data = f.read()
        assert res["verify_mismatch_bytes"] == oracle_lib.verify_pattern(data, 0, 7)[1] == 3


def test_file_mode_random_full_coverage_write_then_random_read(workdir):
    size, block, threads = 2 * MiB, 4 * KiB, 2
    common = dict(num_threads=threads, block_size=block, file_size=size, integrity_check_salt=3,
                  use_random_offsets=True, rand_offset_seed=1234)
    gcfg, ccfg = gpu_and_cpu_configs(workdir, ["a", "b"], **common)
    with WorkerManager(gcfg) as mgr:
        gw = mgr.run_phase(BenchPhase.CREATEFILES)
        rc, ow, opr = oracle_lib.run_oracle_phase(ccfg, BenchPhase.CREATEFILES)
        assert rc == 0
        check_counters(gw, opr, ow, mgr)
        # full coverage: every block written exactly once -> files complete and identical
        for g, c in zip(gcfg.paths, ccfg.paths):
            assert os.path.getsize(g) == size
            assert sha(g) == sha(c)
        gr = mgr.run_phase(BenchPhase.READFILES)
        rc, ow, opr = oracle_lib.run_oracle_phase(ccfg, BenchPhase.READFILES)
        assert rc == 0
        check_counters(gr, opr, ow, mgr)
        assert gr["ops_total"]["bytes"] == 2 * size  # randamount default = fileSize * numFiles


@pytest.mark.parametrize("extra", [
    dict(do_reverse_seq_offsets=True),
    dict(use_strided_access=True),
    dict(use_random_offsets=True, use_random_unaligned=True, rand_offset_seed=9,
         random_amount=3 * MiB),
    dict(use_random_offsets=True, use_explicit_rand_offset_algo=True, rand_offset_seed=10),
    # --randalgo fast / balanced / strong (RandAlgoSelectorTk.h:10-13)
    dict(use_random_offsets=True, use_explicit_rand_offset_algo=True, rand_offset_seed=11,
         rand_offset_algo=1),
    dict(use_random_offsets=True, use_explicit_rand_offset_algo=True, rand_offset_seed=12,
         rand_offset_algo=2),
    dict(use_random_offsets=True, use_random_unaligned=True, rand_offset_seed=13,
         random_amount=2 * MiB, rand_offset_algo=3),
])
def test_file_mode_offset_variants(workdir, extra):
    size, block, threads = 1 * MiB + 4096 * 3, 4 * KiB * 3, 2
    gcfg, ccfg = gpu_and_cpu_configs(workdir, ["f"], num_threads=threads, block_size=block,
                                     file_size=size, integrity_check_salt=5, **extra)
    # complete file first so that reads of any variant have data
    full = WorkerConfig(paths=[gcfg.paths[0]], block_size=MiB, file_size=size,
                        integrity_check_salt=5)
    with WorkerManager(full) as mgr:
        mgr.run_phase(BenchPhase.CREATEFILES)
    shutil.copy(gcfg.paths[0], ccfg.paths[0])
    with WorkerManager(gcfg) as mgr:
        gw = mgr.run_phase(BenchPhase.CREATEFILES)
        rc, ow, opr = oracle_lib.run_oracle_phase(ccfg, BenchPhase.CREATEFILES)
        assert rc == 0
        check_counters(gw, opr, ow, mgr)
        assert sha(gcfg.paths[0]) == sha(ccfg.paths[0])
        gr = mgr.run_phase(BenchPhase.READFILES)
        rc, ow, opr = oracle_lib.run_oracle_phase(ccfg, BenchPhase.READFILES)
        assert rc == 0
        check_counters(gr, opr, ow, mgr)


@pytest.mark.parametrize("block,fsize,batch", [(64 * KiB, 64 * KiB, 0), (16 * KiB, 50 * KiB, 5),
                                               (64 * KiB, 0, 0)])
def test_dir_mode_full_cycle(workdir, block, fsize, batch):
    """BASELINE config 5 shape, scaled down: -d -w -r -n 2 -N 3 --verify"""
    gdir = os.path.join(workdir, "gpu")
    cdir = os.path.join(workdir, "cpu")
    os.mkdir(gdir)
    os.mkdir(cdir)
    common = dict(path_type=PathType.DIR, num_threads=3, num_dirs=2, num_files=3,
                  block_size=block, file_size=fsize, integrity_check_salt=1,
                  pipeline_batch_blocks=batch)
    gcfg = WorkerConfig(paths=[gdir], **common)
    ccfg = WorkerConfig(paths=[cdir], **common)
    with WorkerManager(gcfg) as mgr:
        for phase in (BenchPhase.CREATEDIRS, BenchPhase.CREATEFILES, BenchPhase.STATFILES,
                      BenchPhase.READFILES):
            gres = mgr.run_phase(phase)
            rc, ow, opr = oracle_lib.run_oracle_phase(ccfg, phase)
            assert rc == 0, phase
            check_counters(gres, opr, ow, mgr)
            exp_entries, exp_bytes = mgr.expected_totals(phase)
            assert gres["ops_total"]["entries"] == exp_entries
            assert gres["ops_total"]["bytes"] == exp_bytes
        # same namespace, same bytes (LocalWorker.cpp:3064-3068)
        gfiles = sorted(os.path.relpath(os.path.join(r, f), gdir)
                        for r, _, fs in os.walk(gdir) for f in fs)
        cfiles = sorted(os.path.relpath(os.path.join(r, f), cdir)
                        for r, _, fs in os.walk(cdir) for f in fs)
        assert gfiles == cfiles and len(gfiles) == 3 * 2 * 3
        assert "r1/d0/r1-f2" in gfiles
        for rel in gfiles:
            assert sha(os.path.join(gdir, rel)) == sha(os.path.join(cdir, rel))
        for phase in (BenchPhase.DELETEFILES, BenchPhase.DELETEDIRS):
            gres = mgr.run_phase(phase)
            rc, ow, opr = oracle_lib.run_oracle_phase(ccfg, phase)
            assert rc == 0
            check_counters(gres, opr, ow, mgr)
        assert os.listdir(gdir) == []


@pytest.mark.parametrize("engine,depth,direct", [(IOEngine.AIO, 8, True), (IOEngine.AIO, 4, False),
                                                  (IOEngine.SYNC, 1, True)])
def test_aio_and_direct_io(workdir, engine, depth, direct):
    """BASELINE config 3 shape, scaled down: 4 KiB random reads at iodepth > 1"""
    size, block = 4 * MiB, 4 * KiB
    gcfg, ccfg = gpu_and_cpu_configs(workdir, ["f"], num_threads=2, block_size=block,
                                     file_size=size, integrity_check_salt=11,
                                     use_direct_io=direct)
    seq = WorkerConfig(paths=gcfg.paths, num_threads=2, block_size=256 * KiB, file_size=size,
                       integrity_check_salt=11, use_direct_io=direct, io_depth=depth,
                       io_engine=engine)
    with WorkerManager(seq) as mgr:
        gw = mgr.run_phase(BenchPhase.CREATEFILES)
        assert gw["ops_total"]["bytes"] == size
    rc, _, _ = oracle_lib.run_oracle_phase(
        WorkerConfig(paths=ccfg.paths, num_threads=2, block_size=256 * KiB, file_size=size,
                     integrity_check_salt=11), BenchPhase.CREATEFILES)
    assert rc == 0
    assert sha(gcfg.paths[0]) == sha(ccfg.paths[0])
    rnd = dict(num_threads=2, block_size=block, file_size=size, integrity_check_salt=11,
               use_random_offsets=True, rand_offset_seed=77)
    with WorkerManager(WorkerConfig(paths=gcfg.paths, use_direct_io=direct, io_depth=depth,
                                    io_engine=engine, **rnd)) as mgr:
        gr = mgr.run_phase(BenchPhase.READFILES)
        rc, ow, opr = oracle_lib.run_oracle_phase(WorkerConfig(paths=ccfg.paths, **rnd),
                                                  BenchPhase.READFILES)
        assert rc == 0
        check_counters(gr, opr, ow, mgr)
        assert gr["ops_total"]["iops"] == size // block
        assert gr["verify_mismatch_bytes"] == 0


def test_block_variance_fill_matches_cpu_twin(workdir):
    """--blockvarpct on the GPU: bytes on disk equal the counter-based CPU twin per block
    (block counter = (rank << 40) + numIOPSSubmitted of that worker)."""
    size, block, threads, pct, seed = 3 * MiB, 256 * KiB, 2, 60, 4242
    cfg = WorkerConfig(paths=[os.path.join(workdir, "rnd")], num_threads=threads, block_size=block,
                       file_size=size, block_variance_percent=pct, block_variance_seed=seed,
                       pipeline_batch_blocks=4)
    with WorkerManager(cfg) as mgr:
        res = mgr.run_phase(BenchPhase.CREATEFILES)
        assert res["ops_total"]["bytes"] == size
        assert res["filled_bytes"] == size
        with open(cfg.paths[0], "rb") as f:
            data = f.read()
        blocks_per_rank = (size // block) // threads
        for blk in range(size // block):
            rank = blk // blocks_per_rank
            ctr = (rank << 40) + (blk % blocks_per_rank)
            assert data[blk * block:(blk + 1) * block] == \
                oracle_lib.fill_random_ctr(block, pct, seed, ctr), blk
        # second write phase: counters keep running (numIOPSSubmitted is never reset,
        # LocalWorker.h:121) so the content changes
        mgr.run_phase(BenchPhase.CREATEFILES)
        with open(cfg.paths[0], "rb") as f:
            data2 = f.read()
        assert data2[:block] == oracle_lib.fill_random_ctr(block, pct, seed, blocks_per_rank)
        assert data2 != data


def test_plain_write_read_without_verify(workdir, staging_engine):
    size, block = 2 * MiB, 512 * KiB
    cfg = WorkerConfig(paths=[os.path.join(workdir, "plain")], num_threads=2, block_size=block,
                       file_size=size)
    with WorkerManager(cfg) as mgr:
        w = mgr.run_phase(BenchPhase.CREATEFILES)
        assert w["ops_total"] == {"entries": 0, "bytes": size, "iops": size // block}
        # nothing to fill or verify: the kernel staging engine moves the blocks with its
        # stage-copy kernel (and the ring content written equals the device ring's), the
        # copy-engine one launches nothing
        assert (w["num_kernel_launches"] > 0) == (staging_engine == "kernel")
        assert w["d2h_bytes"] == size and w["filled_bytes"] == 0
        r = mgr.run_phase(BenchPhase.READFILES)
        assert (r["num_kernel_launches"] > 0) == (staging_engine == "kernel")
        assert r["ops_total"] == {"entries": 0, "bytes": size, "iops": size // block}
        assert r["h2d_bytes"] == size
        assert r["first_finish_usec"] > 0 and r["last_finish_usec"] >= r["first_finish_usec"]
        assert r["ops_per_sec"]["bytes"] == mgr._lib.elb_per_sec_from_usec(
            size, r["last_finish_usec"])
        mgr.run_phase(BenchPhase.SYNC)
        d = mgr.run_phase(BenchPhase.DELETEFILES)
        assert d["ops_total"]["entries"] == 2  # every worker tries every file
        assert not os.path.exists(cfg.paths[0])


def test_short_read_error_text(workdir):
    size, block = 1 * MiB, 256 * KiB
    cfg = WorkerConfig(paths=[os.path.join(workdir, "short")], block_size=block, file_size=size,
                       integrity_check_salt=1)
    with WorkerManager(cfg) as mgr:
        mgr.run_phase(BenchPhase.CREATEFILES)
        os.truncate(cfg.paths[0], size - 1000)
        with pytest.raises(WorkerError, match="Unexpected short file read. Path: .*short; "
                                              "Bytes read: %d; Expected read: %d" %
                                              (block - 1000, block)):
            mgr.run_phase(BenchPhase.READFILES)


def test_live_stats_and_interrupt(workdir):
    size, block = 256 * MiB, MiB
    cfg = WorkerConfig(paths=[os.path.join(workdir, "big")], num_threads=2, block_size=block,
                       file_size=size, integrity_check_salt=1)
    with WorkerManager(cfg) as mgr:
        mgr.start_phase(BenchPhase.CREATEFILES)
        seen = 0
        while not mgr.wait_done(5):
            ops, _ = mgr.live_ops()
            assert ops["bytes"] >= seen
            seen = ops["bytes"]
        res = mgr.phase_results()
        assert res["ops_total"]["bytes"] == size
        assert res["ops_stonewall_total"]["bytes"] <= size
        assert res["ops_stonewall_total"]["bytes"] > 0
        lat = mgr.live_latency()
        assert lat["numAvgIOLatValues"] == size // block
        # friendly interruption ends the phase without error (LocalWorker.cpp:372-387)
        mgr.start_phase(BenchPhase.READFILES)
        mgr.interrupt()
        assert mgr.wait_done(-1)
        assert mgr.phase_results()["num_workers_done_with_error"] == 0
        # and the workers are ready for the next phase
        assert mgr.run_phase(BenchPhase.READFILES)["ops_total"]["bytes"] == size


def test_multi_gpu_round_robin_assignment(workdir):
    import torch
    ngpus = torch.cuda.device_count()
    ids = list(range(ngpus))
    cfg = WorkerConfig(paths=[os.path.join(workdir, "mg")], num_threads=max(2, ngpus) + 1,
                       block_size=64 * KiB, file_size=4 * MiB, integrity_check_salt=1, gpu_ids=ids)
    with WorkerManager(cfg) as mgr:
        # rank -> GPU = gpuIDs[rank % size] (LocalWorker.cpp:1420-1422)
        assert [w.gpu_id for w in mgr.workers()] == [ids[w.rank % ngpus] for w in mgr.workers()]
        mgr.run_phase(BenchPhase.CREATEFILES)
        res = mgr.run_phase(BenchPhase.READFILES)
        assert res["verified_bytes"] == 4 * MiB and res["verify_mismatch_bytes"] == 0
        for w in mgr.workers():
            assert w.dev_counters_ptr != 0


def test_rank_offset_sharding_two_managers(workdir):
    """two 'processes' (managers with --rankoffset semantics) x 2 threads over 2 shared files
    produce the same bytes as one manager with 4 threads; this is how bench.py shards ranks under
    torchrun (no data-path collective)."""
    from elbencho_b200 import distributed as elbdist
    size, block = 10 * 4096 + 100, 4096
    shared_paths = [os.path.join(workdir, "shared_%d" % i) for i in range(2)]
    single_paths = [os.path.join(workdir, "single_%d" % i) for i in range(2)]
    total_bytes = 0
    mgrs = []
    for rank in range(2):
        off, total = elbdist.rank_layout(2, rank, 2)
        mgrs.append(WorkerManager(WorkerConfig(paths=shared_paths, num_threads=2, rank_offset=off,
                                               num_dataset_threads=total, block_size=block,
                                               file_size=size, integrity_check_salt=9)))
    try:
        for mgr in mgrs:
            mgr.start_phase(BenchPhase.CREATEFILES)
        for mgr in mgrs:
            assert mgr.wait_done(-1)
            total_bytes += mgr.phase_results()["ops_total"]["bytes"]
        assert [w.rank for m in mgrs for w in m.workers()] == [0, 1, 2, 3]
        res = [m.run_phase(BenchPhase.READFILES) for m in mgrs]
        assert sum(r["verified_bytes"] for r in res) == 2 * size
    finally:
        for mgr in mgrs:
            mgr.close()
    assert total_bytes == 2 * size
    with WorkerManager(WorkerConfig(paths=single_paths, num_threads=4, block_size=block,
                                    file_size=size, integrity_check_salt=9)) as mgr:
        mgr.run_phase(BenchPhase.CREATEFILES)
    for a, b in zip(shared_paths, single_paths):
        assert sha(a) == sha(b)
        with open(a, "rb") as f:
            assert f.read() == oracle_lib.fill_pattern(size, 0, 9)


def test_mid_size_multi_batch_run_bytes_equal_oracle(workdir, staging_engine):
    """4 GiB through many batches per worker (kernel staging: the default cache-resident batches;
    copy-engine staging: 16 MiB batches that are replayed from per-batch CUDA graphs): file
    bytes, counters and verify outcome equal the oracle's."""
    batch_blocks, num_batches = (0, 0) if staging_engine == "kernel" else (16, 2)
    size, block, threads = 4 << 30, MiB, 4
    kwargs = dict(num_threads=threads, block_size=block, file_size=size, integrity_check_salt=11)
    gcfg, ccfg = gpu_and_cpu_configs(workdir, ["big"], **kwargs)
    tuned = WorkerConfig(paths=gcfg.paths, pipeline_batch_blocks=batch_blocks,
                         pipeline_num_batches=num_batches, **kwargs)
    with WorkerManager(tuned) as mgr:
        gw = mgr.run_phase(BenchPhase.CREATEFILES)
        gr = mgr.run_phase(BenchPhase.READFILES)
    rc, _, opw = oracle_lib.run_oracle_phase(ccfg, BenchPhase.CREATEFILES)
    assert rc == 0
    assert gw["ops_total"]["bytes"] == opw.opsTotal.numBytesDone == size
    assert gw["ops_total"]["iops"] == opw.opsTotal.numIOPSDone == size // block
    assert gw["filled_bytes"] == size and gw["d2h_bytes"] == size
    assert gr["verified_bytes"] == size and gr["verify_mismatch_bytes"] == 0
    assert gr["h2d_bytes"] == size
    assert gr["dev_kernel_usec"] > 0 and gw["dev_kernel_usec"] > 0
    assert sha(gcfg.paths[0]) == sha(ccfg.paths[0])
