"""GDS code path of the worker (--cufile / --gds, incl. iodepth > 1 through the cuFile batch API),
run on a real GPU against tests/mock_cufile (a POSIX + cudaMemcpy stand-in for libcufile: the real
library cannot register file handles on the graft GPU boxes, see tests/mock_cufile/mock_cufile.cpp)
and checked bit-exactly against the CPU oracle."""
import ctypes
import hashlib
import os
import shutil
import subprocess
import sys
import tempfile

import pytest

MOCK_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mock_cufile")
MOCK_LIB = os.path.join(MOCK_DIR, "libmock_cufile.so")

# must be set before the native library binds cuFile for the first time in this process
os.environ.setdefault("ELB_CUFILE_LIB", MOCK_LIB)

from elbencho_b200 import BenchPhase, PathType, WorkerConfig, WorkerError, WorkerManager  # noqa: E402
from elbencho_b200.worker import IOEngine  # noqa: E402
from tests import oracle_lib  # noqa: E402

pytestmark = pytest.mark.gpu

MiB = 1 << 20
KiB = 1 << 10
ST = ["driver_open", "handle_reg", "handle_dereg", "buf_reg", "buf_dereg", "read", "write",
      "batch_submit", "batch_ops"]


@pytest.fixture(scope="module")
def mock():
    if not os.path.exists(MOCK_LIB):
        subprocess.run(["make", "-s", "-C", MOCK_DIR], check=True)
    lib = ctypes.CDLL(MOCK_LIB)  # same handle the worker dlopen()s
    lib.mock_cufile_get_stats.argtypes = [ctypes.POINTER(ctypes.c_uint64)]

    class Mock:
        def stats(self):
            out = (ctypes.c_uint64 * len(ST))()
            lib.mock_cufile_get_stats(out)
            return dict(zip(ST, out))

        def reset(self):
            lib.mock_cufile_reset_stats()
    return Mock()


@pytest.fixture()
def workdir(cuda_device):
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    path = tempfile.mkdtemp(prefix="elb_cufile_", dir=base)
    yield path
    shutil.rmtree(path, ignore_errors=True)


def sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def test_cufile_sync_write_read_verify(workdir, mock):
    size, block, threads = 6 * MiB, 512 * KiB, 2
    gpath, cpath = os.path.join(workdir, "g"), os.path.join(workdir, "c")
    cfg = WorkerConfig(paths=[gpath], num_threads=threads, block_size=block, file_size=size,
                       integrity_check_salt=3, use_cufile=True, use_gds_buf_reg=True,
                       pipeline_batch_blocks=3)
    mock.reset()
    with WorkerManager(cfg) as mgr:
        st = mock.stats()
        assert st["driver_open"] <= 1 and st["handle_reg"] == 1 and st["buf_reg"] == threads
        w = mgr.run_phase(BenchPhase.CREATEFILES)
        assert w["ops_total"] == {"entries": 0, "bytes": size, "iops": size // block}
        assert w["d2h_bytes"] == 0 and w["filled_bytes"] == size  # no host staging at all
        assert mock.stats()["write"] == size // block
        rc, ow, opr = oracle_lib.run_oracle_phase(
            WorkerConfig(paths=[cpath], num_threads=threads, block_size=block, file_size=size,
                         integrity_check_salt=3), BenchPhase.CREATEFILES)
        assert rc == 0 and sha(gpath) == sha(cpath)
        r = mgr.run_phase(BenchPhase.READFILES)
        assert r["ops_total"]["bytes"] == size and r["h2d_bytes"] == 0
        assert r["verified_bytes"] == size and r["verify_mismatch_bytes"] == 0
        assert mock.stats()["read"] == size // block
        # corrupt one byte: the message needs the actual byte, which only exists on the device
        with open(gpath, "r+b") as f:
            f.seek(4 * MiB + 77)
            f.write(b"\x99")
        with open(cpath, "r+b") as f:
            f.seek(4 * MiB + 77)
            f.write(b"\x99")
        rc, ow, _ = oracle_lib.run_oracle_phase(
            WorkerConfig(paths=[cpath], num_threads=threads, block_size=block, file_size=size,
                         integrity_check_salt=3), BenchPhase.READFILES)
        oracle_msg = [w_.errorMsg.decode() for w_ in ow if w_.hadError][0]
        with pytest.raises(WorkerError) as excinfo:
            mgr.run_phase(BenchPhase.READFILES)
        assert str(excinfo.value) == oracle_msg
    st = mock.stats()
    assert st["buf_dereg"] == threads and st["handle_dereg"] == 1


def test_cufile_batch_iodepth_random_reads(workdir, mock):
    """BASELINE config 3 shape (scaled down): 4 KiB random reads, iodepth 16, --gds cuFile batch."""
    size, block, threads, depth = 2 * MiB, 4 * KiB, 2, 16
    gpath, cpath = os.path.join(workdir, "g"), os.path.join(workdir, "c")
    with WorkerManager(WorkerConfig(paths=[gpath], block_size=MiB, file_size=size,
                                    integrity_check_salt=5)) as mgr:
        mgr.run_phase(BenchPhase.CREATEFILES)
    shutil.copy(gpath, cpath)
    rnd = dict(num_threads=threads, block_size=block, file_size=size, integrity_check_salt=5,
               use_random_offsets=True, rand_offset_seed=31)
    mock.reset()
    with WorkerManager(WorkerConfig(paths=[gpath], use_cufile=True, use_gds_buf_reg=True,
                                    io_depth=depth, **rnd)) as mgr:
        r = mgr.run_phase(BenchPhase.READFILES)
        rc, ow, opr = oracle_lib.run_oracle_phase(WorkerConfig(paths=[cpath], **rnd),
                                                  BenchPhase.READFILES)
        assert rc == 0
        assert r["ops_total"]["bytes"] == opr.opsTotal.numBytesDone == size
        assert r["ops_total"]["iops"] == opr.opsTotal.numIOPSDone == size // block
        assert r["verify_mismatch_bytes"] == 0 and r["verified_bytes"] == size
        st = mock.stats()
        assert st["batch_ops"] == size // block and st["read"] == 0
        assert st["batch_submit"] == (size // block) // depth  # groups of --iodepth requests
        # random writes through the batch API too (full coverage), then verify again
        w = mgr.run_phase(BenchPhase.CREATEFILES)
        assert w["ops_total"]["bytes"] == size
        assert mgr.run_phase(BenchPhase.READFILES)["verify_mismatch_bytes"] == 0
    with open(gpath, "rb") as f:
        assert f.read() == oracle_lib.fill_pattern(size, 0, 5)


def test_cufile_dir_mode_registers_every_file(workdir, mock):
    gdir, cdir = os.path.join(workdir, "g"), os.path.join(workdir, "c")
    os.mkdir(gdir)
    os.mkdir(cdir)
    common = dict(path_type=PathType.DIR, num_threads=2, num_dirs=2, num_files=2,
                  block_size=16 * KiB, file_size=48 * KiB, integrity_check_salt=1)
    mock.reset()
    with WorkerManager(WorkerConfig(paths=[gdir], use_cufile=True, **common)) as mgr:
        for phase in (BenchPhase.CREATEDIRS, BenchPhase.CREATEFILES, BenchPhase.READFILES):
            res = mgr.run_phase(phase)
            rc, ow, opr = oracle_lib.run_oracle_phase(WorkerConfig(paths=[cdir], **common), phase)
            assert rc == 0
            assert res["ops_total"]["bytes"] == opr.opsTotal.numBytesDone
            assert res["ops_total"]["entries"] == opr.opsTotal.numEntriesDone
    st = mock.stats()
    nfiles = 2 * 2 * 2
    assert st["handle_reg"] == st["handle_dereg"] == 2 * nfiles  # per file and phase (:3091)
    for root, _, files in os.walk(gdir):
        for name in files:
            rel = os.path.relpath(os.path.join(root, name), gdir)
            assert sha(os.path.join(gdir, rel)) == sha(os.path.join(cdir, rel))


def test_real_libcufile_failure_is_loud(workdir):
    """without the stand-in the worker must fail with the cuFile error, never fall back"""
    script = (
        "import os, sys; sys.path.insert(0, %r); os.environ.pop('ELB_CUFILE_LIB', None)\n"
        "from elbencho_b200 import WorkerConfig, WorkerManager, WorkerError\n"
        "try:\n"
        "    WorkerManager(WorkerConfig(paths=[%r], file_size=1<<20, use_cufile=True))\n"
        "    print('CREATED')\n"
        "except WorkerError as e:\n"
        "    print('ERROR:', e)\n" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                      os.path.join(workdir, "x")))
    env = dict(os.environ)
    env.pop("ELB_CUFILE_LIB", None)
    out = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=env,
                         timeout=120).stdout
    # either real GDS works on this box (then fine) or the error names cuFile
    assert "CREATED" in out or "cuFile" in out, out
