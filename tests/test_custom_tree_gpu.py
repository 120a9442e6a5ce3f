"""GPU parity tests of custom tree mode (--treefile; LocalWorker.cpp:2927-3010, 3261-3470): dirs and
files of different sizes from a tree file, small files whole per worker, large files shared as
block ranges. Written with --verify, so every file must equal the closed-form pattern whatever
worker wrote which range; counters are checked against the independent partition model."""
import os
import shutil
import subprocess
import tempfile

import pytest

from elbencho_b200 import BenchPhase, PathType, WorkerConfig, WorkerError, WorkerManager
from elbencho_b200.build import CLI_PATH
from tests import oracle_lib, tree_model

pytestmark = pytest.mark.gpu

KiB = 1 << 10
MiB = 1 << 20

TREE_TEXT = """# test tree
d top
d top/sub1
d top/sub1/deep
d other
f 0 top/empty.bin
f 1 top/one_byte.bin
f 5000 top/sub1/small.bin
f 65536 top/sub1/deep/one_block.bin
f 200000 other/odd.bin
f 2097152 other/share_a.bin
f 3000001 top/share_b.bin
f 4194304 top/sub1/share_c.bin
"""


@pytest.fixture()
def workdir(cuda_device):
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    path = tempfile.mkdtemp(prefix="elb_tree_", dir=base)
    yield path
    shutil.rmtree(path, ignore_errors=True)


def expected_counters(files, nthreads, block, share=0, round_up=0):
    entries = iops = total = 0
    for rank in range(nthreads):
        for _, size, _, length in tree_model.worker_files(files, rank, nthreads, block, share,
                                                          round_up):
            total += length
            iops += max(1, tree_model.num_blocks(length, block)) if length else 0
            entries += 1 if size == length else 0
    return entries, iops, total


@pytest.mark.parametrize("threads,extra", [
    (1, {}), (3, {}), (4, dict(io_depth=4)), (3, dict(use_custom_tree_randomize=True,
                                                      tree_randomize_seed=7)),
    (2, dict(tree_round_up_size=4096)),
])
def test_tree_full_cycle(workdir, threads, extra):
    block, salt = 64 * KiB, 11
    tree_path = os.path.join(workdir, "tree.txt")
    bench_dir = os.path.join(workdir, "bench")
    os.mkdir(bench_dir)
    with open(tree_path, "w") as f:
        f.write(TREE_TEXT)
    dirs, files = tree_model.parse_tree(TREE_TEXT)
    round_up = extra.get("tree_round_up_size", 0)
    sizes = {p: (s if not (round_up and s % round_up) else s - s % round_up + round_up)
             for p, s in files}
    entries, iops, total = expected_counters(files, threads, block, 0, round_up)
    assert total == sum(sizes.values())
    cfg = WorkerConfig(paths=[bench_dir], path_type=PathType.DIR, num_threads=threads,
                       block_size=block, integrity_check_salt=salt, tree_file_path=tree_path,
                       **extra)
    with WorkerManager(cfg) as mgr:
        res = mgr.run_phase(BenchPhase.CREATEDIRS)
        assert res["ops_total"]["entries"] == len(dirs)
        for d in dirs:
            assert os.path.isdir(os.path.join(bench_dir, d))

        res = mgr.run_phase(BenchPhase.CREATEFILES)
        assert res["ops_total"]["bytes"] == total
        assert res["ops_total"]["iops"] == iops
        assert res["ops_total"]["entries"] == entries
        assert res["filled_bytes"] == total
        for path, size in sizes.items():
            with open(os.path.join(bench_dir, path), "rb") as f:
                data = f.read()
            assert len(data) == size, path
            assert data == oracle_lib.fill_pattern(size, 0, salt), path

        res = mgr.run_phase(BenchPhase.STATFILES)
        assert res["ops_total"]["entries"] == entries

        res = mgr.run_phase(BenchPhase.READFILES)
        assert res["ops_total"]["bytes"] == total
        assert res["ops_total"]["iops"] == iops
        assert res["ops_total"]["entries"] == entries
        assert res["verified_bytes"] == total
        assert res["verify_mismatch_bytes"] == 0

        res = mgr.run_phase(BenchPhase.DELETEFILES)
        assert res["ops_total"]["entries"] == entries
        for path in sizes:
            assert not os.path.exists(os.path.join(bench_dir, path)), path

        res = mgr.run_phase(BenchPhase.DELETEDIRS)
        assert res["ops_total"]["entries"] == len(dirs)  # the first worker removes all dirs
        assert os.listdir(bench_dir) == []


def test_tree_read_finds_corruption_in_a_shared_file(workdir):
    block, salt = 64 * KiB, 3
    tree_path = os.path.join(workdir, "tree.txt")
    bench_dir = os.path.join(workdir, "bench")
    os.mkdir(bench_dir)
    with open(tree_path, "w") as f:
        f.write(TREE_TEXT)
    cfg = WorkerConfig(paths=[bench_dir], path_type=PathType.DIR, num_threads=3, block_size=block,
                       integrity_check_salt=salt, tree_file_path=tree_path)
    with WorkerManager(cfg) as mgr:
        mgr.run_phase(BenchPhase.CREATEDIRS)
        mgr.run_phase(BenchPhase.CREATEFILES)
        victim = os.path.join(bench_dir, "top/share_b.bin")
        pos = 2500017
        with open(victim, "r+b") as f:
            f.seek(pos)
            byte = f.read(1)
            f.seek(pos)
            f.write(bytes([byte[0] ^ 1]))
        with pytest.raises(WorkerError) as err:
            mgr.run_phase(BenchPhase.READFILES)
        # (LocalWorker.cpp:2174-2177: offset of the bad byte in the file)
        assert "Data verification failed. Offset: %d;" % pos in str(err.value)


def test_cli_treescan_then_treefile_run(workdir):
    """scan an existing directory into a tree file, recreate it elsewhere from that tree file"""
    src = os.path.join(workdir, "src")
    os.makedirs(os.path.join(src, "x", "y"))
    for name, size in (("x/a.bin", 300000), ("x/y/b.bin", 70000), ("c.bin", 9 * MiB)):
        with open(os.path.join(src, name), "wb") as f:
            f.write(b"\0" * size)
    tree_path = os.path.join(workdir, "scan.txt")
    res = subprocess.run([CLI_PATH, "--treescan", src, "--treefile", tree_path],
                         capture_output=True, text=True, timeout=60)
    assert res.returncode == 0, res.stderr
    assert "Dirs: 2; Files: 3; Bytes: %d" % (300000 + 70000 + 9 * MiB) in res.stdout
    dst = os.path.join(workdir, "dst")
    os.mkdir(dst)
    res = subprocess.run([CLI_PATH, "-d", "-w", "-r", "--stat", "-t", "3", "-b", "128K", "--verify",
                          "5", "--sharesize", "1M", "--treefile", tree_path, "--gpuids", "0",
                          "--nolive", dst], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    for name, size in (("x/a.bin", 300000), ("x/y/b.bin", 70000), ("c.bin", 9 * MiB)):
        with open(os.path.join(dst, name), "rb") as f:
            assert f.read() == oracle_lib.fill_pattern(size, 0, 5), name
    res = subprocess.run([CLI_PATH, "-F", "-D", "-t", "3", "--treefile", tree_path, "--gpuids", "0",
                          "--nolive", dst], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert os.listdir(dst) == []


def test_tree_mode_through_two_services(workdir):
    """the master uploads the tree file to every service (/preparefile), the services share the
    data set: ranks 0-1 on the first, 2-3 on the second (RemoteWorker.cpp:286-330)"""
    import socket
    import time

    def free_port():
        with socket.socket() as sock:
            sock.bind(("127.0.0.1", 0))
            return sock.getsockname()[1]

    tree_path = os.path.join(workdir, "tree.txt")
    bench_dir = os.path.join(workdir, "bench")
    os.mkdir(bench_dir)
    with open(tree_path, "w") as f:
        f.write(TREE_TEXT)
    _, files = tree_model.parse_tree(TREE_TEXT)
    ports = [free_port(), free_port()]
    hosts = ",".join("127.0.0.1:%d" % p for p in ports)
    services = [subprocess.Popen([CLI_PATH, "--service", "--foreground", "--port", str(p)],
                                 stdout=subprocess.PIPE, stderr=subprocess.PIPE) for p in ports]
    try:
        res = subprocess.run([CLI_PATH, "-d", "-w", "-r", "--stat", "-t", "2", "-b", "64K", "--verify",
                              "9", "--treefile", tree_path, "--gpuids", "0", "--hosts", hosts,
                              "--svcwait", "30", "--nolive", bench_dir],
                             capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stdout + res.stderr
        for path, size in files:
            with open(os.path.join(bench_dir, path), "rb") as f:
                assert f.read() == oracle_lib.fill_pattern(size, 0, 9), path
        res = subprocess.run([CLI_PATH, "-F", "-D", "-t", "2", "--treefile", tree_path, "--gpuids",
                              "0", "--hosts", hosts, "--nolive", bench_dir],
                             capture_output=True, text=True, timeout=300)
        assert res.returncode == 0, res.stdout + res.stderr
        assert os.listdir(bench_dir) == []
    finally:
        subprocess.run([CLI_PATH, "--quit", "--hosts", hosts], capture_output=True, timeout=60)
        for svc in services:
            try:
                svc.wait(timeout=20)
            except subprocess.TimeoutExpired:
                svc.kill()
