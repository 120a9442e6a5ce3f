"""Exploration: does binding the worker threads (and with them the pinned rings and the tmpfs pages
they first-touch) to the GPU-local NUMA node change write / read GiB/s?  Affinity is set on the
calling thread before the manager is created; the worker threads inherit it."""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from elbencho_b200 import BenchPhase, WorkerConfig, WorkerManager  # noqa: E402

GiB = 1 << 30
MiB = 1 << 20


def parse_cpulist(text):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            lo, hi = part.split("-")
            cpus.extend(range(int(lo), int(hi) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_topology(gpu_index):
    import torch
    bus = torch.cuda.get_device_properties(gpu_index)
    out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i",
                          str(gpu_index)], capture_output=True, text=True).stdout.strip()
    bus_id = out.lower()
    if bus_id.startswith("00000000:"):
        bus_id = bus_id[4:]
    base = "/sys/bus/pci/devices/" + bus_id
    info = {"bus_id": bus_id, "name": bus.name}
    for name in ("numa_node", "local_cpulist"):
        try:
            info[name] = open(os.path.join(base, name)).read().strip()
        except OSError as err:
            info[name] = "unreadable: %s" % err
    return info


def gpu_run(path, threads, size, gate=True):
    if os.path.exists(path):
        os.unlink(path)
    cfg = WorkerConfig(paths=[path], num_threads=threads, block_size=MiB, file_size=size,
                       integrity_check_salt=1, serialize_buffered_writes=gate)
    out = {}
    with WorkerManager(cfg) as mgr:
        for phase in (BenchPhase.CREATEFILES, BenchPhase.READFILES):
            res = mgr.run_phase(phase)
            out[phase.name[:1]] = round(
                res["ops_total"]["bytes"] / GiB / (res["last_finish_usec"] / 1e6), 2)
    os.unlink(path)
    return out


def cpu_run(path, threads, size):
    res = bench.run_cpu_localworker([path], threads, size, MiB, 1, False)
    if os.path.exists(path):
        os.unlink(path)
    return {name[:1]: round(ph["bytes"] / GiB / (ph["usec"] / 1e6), 2)
            for name, ph in res["phases"].items()}


def main():
    base = sys.argv[1] if len(sys.argv) > 1 else "/dev/shm"
    size = int(float(sys.argv[2]) * GiB) if len(sys.argv) > 2 else 16 * GiB
    path = os.path.join(base, "elb_explore_affinity.bin")
    all_cpus = sorted(os.sched_getaffinity(0))
    topo = gpu_topology(0)
    print(json.dumps({"topology": topo, "num_cpus": len(all_cpus)}), flush=True)
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout,
          file=sys.stderr)
    print(subprocess.run(["lscpu"], capture_output=True, text=True).stdout, file=sys.stderr)
    local = parse_cpulist(topo["local_cpulist"]) if "unreadable" not in topo["local_cpulist"] \
        else all_cpus
    local = [c for c in local if c in all_cpus] or all_cpus
    remote = [c for c in all_cpus if c not in local] or all_cpus
    # physical cores first: on this box SMT siblings are the upper half of the cpu numbers
    half = len(all_cpus) // 2
    local_phys = [c for c in local if c < half] or local
    sets = {"all": all_cpus, "gpu_local_node": local, "gpu_local_phys": local_phys,
            "remote_node": remote}
    gpu_run(path, 2, 1 * GiB)
    for threads in (1, 16):
        for name, cpus in sets.items():
            os.sched_setaffinity(0, cpus)
            for rep in range(2):
                res = {"threads": threads, "cpus": name, "ncpus": len(cpus), "rep": rep,
                       "gpu": gpu_run(path, threads, size)}
                if rep == 0:
                    res["cpu_localworker"] = cpu_run(path, threads, size // 2)
                print(json.dumps(res), flush=True)
            os.sched_setaffinity(0, all_cpus)


if __name__ == "__main__":
    main()
