"""Round 2 exploration, third pass: write gate tuning (near distance, L2 prefetch while queued) on
the single-file sequential write, one GPU. Each variant runs in its own process (the knobs are
read once per process from the environment)."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import json, os, sys
sys.path.insert(0, %r)
from scripts.explore_r2_pipeline import GiB, MiB, gpu_seq
path = os.path.join(sys.argv[1], "elb_explore_gate.bin")
gpu_seq(path, 1 * GiB, 4)
out = []
for rep in range(2):
    out.append(gpu_seq(path, int(sys.argv[2]) * GiB, 16, pipeline_batch_blocks=int(sys.argv[3])))
os.unlink(path)
print(json.dumps(out))
''' % REPO


def main():
    base = sys.argv[1] if len(sys.argv) > 1 else "/dev/shm"
    size = sys.argv[2] if len(sys.argv) > 2 else "16"
    for near in ("2", "3", "4"):
        for prefetch in ("1", "0"):
            for bb in ("0", "1", "4"):
                if bb != "0" and (near != "3" or prefetch != "1"):
                    continue
                env = dict(os.environ, ELB_GATE_NEAR=near, ELB_GATE_PREFETCH=prefetch)
                res = subprocess.run([sys.executable, "-c", CHILD, base, size, bb], env=env,
                                     capture_output=True, text=True)
                try:
                    runs = json.loads(res.stdout.strip().splitlines()[-1])
                except Exception:
                    runs = {"error": res.stderr[-300:]}
                print(json.dumps({"test": "gate", "near": int(near), "prefetch": int(prefetch),
                                  "batch_blocks": int(bb), "runs": runs}), flush=True)


if __name__ == "__main__":
    main()
