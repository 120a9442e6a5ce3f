"""Independent Python restatement of the reference's custom tree partitioning, used as the checker
for the product's PathStore (TEST INFRASTRUCTURE, like oracle/):
source/PathStore.cpp:89-155 (load + round up + size window), :192-212 (sort orders), :258-300
(non-shared: every n-th element), :322-437 (shared: contiguous block ranges), ProgArgs.cpp:2740-2803
(non-shared = size < share size, shared = the rest) and LocalWorker.cpp:1520-1560 (per worker:
non-shared sublist first, then the shared ranges)."""


def parse_tree(text):
    dirs, files = [], []
    for line in text.splitlines():
        parts = line.split(None, 1)
        if not parts:
            continue
        if parts[0] == "d" and len(parts) == 2:
            dirs.append(parts[1].strip())
        elif parts[0] == "f" and len(parts) == 2:
            size, path = parts[1].split(None, 1)
            files.append((path.strip(), int(size)))
    return dirs, files


def num_blocks(size, block):
    return size // block + (1 if size % block else 0)


def worker_dirs(dirs, rank, nthreads):
    ordered = sorted(dirs, key=lambda p: (len(p), p))
    return ordered[rank::nthreads]


def worker_files(files, rank, nthreads, block, share_size=0, round_up=0):
    """-> list of (path, totalLen, rangeStart, rangeLen)"""
    share_size = share_size or 32 * block
    sized = []
    for path, size in files:
        if round_up and size % round_up:
            size = size - size % round_up + round_up
        sized.append((path, size))
    non_shared = sorted([f for f in sized if f[1] < share_size], key=lambda f: (f[1], f[0]))
    shared = [f for f in sized if f[1] >= share_size]  # (file order of the tree file)
    out = [(p, s, 0, s) for p, s in non_shared[rank::nthreads]]

    total_blocks = sum(num_blocks(s, block) for _, s in shared)
    standard = total_blocks // nthreads
    mine = standard
    if rank == nthreads - 1 and total_blocks % nthreads:
        mine = total_blocks - standard * (nthreads - 1)
    start, end = rank * standard, rank * standard + mine
    pos = 0
    left = mine
    for path, size in shared:
        nblocks = num_blocks(size, block)
        first, last = pos, pos + nblocks - 1
        pos += nblocks
        if not left or first >= end:
            break
        if last < start:
            continue
        if start <= first:
            range_start, remaining = 0, nblocks
        else:
            inner = start - first
            range_start, remaining = inner * block, nblocks - inner
        if left < remaining:
            range_len, left = left * block, 0
        else:
            range_len, left = size - range_start, left - remaining
        out.append((path, size, range_start, range_len))
    return out
