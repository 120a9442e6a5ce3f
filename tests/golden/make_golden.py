"""Generate tests/golden/ref_vectors.json from the REFERENCE's own headers.

Run in the build container (needs /root/reference and oracle/_ref/libelb_ref.so, built by
`make -C oracle`):   python tests/golden/make_golden.py

The vectors pin the oracle (and through it the product) to the reference's PRNG streams and offset
generator sequences for injected states. The reference has no golden data of its own (SURVEY.md
§4, §8c); what its tests pin is the closed-form --verify pattern, which is added here from the
closed form itself (pure Python, independent of both the oracle and the product).
"""
import ctypes
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO_ROOT)

from tests import oracle_lib  # noqa: E402

XOSHIRO_STATES = [[1, 2, 3, 4], [0x0123456789ABCDEF, 0xFEDCBA9876543210, 0xDEADBEEFCAFEBABE, 42]]
GOLDEN_SEEDS = [12345, 0xFFFFFFFFFFFFFFF1]

# (kind, numBytesTotal, len, offset, blockSize, numDataSetThreads, lcgSeed)
OFFSETGEN_CASES = [
    (0, 0, 10 * 1048576, 0, 1048576, 1, 0),
    (0, 0, 10 * 1048576 + 123, 4096, 1048576, 1, 0),
    (1, 0, 10 * 1048576, 0, 1048576, 1, 0),
    (1, 0, 5 * 65536 + 1000, 65536, 65536, 1, 0),
    (2, 40 * 4096, 1048576, 8192, 4096, 1, 0),
    (2, 10 * 4096 + 77, 100000, 0, 4096, 1, 0),
    (3, 64 * 4096, 16 * 1048576, 1048576, 4096, 1, 0),
    (3, 20 * 65536 + 5, 65536 * 7, 0, 65536, 1, 0),
    (4, 0, 8 * 4096, 3 * 4096, 4096, 4, 0),
    (5, 16 * 4096, 16 * 4096, 0, 4096, 1, 7),
    # (full coverage: one cycle only - the reference re-seeds later cycles from random_device)
    (5, 13 * 4096, 13 * 4096, 26 * 4096, 4096, 1, 0xDEADBEEF),
    (5, 1000 * 512, 1000 * 512, 512 * 3, 512, 1, 123456789),
]


OFFSETGEN_ALGO_CASES = [
    (2, 40 * 4096, 1048576, 8192, 4096, 1, 0),
    (3, 64 * 4096, 16 * 1048576, 1048576, 4096, 1, 0),
]


CUSTOM_TREE_TEXT = """# golden tree
d top
d top/sub1
d top/sub1/deep
d other
d a
f 0 top/empty.bin
f 1 top/one_byte.bin
f 5000 top/sub1/small.bin
f 65536 top/sub1/deep/one_block.bin
f 65537 a/one_block_plus.bin
f 200000 other/odd.bin
f 200000 a/same size as odd.bin
f 2097152 other/share_a.bin
f 3000001 top/share_b.bin
f 4194304 top/sub1/share_c.bin
f 2097152 a/share_d.bin
x ignored
"""


def pattern_closed_form(length, file_offset, salt):
    """byte x of the file = byte (x % 8) of little-endian u64 ((x & ~7) + salt) mod 2^64
    (LocalWorker.cpp:2091-2128)."""
    out = bytearray()
    for x in range(file_offset, file_offset + length):
        word = ((x & ~7) + salt) & 0xFFFFFFFFFFFFFFFF
        out.append((word >> ((x % 8) * 8)) & 0xFF)
    return bytes(out)


def main():
    ref = oracle_lib.load_ref()
    if ref is None:
        raise SystemExit("oracle/_ref/libelb_ref.so missing: run `make -C oracle` first")

    vectors = {"generator": "tests/golden/make_golden.py", "reference": "breuner/elbencho @ v3.1-4"}

    xo = []
    for state in XOSHIRO_STATES:
        algo = ref.ref_xoshiro256ss_create(oracle_lib.u64x4(state))
        nexts = [ref.ref_randalgo_next(algo) for _ in range(16)]
        buf = ctypes.create_string_buffer(37)
        ref.ref_randalgo_fill_buf(algo, buf, 37)
        ref.ref_randalgo_destroy(algo)
        xo.append({"state": state, "next16": nexts, "then_fill37_hex": buf.raw.hex()})
    vectors["xoshiro256ss"] = xo

    gp = []
    for seed in GOLDEN_SEEDS:
        algo = ref.ref_goldenprime_create(seed, oracle_lib.u64x4(XOSHIRO_STATES[0]))
        nexts = [ref.ref_randalgo_next(algo) for _ in range(8)]
        length = 600000 + 5  # two full reseed chunks + tail with a partial word
        buf = ctypes.create_string_buffer(length)
        ref.ref_randalgo_fill_buf(algo, buf, length)
        after = ref.ref_randalgo_next(algo)
        ref.ref_randalgo_destroy(algo)
        gp.append({"seed": seed, "seeder_state": XOSHIRO_STATES[0], "next8": nexts,
                   "fill_len": length, "fill_sha256": hashlib.sha256(buf.raw).hexdigest(),
                   "fill_first64_hex": buf.raw[:64].hex(), "fill_last16_hex": buf.raw[-16:].hex(),
                   "next_after_fill": after})
    vectors["goldenprime"] = gp

    # the four --randalgo generators behind RandAlgoInterface::next(), injected state
    # (algo ids = enum elb_offset_rand_algo: 0 balanced_single, 1 fast, 2 balanced, 3 strong)
    ra = []
    for algo_id in (0, 1, 2, 3):
        for state in XOSHIRO_STATES:
            algo = ref.ref_randalgo_create(algo_id, oracle_lib.u64x4(state))
            nexts = [ref.ref_randalgo_next(algo) for _ in range(700)]  # mt19937_64: > 2 x 312
            ref.ref_randalgo_destroy(algo)
            digest = hashlib.sha256(b"".join(v.to_bytes(8, "little") for v in nexts)).hexdigest()
            ra.append({"algo": algo_id, "state": state, "next8": nexts[:8],
                       "next700_sha256": digest, "last": nexts[-1]})
    vectors["randalgo"] = ra

    # random offset generators driven by each non-default algorithm
    oa = []
    for algo_id in (1, 2, 3):
        for case in OFFSETGEN_ALGO_CASES:
            kind, total, length, offset, block, threads, lcg = case
            seq = oracle_lib.offsetgen_sequence(ref, "ref", kind, total, length, offset, block,
                                                threads, XOSHIRO_STATES[1], lcg, rand_algo=algo_id)
            oa.append({"algo": algo_id, "kind": kind, "numBytesTotal": total, "len": length,
                       "offset": offset, "blockSize": block, "numDataSetThreads": threads,
                       "lcgSeed": lcg, "randState": XOSHIRO_STATES[1], "sequence": seq})
    vectors["offsetgen_randalgo"] = oa

    # LatencyHistogram.h (bucket rule, min/max/avg, percentile rule, the two string formats),
    # UnitTk::getPerSecFromUSec and the human-readable formats of UnitTk.cpp
    import random
    histos = []
    for seed, count, scale in ((1, 1, 1), (2, 50, 100), (3, 2000, 5000), (4, 5000, 3000000)):
        rng = random.Random(seed)
        lats = [0] if count == 1 else [int(rng.random() ** 3 * scale) for _ in range(count)]
        histo = ref.ref_histogram_create()
        for lat in lats:
            ref.ref_histogram_add(histo, lat)
        nbuckets = ref.ref_histogram_num_buckets(histo)
        buckets = (ctypes.c_uint64 * nbuckets)()
        ref.ref_histogram_buckets(histo, buckets)
        buf = ctypes.create_string_buffer(8192)
        ref.ref_histogram_str(histo, buf, len(buf))
        histo_str = buf.value.decode()
        pcts = {}
        for pct in (1.0, 50.0, 75.0, 99.0, 99.9, 99.999):
            ref.ref_histogram_percentile_str(histo, pct, buf, len(buf))
            pcts[str(pct)] = {"value": ref.ref_histogram_percentile(histo, pct),
                              "str": buf.value.decode()}
        histos.append({"latencies": lats, "num_buckets": nbuckets, "buckets": list(buckets),
                       "num": ref.ref_histogram_num(histo), "sum": ref.ref_histogram_sum(histo),
                       "min": ref.ref_histogram_min(histo), "max": ref.ref_histogram_max(histo),
                       "avg": ref.ref_histogram_avg(histo),
                       "exceeded": ref.ref_histogram_exceeded(histo), "histogram_str": histo_str,
                       "percentiles": pcts})
        ref.ref_histogram_destroy(histo)
    vectors["latency_histogram"] = histos

    units = {"latency_us": {}, "elapsed_ms": {}, "elapsed_sec": {}, "per_sec": [],
             "human_to_bytes": {}}
    buf = ctypes.create_string_buffer(256)
    for val in (0, 1, 9, 10, 99, 999, 1000, 1234, 9999, 10000, 99999, 123456, 999999, 1000000,
                1234567, 59999999, 60000000, 3599999999, 3600000000, 86400000001):
        ref.ref_unit_str(0, val, buf, len(buf))
        units["latency_us"][str(val)] = buf.value.decode()
    for val in (0, 1, 999, 1000, 1001, 59999, 60000, 61007, 3599999, 3600000, 3661001, 90061001):
        ref.ref_unit_str(1, val, buf, len(buf))
        units["elapsed_ms"][str(val)] = buf.value.decode()
    for val in (0, 1, 59, 60, 61, 3599, 3600, 3661, 90061):
        ref.ref_unit_str(2, val, buf, len(buf))
        units["elapsed_sec"][str(val)] = buf.value.decode()
    for total, usec in ((0, 1), (1, 1), (1234567, 345), (68719476736, 21220070), (1 << 60, 3),
                        (999999, 1000000), (1000001, 1000000), (7, 13)):
        units["per_sec"].append([total, usec, ref.ref_per_sec_from_usec(total, usec)])
    err = ctypes.create_string_buffer(512)
    out = ctypes.c_uint64()
    for text in ("0", "1", "4k", "4K", "1m", "64G", "2t", "1P", "1E", "512", "1.5g", "1,5g", "-4k",
                 "4x", "k", ""):
        rc = ref.ref_num_human_to_bytes(text.encode(), ctypes.byref(out), err, len(err))
        units["human_to_bytes"][text] = out.value if rc == 0 else {"error": err.value.decode()}
    units["simple128"] = {}
    for text in ("a", "secret", "correct horse battery staple", "p\u00e4ssw\u00f6rd", "x" * 200):
        ref.ref_simple128(text.encode(), buf, len(buf))
        units["simple128"][text] = buf.value.decode()
    vectors["units"] = units

    # custom tree partition by the reference's own PathStore (oracle/ref_harness_tree.cpp)
    import tempfile
    tree_text = CUSTOM_TREE_TEXT
    trees = []
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as tmp:
        tmp.write(tree_text)
        tree_path = tmp.name
    buf = ctypes.create_string_buffer(1 << 20)
    for block, share, round_up, nthreads in ((65536, 0, 0, 1), (65536, 0, 0, 3), (4096, 16384, 0, 4),
                                            (65536, 0, 4096, 2), (65536, 1 << 62, 0, 5)):
        per_rank = []
        for rank in range(nthreads):
            lists = {}
            for kind, name in ((0, "dirs"), (1, "files")):
                res = ref.ref_custom_tree_worker_list(tree_path.encode(), block, share, round_up,
                                                      rank, nthreads, kind, buf, len(buf))
                assert res >= 0, buf.value
                lists[name] = buf.value.decode()
            per_rank.append(lists)
        trees.append({"blockSize": block, "fileShareSize": share, "treeRoundUpSize": round_up,
                      "numDataSetThreads": nthreads, "per_rank": per_rank})
    os.unlink(tree_path)
    vectors["custom_tree"] = {"tree_text": tree_text, "cases": trees}

    og = []
    for case in OFFSETGEN_CASES:
        kind, total, length, offset, block, threads, lcg = case
        seq = oracle_lib.offsetgen_sequence(ref, "ref", kind, total, length, offset, block,
                                            threads, XOSHIRO_STATES[1], lcg)
        og.append({"kind": kind, "numBytesTotal": total, "len": length, "offset": offset,
                   "blockSize": block, "numDataSetThreads": threads, "lcgSeed": lcg,
                   "randState": XOSHIRO_STATES[1], "sequence": seq})
    vectors["offsetgen"] = og

    pat = []
    for length, off, salt in [(16, 0, 1), (24, 5, 1), (7, 13, 0xFFFFFFFFFFFFFFFF), (40, 1048573, 42),
                              (3, 6, 0x0102030405060708), (17, 0xFFFFFFFFFFFFFFF0 - 8, 99)]:
        pat.append({"len": length, "fileOffset": off, "salt": salt,
                    "hex": pattern_closed_form(length, off, salt).hex()})
    vectors["pattern_closed_form"] = pat

    out_path = os.path.join(HERE, "ref_vectors.json")
    with open(out_path, "w") as f:
        json.dump(vectors, f, indent=1)
    print("wrote", out_path, os.path.getsize(out_path), "bytes")


if __name__ == "__main__":
    main()
