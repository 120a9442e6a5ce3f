"""Kernel-level entry points of the C ABI on raw device pointers.

Arguments are plain integers (device addresses, CUDA stream handles) so that any allocator can be
used; tests and bench.py pass ``tensor.data_ptr()`` and ``torch.cuda.current_stream().cuda_stream``.
These replace the reference's CPU block modifiers (source/workers/LocalWorker.cpp:2091-2277).
"""
import ctypes

from . import _native
from ._native import BlockDesc, VerifyResult, DEVCTR_NUM  # noqa: F401

RANDALGO_SPLITMIX64 = 0

DEVCTR_VERIFY_MISMATCH_BYTES = 0
DEVCTR_VERIFIED_BYTES = 1
DEVCTR_FILLED_BYTES = 2

BLOCK_DESC_BYTES = ctypes.sizeof(BlockDesc)        # 32
VERIFY_RESULT_BYTES = ctypes.sizeof(VerifyResult)  # 16


class KernelError(RuntimeError):
    pass


def _check(res, what):
    if res != 0:
        raise KernelError("%s failed: %s" % (what, _native.last_error()))


def fill_pattern(dev_ptr, length, file_offset, salt, stream=0):
    """K1 (replaces preWriteIntegrityCheckFillBuf, LocalWorker.cpp:2091-2128)."""
    _check(_native.load().elb_fill_pattern(dev_ptr, length, file_offset, salt, stream),
           "elb_fill_pattern")


def verify_pattern(dev_ptr, length, file_offset, salt, dev_result_ptr, stream=0):
    """K2 (replaces postReadIntegrityCheckVerifyBuf, LocalWorker.cpp:2137-2179).
    dev_result_ptr: device address of 16 bytes {numMismatchBytes, firstMismatchIdx}."""
    _check(_native.load().elb_verify_pattern(dev_ptr, length, file_offset, salt, dev_result_ptr,
                                             stream), "elb_verify_pattern")


def fill_random(dev_ptr, length, pct, seed, block_counter, stream=0, algo=RANDALGO_SPLITMIX64):
    """K3 (replaces preWriteBufRandRefillCuda, LocalWorker.cpp:2236-2277)."""
    _check(_native.load().elb_fill_random(dev_ptr, length, pct, seed, block_counter, algo, stream),
           "elb_fill_random")


def fill_pattern_batch(dev_descs_ptr, num_descs, salt, dev_counters_ptr=0, stream=0,
                       total_bytes=0, max_block_len=0):
    """total_bytes / max_block_len: size hints (0 = unknown) that pick the launch shape: both
    given and (nearly) uniform blocks -> hardware-scheduled tiles, else a persistent grid."""
    _check(_native.load().elb_fill_pattern_batch_sized(dev_descs_ptr, num_descs, salt,
                                                       dev_counters_ptr or None, total_bytes,
                                                       max_block_len, stream),
           "elb_fill_pattern_batch")


def verify_pattern_batch(dev_descs_ptr, num_descs, salt, dev_results_ptr, dev_counters_ptr=0,
                         stream=0, total_bytes=0, max_block_len=0):
    _check(_native.load().elb_verify_pattern_batch_sized(dev_descs_ptr, num_descs, salt,
                                                         dev_results_ptr,
                                                         dev_counters_ptr or None, total_bytes,
                                                         max_block_len, stream),
           "elb_verify_pattern_batch")


def fill_random_batch(dev_descs_ptr, num_descs, pct, seed, dev_counters_ptr=0, stream=0,
                      total_bytes=0, algo=RANDALGO_SPLITMIX64, max_block_len=0):
    _check(_native.load().elb_fill_random_batch_sized(dev_descs_ptr, num_descs, pct, seed, algo,
                                                      dev_counters_ptr or None, total_bytes,
                                                      max_block_len, stream),
           "elb_fill_random_batch")


def fill_pattern_staged(descs_ptr, num_descs, salt, host_delta, dev_counters_ptr=0, stream=0,
                        total_bytes=0, max_block_len=0):
    """K1 + stage-out: the block goes to the device buffer and to (devPtr + host_delta)."""
    _check(_native.load().elb_fill_pattern_staged(descs_ptr, num_descs, salt, host_delta,
                                                  dev_counters_ptr or None, total_bytes,
                                                  max_block_len, stream),
           "elb_fill_pattern_staged")


def fill_random_staged(descs_ptr, num_descs, pct, seed, host_delta, dev_counters_ptr=0, stream=0,
                       total_bytes=0, max_block_len=0, algo=RANDALGO_SPLITMIX64):
    _check(_native.load().elb_fill_random_staged(descs_ptr, num_descs, pct, seed, algo, host_delta,
                                                 dev_counters_ptr or None, total_bytes,
                                                 max_block_len, stream),
           "elb_fill_random_staged")


def verify_pattern_staged(descs_ptr, num_descs, salt, host_delta, dev_results_ptr,
                          host_results_ptr=0, dev_ticket_ptr=0, dev_counters_ptr=0, stream=0,
                          total_bytes=0, max_block_len=0):
    """stage-in + K2: the block is read from (devPtr + host_delta), stored to the device buffer
    and compared; with host_results_ptr/dev_ticket_ptr the results are published to pinned host
    memory by the last CTA of the launch."""
    _check(_native.load().elb_verify_pattern_staged(descs_ptr, num_descs, salt, host_delta,
                                                    dev_results_ptr, host_results_ptr or None,
                                                    dev_ticket_ptr or None,
                                                    dev_counters_ptr or None, total_bytes,
                                                    max_block_len, stream),
           "elb_verify_pattern_staged")


def stage_copy(descs_ptr, num_descs, host_to_device, host_delta, stream=0, total_bytes=0,
               max_block_len=0):
    _check(_native.load().elb_stage_copy(descs_ptr, num_descs, int(bool(host_to_device)),
                                         host_delta, total_bytes, max_block_len, stream),
           "elb_stage_copy")


def verify_results_init(dev_results_ptr, num_descs, stream=0):
    _check(_native.load().elb_verify_results_init(dev_results_ptr, num_descs, stream),
           "elb_verify_results_init")


def num_kernel_launches():
    return int(_native.load().elb_num_kernel_launches())


def pack_block_descs(blocks):
    """blocks: iterable of (dev_ptr, len, file_offset, block_counter) -> bytes of elb_block_desc[]
    (to be copied into device-readable memory by the caller)."""
    blocks = list(blocks)
    arr = (BlockDesc * len(blocks))()
    for i, (ptr, length, off, ctr) in enumerate(blocks):
        arr[i].devPtr = ptr
        arr[i].len = length
        arr[i].fileOffset = off
        arr[i].blockCounter = ctr
    return bytes(arr)
