"""nanorq_b200 -- B200-native hot path of the nanorq RaptorQ codec.

The product is the C-ABI shared library ``libnanorq_b200.so`` (nanorq.h API +
rqb200.h solver/row-op entry points, CUDA kernels for sm_100a).  This package is
the thin ctypes mirror of that ABI used by the tests and by bench.py.
"""
from .api import (  # noqa: F401
    Decoder,
    Encoder,
    FileIO,
    Matrix,
    MemIO,
    PinnedBuffer,
    PinnedMemIO,
    Solver,
    SolveRequest,
    block_params,
    cache_stats,
    device_mem_info,
    release_cached,
    set_cache_limit,
    device_count,
    host_profile,
    kernel_launches,
    last_error,
    lib,
    lt_row_indices,
    plan_blob,
    schedule_plan_blob,
    set_device,
    slow_path_counters,
    transfer_bytes,
    SYM_ADDED,
    SYM_DUP,
    SYM_ERR,
    SYM_IGN,
    NO_ROW,
)
