"""B200-native ERP descriptor matching + eight-point essential-matrix hot path.

Everything computes in ``lib/liberp_b200.so`` (hand-written sm_100a CUDA behind the C ABI
of ``include/erp_b200.h``).  This package is the thin Python face of that ABI used by the
tests and ``bench.py``; the C++ drop-in classes live under ``host/``.
"""
from .binding import (  # noqa: F401
    DMATCH, ENGINE_AUTO, ENGINE_EXACT_SIMT, ENGINE_TCGEN05, ENGINE_TCGEN05_1X, METRIC_ALGEBRAIC, METRIC_ANGULAR,
    METRIC_SAMPSON, POSE_FLOATS, Context, ErpError, Group, LIB_PATH, comm_unique_id, lib, libstdcxx_sample_table, shard_range,
)

__all__ = ["Context", "Group", "ErpError", "DMATCH", "lib", "LIB_PATH", "libstdcxx_sample_table", "comm_unique_id", "shard_range"]
