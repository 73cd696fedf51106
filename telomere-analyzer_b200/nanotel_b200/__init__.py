"""nanotel_b200 -- Python host mirror of NanoTel.R's interface over libnanotel_b200.so (B200, sm_100a).

The compute path is the CUDA library; nothing in this package computes telomere calls on the CPU.
"""
from ._lib import (LIB_PATH, OPT_DEBUG_STAGES, OPT_NO_JIT, OPT_REQUIRE_JIT, READ_FILTERED, READ_IUPAC, READ_KEEP,
                   READ_NO_WINDOWS, READ_REF_ERROR, RESULT_DTYPE)
from .scanner import NanoTelError, Scanner, assign_serials, count_windows

__all__ = ["Scanner", "NanoTelError", "assign_serials", "count_windows", "RESULT_DTYPE", "LIB_PATH",
           "READ_KEEP", "READ_FILTERED", "READ_REF_ERROR", "READ_NO_WINDOWS", "READ_IUPAC",
           "OPT_NO_JIT", "OPT_REQUIRE_JIT", "OPT_DEBUG_STAGES"]
