"""Compile the NVRTC specialisation of the scan kernel for a pattern set (no GPU needed) and write the cubin, so that
its SASS can be inspected with cuobjdump:   python tools/jit_dump.py OUT.cubin PATTERNS [TVR] [S]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "telomere-analyzer_b200"))
from nanotel_b200 import _lib as _l  # noqa: E402

out, pats = sys.argv[1], sys.argv[2]
tvr = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3] else None
S = int(sys.argv[4]) if len(sys.argv) > 4 else 100
L = _l.load()
P = _l.make_params(pats, tvr, subseq_length=S, rc=True)
log = C.create_string_buffer(1 << 16)
n = L.ntl_jit_compile_check(C.byref(P), b"sm_100a", log, 1 << 16, out.encode())
print(n, log.value.decode(errors="replace")[:2000])
