import sys, os
sys.path.insert(0, "telomere-analyzer_b200")
from nanotel_b200 import _lib
from nanotel_b200.synth import synth_reads
L = _lib.load()
buf, off, meta = synth_reads(100000, 20261020)
n = int(off[-1])
for t in (1, 4, 8, 12, 16):
    print("threads %2d  read %.1f GB/s   read + 1/4 NT write %.1f GB/s" % (t, L.ntl_host_read_gbs(buf.ctypes.data, n, t, 3), L.ntl_host_read_gbs(buf.ctypes.data, n, t, -3)))
