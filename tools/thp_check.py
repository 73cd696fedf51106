"""Are large numpy buffers (the bench's ASCII input) backed by transparent huge pages on this host?"""
import re

import numpy as np

print(open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), "|", open("/sys/kernel/mm/transparent_hugepage/defrag").read().strip())
a = np.empty(1 << 30, np.uint8)
a[:] = 1
s = open("/proc/self/smaps").read()
print("AnonHugePages of this process: %.0f MB (1024 MB array touched)" % (sum(int(x) for x in re.findall(r"AnonHugePages:\s+(\d+) kB", s)) / 1024))
