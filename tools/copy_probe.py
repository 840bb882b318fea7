"""Host staging-copy throughput (dipsb_host_copy2d) for a 1080p RGBA frame, by DIPSB_COPY_THREADS.  Host only."""
import os
import subprocess
import sys

CHILD = r'''
import numpy as np, time
from dips_b200 import _lib
lib = _lib.load()
n = 8294400
src = np.random.default_rng(0).integers(0, 256, n * 8, dtype=np.uint8).reshape(8, n)
dst = np.zeros((2, n), np.uint8)
best = 0
for rep in range(5):
    t = time.perf_counter()
    for k in range(40):
        lib.dipsb_host_copy2d(dst[k & 1].ctypes.data, n, src[k % 8].ctypes.data, n, n, 1)
    best = max(best, 40 * n / (time.perf_counter() - t) / 1e9)
print("threads %d  %.1f GB/s  (%.3f ms per 8.3 MB frame)" % (lib.dipsb_host_copy_threads(), best, n / best / 1e6))
'''

if __name__ == "__main__":
    print("host cores:", os.cpu_count())
    for t in (1, 2, 3, 4, 6, 8):
        env = dict(os.environ, DIPSB_COPY_THREADS=str(t))
        sys.stdout.write(subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True).stdout)
