"""Box sizes of the scatter / gather tiles on the BENCH workload: words of the bounding box of all supports of a
TX x TY x 32 tile (z origin aligned to 4, z stride rounded up to `zr`), cubic order."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..'))
from bench import make_workload
size = 256
vol, grid = make_workload(size, 'cpu')
i0 = np.floor(grid[0].numpy() - 1.0).astype(np.int64)
for (TX, TY, TZ) in ((8, 8, 32), (4, 8, 32), (4, 4, 32), (8, 8, 16), (2, 8, 32), (8, 16, 16), (4, 16, 16)):
    t = i0.reshape(size // TX, TX, size // TY, TY, size // TZ, TZ, 3)
    lo = t.min(axis=(1, 3, 5)); hi = t.max(axis=(1, 3, 5)) + 3
    lo[..., 2] &= ~3
    ext = hi - lo + 1
    for zr in (4, 32):
        sz = (ext[..., 2] + zr - 1) // zr * zr
        words = ext[..., 0] * ext[..., 1] * sz
        q = np.percentile(words, [50, 90, 99, 100])
        print('tile %dx%dx%d zr %2d: ext mean %s max %s | words p50 %.0f p90 %.0f p99 %.0f max %.0f | words/source mean %.2f | fit<=10240: %.3f  <=16384: %.3f <=8192 %.3f'
              % (TX, TY, TZ, zr, ext.reshape(-1, 3).mean(0).round(1), ext.reshape(-1, 3).max(0), *q, words.mean() / (TX * TY * TZ),
                 (words <= 10240).mean(), (words <= 16384).mean(), (words <= 8192).mean()))
