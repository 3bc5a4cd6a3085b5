"""Wavefronts per shared atomic / load as a function of the row skew (row stride mod 32) for z-row warps."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..'))
from bench import make_workload
from bank_search import count_wavefronts
size = 256
vol, grid = make_workload(size, 'cpu')
i0 = np.floor(grid[0].numpy() - 1.0).astype(np.int64)
rng = np.random.default_rng(2)
n = 3000
for (LY, LZ) in ((1, 32), (2, 16)):
    x = rng.integers(0, size, n); y = rng.integers(0, size // LY, n) * LY; z = rng.integers(0, size // LZ, n) * LZ
    ly, lz = np.meshgrid(np.arange(LY), np.arange(LZ), indexing='ij'); ly, lz = ly.ravel(), lz.ravel()
    seg = i0[x[:, None], y[:, None] + ly[None], z[:, None] + lz[None]]
    rx, ry, z0 = seg[..., 0], seg[..., 1], seg[..., 2]
    for sx_mode in ('sx = 0', 'sx = 13 * sy (plane = 13 rows)'):
        line = []
        for sy in range(0, 32, 4):
            sx = 0 if sx_mode.startswith('sx = 0') else (13 * sy) % 32
            res = []
            for atomic in (False, True):
                tot = 0.0
                for k in range(4):
                    word = z0 + k
                    tot += count_wavefronts((word + sy * ry + sx * rx) % 32, (rx * 4096 + ry) * 100000 + word, atomic).mean()
                res.append(tot / 4)
            line.append('sy=%2d: %.2f/%.2f' % (sy, res[0], res[1]))
        print('lanes %dx%-2d %-32s LDS/ATOMS  ' % (LY, LZ, sx_mode) + '  '.join(line))
