"""Search over lane <-> voxel tilings (LX x LY x LZ lanes of a warp) and box row / plane skews for the mapping
with the fewest shared-memory wavefronts per tap on the BENCH workload.  Same bank model as bank_sim.py."""
import sys, os, itertools
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..'))
from bench import make_workload

def count_wavefronts(bank, key, atomic):
    n = bank.shape[0]
    order = np.lexsort((key, bank))
    rows = np.arange(n)[:, None]
    sb = bank[rows, order]; sk = key[rows, order]
    if atomic:
        new = np.ones_like(sb, dtype=np.int64)
    else:
        new = np.ones_like(sb, dtype=np.int64)
        new[:, 1:] = ((sb[:, 1:] != sb[:, :-1]) | (sk[:, 1:] != sk[:, :-1])).astype(np.int64)
    cnt = np.zeros((n, 32), dtype=np.int64)
    np.add.at(cnt, (np.repeat(rows, 32, 1), sb), new)
    return cnt.max(1)

def main(nwarps=3000, size=256):
    vol, grid = make_workload(size, 'cpu')
    g = grid[0].numpy()
    i0 = np.floor(g - 1.0).astype(np.int64)
    rng = np.random.default_rng(1)
    shapes = [(1, 1, 32), (1, 2, 16), (1, 4, 8), (2, 2, 8), (1, 8, 4), (2, 4, 4), (4, 4, 2), (2, 1, 16), (4, 1, 8)]
    results = []
    for (LX, LY, LZ) in shapes:
        x = rng.integers(0, size // LX, nwarps) * LX; y = rng.integers(0, size // LY, nwarps) * LY; z = rng.integers(0, size // LZ, nwarps) * LZ
        lx, ly, lz = np.meshgrid(np.arange(LX), np.arange(LY), np.arange(LZ), indexing='ij')
        lx, ly, lz = lx.ravel(), ly.ravel(), lz.ravel()
        seg = i0[x[:, None] + lx[None], y[:, None] + ly[None], z[:, None] + lz[None]]     # (n, 32, 3)
        rx, ry, z0 = seg[..., 0], seg[..., 1], seg[..., 2]
        best = {}
        for sy in range(0, 32):
            for sx in ([0] if LX == 1 else range(0, 32)):
                for atomic in (False, True):
                    tot = 0.0
                    for k in range(4):
                        word = z0 + k
                        bank = (word + sy * ry + sx * rx) % 32
                        key = (rx * 4096 + ry) * 100000 + word
                        tot += count_wavefronts(bank, key, atomic).mean()
                    tot /= 4
                    kk = 'ATOMS' if atomic else 'LDS'
                    if kk not in best or tot < best[kk][0]:
                        best[kk] = (tot, sx, sy)
        # baseline skew 0
        for kk in best:
            print('lanes %dx%dx%-2d %-5s best wavefronts/instr %.3f at row skew sy=%d, plane skew sx=%d' % (LX, LY, LZ, kk, best[kk][0], best[kk][2], best[kk][1]))
        sys.stdout.flush()

if __name__ == '__main__':
    main()
