"""Shared-memory bank model of the tap loops on the BENCH workload (256^3 smooth deformation): mean wavefronts
per warp instruction for a lane <-> voxel mapping.  Model (measured, profiles/micro/smem_micro.cu): one wavefront
serves 32 distinct banks; lanes reading the SAME word are merged (loads) or serialised (atomics); distinct words
in one bank serialise.  Rows of the box are 64 words, so bank = z mod 32."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', '..'))
from bench import make_workload

def wavefronts(words, rows, atomic):
    """words, rows: (n_instr, 32) int arrays (z word and row id per lane) -> wavefronts per instruction"""
    n = words.shape[0]
    out = np.zeros(n, dtype=np.int64)
    bank = words % 32
    key = rows.astype(np.int64) * 100000 + words           # unique id of the addressed word
    for i in range(n):
        if atomic:
            # every lane is one access; accesses to one bank serialise (same word or not)
            out[i] = np.bincount(bank[i], minlength=32).max()
        else:
            u = np.unique(np.stack([bank[i], key[i]], 1), axis=0)
            out[i] = np.bincount(u[:, 0], minlength=32).max()
    return out

def main(nrows=4000, size=256):
    vol, grid = make_workload(size, 'cpu')
    g = grid[0].numpy()
    rng = np.random.default_rng(0)
    res = {}
    xs = rng.integers(0, size, nrows); ys = rng.integers(0, size, nrows); zs = rng.integers(0, size // 32, nrows) * 32
    i0 = np.floor(g - 1.0).astype(np.int64)
    for name in ('A: 1 lane per voxel', 'B: k-split adjacent', 'C: k-split stride 2'):
        for atomic in (False, True):
            tot = 0; cnt = 0
            W, R = [], []
            for x, y, z in zip(xs, ys, zs):
                seg = i0[x, y, z:z + 32]                     # (32, 3)
                row = seg[:, 0] * 1000 + seg[:, 1]
                z0 = seg[:, 2]
                if name.startswith('A'):
                    for k in range(4):
                        W.append(z0 + k); R.append(row)
                else:
                    for p in range(2):
                        vox = (np.arange(16) + 16 * p) if name.startswith('B') else (2 * np.arange(16) + p)
                        lane_v = np.repeat(vox, 2); h = np.tile([0, 1], 16)
                        for a in range(2):
                            W.append(z0[lane_v] + 2 * h + a); R.append(row[lane_v])
            W = np.array(W); R = np.array(R)
            wf = wavefronts(W, R, atomic)
            per_instr = wf.mean()
            instr_per_32vox = 4 if name.startswith('A') else 8   # per (i, j) row of the support
            print('%-22s %-6s wavefronts/instr %.3f  -> %.1f wavefront-clk per support row per 32 voxels (x16 = %.0f per z-row)'
                  % (name, 'ATOMS' if atomic else 'LDS', per_instr, per_instr * instr_per_32vox, per_instr * instr_per_32vox * 16))

if __name__ == '__main__':
    main()
