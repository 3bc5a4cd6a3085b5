#!/usr/bin/env python
"""Differential fuzz: the default dispatch (persistent / tiled / boxed kernels) against the generic
one-thread-per-point kernels (IB200_FLAG_NO_TILES) on random problems -- lattice and volume shapes with partial
tiles, batch / channel counts, orders 1-7, per-axis bounds, extrapolation modes, f32 / f16, deformation amplitudes
from gentle to folding, coordinate shifts that leave the field of view, displacement-field mode, strided volumes.

    python profiles/fuzz_fast_vs_generic.py [cases] [seed]

Prints one line per failing case and a summary; exit code 1 if anything differs beyond the tolerance
(f32: 2e-5 of the largest generic value, 4e-5 for scatters; f16: 2e-2)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'torch-interpol_b200')); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import random  # noqa: E402
import torch  # noqa: E402
from test_gpu_ops import smooth_grid  # noqa: E402
import interpol_b200 as ib  # noqa: E402
from interpol_b200 import pushpull as pp  # noqa: E402

NO_TILES, FORCE_PIPE = 1, 8


def run(flags, fn):
    old = pp.flags
    pp.flags = flags
    try:
        out = fn()
        return out, ib.last_kernel()
    finally:
        pp.flags = old


def main(ncases=None, seed=None):
    if ncases is None:
        ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    if seed is None:
        seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = random.Random(seed)
    bad, kernels = 0, {}
    for case in range(ncases):
        gen = torch.Generator().manual_seed(seed * 100003 + case)
        while True:
            shape = (rng.randint(9, 70), rng.randint(9, 60), rng.choice([16, 20, 24, 32, 36, 40, 44, 64, 72, 96]))
            if shape[0] * shape[1] * shape[2] >= 32768:
                break
        vshape = tuple(rng.randint(5, 60) for _ in range(3))
        B, C = rng.choice([1, 1, 2]), rng.choice([1, 1, 2, 3, 4, 4])
        order = rng.choice([1, 1, 2, 3, 3, 3, 4, 5, 6, 7])
        bound = [rng.randint(0, 6) for _ in range(3)] if rng.random() < 0.7 else [rng.randint(0, 6)]
        ex = rng.choice([0, 1, 1, 2])
        half = rng.random() < 0.25
        dt = torch.float16 if half else torch.float32
        amp = rng.choice([0.5, 3.0, 3.0, 8.0, 30.0])
        disp = rng.random() < 0.2 and not half
        op = rng.choice(['pull', 'pull', 'grad', 'push', 'push', 'count'])
        force = FORCE_PIPE if rng.random() < 0.4 else 0
        grid = smooth_grid(shape, gen, amp=amp, batch=B)
        grid = grid * torch.tensor([vshape[d] / shape[d] for d in range(3)]) + rng.choice([0.0, 0.0, -2.5, 40.0])
        if rng.random() < 0.15:
            grid = grid + torch.randn(grid.shape, generator=gen) * rng.choice([0.3, 5.0])
        if disp:
            ident = torch.stack(torch.meshgrid(*[torch.arange(float(s)) for s in shape], indexing='ij'), dim=-1)
            grid = grid - ident
        grid = grid.to(dt).contiguous().cuda()
        vol = torch.randn([B, C, *vshape], generator=gen).to(dt)
        img = torch.randn([B, C, *shape], generator=gen).to(dt)
        if rng.random() < 0.2 and op in ('pull', 'grad'):
            big = torch.zeros([B, C, vshape[0], vshape[1], vshape[2] + 4], dtype=dt)     # padded rows (strided volume)
            big[..., :vshape[2]] = vol
            vol_d = big.cuda()[..., :vshape[2]]
        else:
            vol_d = vol.cuda()
        img_d = img.cuda()
        if op == 'pull':
            fn = lambda: pp.grid_pull(vol_d, grid, bound, [order], ex, disp)
        elif op == 'grad':
            fn = lambda: pp.grid_grad(vol_d, grid, bound, [order], ex, disp)
        elif op == 'push':
            fn = lambda: pp.grid_push(img_d, grid, list(vshape), bound, [order], ex, disp)
        else:
            fn = lambda: pp.grid_count(grid, list(vshape), bound, [order], ex, disp)
        fast, kf = run(force, fn)
        slow, ks = run(NO_TILES, fn)
        kernels[kf] = kernels.get(kf, 0) + 1
        scale = slow.float().abs().max().item()
        err = (fast.float() - slow.float()).abs().max().item()
        # (scatters: the generic arm accumulates with float32 atomics in arrival order -- its own noise is ~1e-5 when a
        # few hundred sources land on one voxel; case 125 of seed 1 is 4.6e-6 (boxed) / 7.8e-6 (generic) off the oracle)
        tol = (2e-2 if half else (4e-5 if op in ('push', 'count') else 2e-5)) * max(scale, 1e-30)
        if half:
            tol = max(tol, 4 * 5.96e-8)          # results that vanish: a few float16 subnormal quanta
        # pile-ups (a folding deformation splatted through `replicate` onto one face: 1e5 sources on one voxel):
        # float32 accumulation drops increments below half an ulp of the running sum in BOTH kernels and in the
        # reference (each is ~2 % off the float64 oracle there, profiles/diag/fuzz_check.py); they only agree loosely
        if op in ('push', 'count') and not half:
            typical = slow.float().abs().mean().item()
            if scale > 200 * max(typical, 1e-30):
                tol = max(tol, 2e-3 * scale)
        ok = (err <= tol) and bool(torch.isfinite(fast.float()).all() == torch.isfinite(slow.float()).all())
        if not ok:
            bad += 1
            print('MISMATCH case %d: op=%s shape=%s vshape=%s B=%d C=%d order=%d bound=%s ex=%d dtype=%s amp=%g disp=%s force=%d  %s vs %s  err %.3g scale %.3g'
                  % (case, op, shape, vshape, B, C, order, bound, ex, dt, amp, disp, force, kf, ks, err, scale), flush=True)
    print('%d cases, %d mismatches; fast kernels seen: %s' % (ncases, bad, sorted(kernels.items())))
    return 1 if bad else 0


if __name__ == '__main__':
    sys.exit(main())
