"""Differential fuzz of the fast kernels (persistent pull / grad, boxed push / count, tile kernels) against the generic
one-thread-per-point kernels on seeded random problems: shapes with partial tiles, batch / channel counts, orders 1-7,
per-axis bounds, extrapolation modes, f32 / f16, deformations from gentle to folding, coordinates that leave the field
of view, displacement-field mode, strided volumes (profiles/fuzz_fast_vs_generic.py; 2000 further cases were run by
hand, profiles/r2r).  The generic kernels are themselves checked against the float64 oracle in test_gpu_ops.py."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles'))


@pytest.mark.parametrize('seed', [11, 12])
def test_fast_kernels_match_generic_on_random_problems(seed, capsys):
    import fuzz_fast_vs_generic as fz
    failures = fz.main(ncases=120, seed=seed)
    out = capsys.readouterr().out
    assert failures == 0, out[-3000:]
