"""Boundary-condition vocabulary (codes shared with the C ABI).

The index / sign maps themselves are evaluated in registers by the CUDA kernels
(csrc/support.cuh); this module only carries the enums of the reference's
`interpol/bounds.py:8-21` so user code written against it keeps working.
"""
from enum import Enum


class BoundType(Enum):
    zero = zeros = 0
    replicate = nearest = 1
    dct1 = mirror = 2
    dct2 = reflect = 3
    dst1 = antimirror = 4
    dst2 = antireflect = 5
    dft = wrap = 6


class ExtrapolateType(Enum):
    no = 0     # threshold: (0, n-1)
    yes = 1
    hist = 2   # threshold: (-0.5, n-0.5)
