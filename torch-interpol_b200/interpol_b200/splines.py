"""Spline-order vocabulary (reference: `interpol/splines.py:7-15`).  The basis
polynomials are evaluated in registers by the CUDA kernels (csrc/splines.cuh)."""
from enum import Enum


class InterpolationType(Enum):
    nearest = zeroth = 0
    linear = first = 1
    quadratic = second = 2
    cubic = third = 3
    fourth = 4
    fifth = 5
    sixth = 6
    seventh = 7
