// B-spline basis functions of order 0..7 and their first / second derivatives,
// evaluated in registers.  Same piecewise polynomials (same knots, same
// comparison direction at the knots) as the reference's Spline.fastweight /
// fastgrad / fasthess (interpol/splines.py:30-80, 90-139, 149-195) so that
// discontinuous pieces (order-2 curvature, order-1 slope) pick the same side.
#pragma once
#include "common.cuh"

namespace ib200 {

template <typename R> __device__ __forceinline__ R sq(R x) { return x * x; }
template <typename R> __device__ __forceinline__ R cu(R x) { return x * x * x; }

// B_order(|x|): no truncation outside the support (callers only evaluate it
// at the order+1 nodes of the support, like nd.py:60-61).
template <typename R>
__device__ __forceinline__ R spline_weight(int order, R x) {
    x = fabs(x);
    switch (order) {
    case 0: return R(1);
    case 1: return R(1) - x;
    case 2: return x < R(0.5) ? R(0.75) - x * x : R(0.5) * sq(R(1.5) - x);
    case 3: return x < R(1) ? (x * x * (x - R(2)) * R(3) + R(4)) * R(1. / 6.)
                            : cu(R(2) - x) * R(1. / 6.);
    case 4: {
        if (x < R(0.5)) { R y = x * x; return y * (y * R(0.25) - R(0.625)) + R(115. / 192.); }
        if (x < R(1.5)) return x * (x * (x * (R(5) - x) * R(1. / 6.) - R(1.25)) + R(5. / 24.)) + R(55. / 96.);
        return sq(sq(x - R(2.5))) * R(1. / 24.);
    }
    case 5: {
        if (x < R(1)) { R y = x * x; return y * (y * (R(0.25) - x * R(1. / 12.)) - R(0.5)) + R(0.55); }
        if (x < R(2)) return x * (x * (x * (x * (x * R(1. / 24.) - R(0.375)) + R(1.25)) - R(1.75)) + R(0.625)) + R(0.425);
        R y = R(3) - x; return sq(sq(y)) * y * R(1. / 120.);
    }
    case 6: {
        if (x < R(0.5)) { R y = x * x; return y * (y * (R(7. / 48.) - y * R(1. / 36.)) - R(77. / 192.)) + R(5887. / 11520.); }
        if (x < R(1.5)) return x * (x * (x * (x * (x * (x * R(1. / 48.) - R(7. / 48.)) + R(0.328125)) - R(35. / 288.)) - R(91. / 256.)) - R(7. / 768.)) + R(7861. / 15360.);
        if (x < R(2.5)) return x * (x * (x * (x * (x * (R(7. / 60.) - x * R(1. / 120.)) - R(0.65625)) + R(133. / 72.)) - R(2.5703125)) + R(1267. / 960.)) + R(1379. / 7680.);
        return sq(cu(x - R(3.5))) * R(1. / 720.);
    }
    case 7: {
        if (x < R(1)) { R y = x * x; return y * (y * (y * (x * R(1. / 144.) - R(1. / 36.)) + R(1. / 9.)) - R(1. / 3.)) + R(151. / 315.); }
        if (x < R(2)) return x * (x * (x * (x * (x * (x * (R(0.05) - x * R(1. / 240.)) - R(7. / 30.)) + R(0.5)) - R(7. / 18.)) - R(0.1)) - R(7. / 90.)) + R(103. / 210.);
        if (x < R(3)) return x * (x * (x * (x * (x * (x * (x * R(1. / 720.) - R(1. / 36.)) + R(7. / 30.)) - R(19. / 18.)) + R(49. / 18.)) - R(23. / 6.)) + R(217. / 90.)) - R(139. / 630.);
        R y = R(4) - x; return sq(cu(y)) * y * R(1. / 5040.);
    }
    }
    return R(0);
}

// d/dx B_order(x) = B'_order(|x|) * sign(x)   (orders >= 2; order 0 -> 0,
// order 1 is handled by the caller because the reference has two behaviours)
template <typename R>
__device__ __forceinline__ R spline_grad(int order, R xs) {
    const R x = fabs(xs);
    R g;
    switch (order) {
    case 2: g = x < R(0.5) ? R(-2) * x : x - R(1.5); break;
    case 3: g = x < R(1) ? x * (x * R(1.5) - R(2)) : R(-0.5) * sq(R(2) - x); break;
    case 4:
        if (x < R(0.5)) g = x * (x * x - R(1.25));
        else if (x < R(1.5)) g = x * (x * (x * R(-2. / 3.) + R(2.5)) - R(2.5)) + R(5. / 24.);
        else g = cu(R(2) * x - R(5)) * R(1. / 48.);
        break;
    case 5:
        if (x < R(1)) g = x * (x * (x * (x * R(-5. / 12.) + R(1))) - R(1));
        else if (x < R(2)) g = x * (x * (x * (x * R(5. / 24.) - R(1.5)) + R(3.75)) - R(3.5)) + R(0.625);
        else g = sq(sq(x - R(3))) * R(-1. / 24.);
        break;
    case 6:
        if (x < R(0.5)) { R y = x * x; g = x * (y * R(7. / 12.) - y * y * R(1. / 6.) - R(77. / 96.)); }
        else if (x < R(1.5)) g = x * (x * (x * (x * (x * R(0.125) - R(35. / 48.)) + R(1.3125)) - R(35. / 96.)) - R(0.7109375)) - R(7. / 768.);
        else if (x < R(2.5)) g = x * (x * (x * (x * (x * R(-1. / 20.) + R(7. / 12.)) - R(2.625)) + R(133. / 24.)) - R(5.140625)) + R(1267. / 960.);
        else { R y = R(2) * x - R(7); g = sq(sq(y)) * y * R(1. / 3840.); }
        break;
    case 7:
        if (x < R(1)) { R y = x * x; g = x * (y * (y * (x * R(7. / 144.) - R(1. / 6.)) + R(4. / 9.)) - R(2. / 3.)); }
        else if (x < R(2)) g = x * (x * (x * (x * (x * (x * R(-7. / 240.) + R(3. / 10.)) - R(7. / 6.)) + R(2)) - R(7. / 6.)) - R(1. / 5.)) - R(7. / 90.);
        else if (x < R(3)) g = x * (x * (x * (x * (x * (x * R(7. / 720.) - R(1. / 6.)) + R(7. / 6.)) - R(38. / 9.)) + R(49. / 6.)) - R(23. / 3.)) + R(217. / 90.);
        else g = sq(cu(x - R(4))) * R(-1. / 720.);
        break;
    default: return R(0);
    }
    return xs > R(0) ? g : (xs < R(0) ? -g : R(0));
}

// d2/dx2 B_order(x)
template <typename R>
__device__ __forceinline__ R spline_hess(int order, R x) {
    x = fabs(x);
    switch (order) {
    case 2: return x < R(0.5) ? R(-2) : R(1);
    case 3: return x < R(1) ? R(3) * x - R(2) : R(2) - x;
    case 4:
        if (x < R(0.5)) return R(3) * x * x - R(1.25);
        if (x < R(1.5)) return x * (R(-2) * x + R(5)) - R(2.5);
        return sq(R(2) * x - R(5)) * R(0.125);
    case 5:
        if (x < R(1)) return -(x * x) * (x * R(5. / 3.) - R(3)) - R(1);
        if (x < R(2)) return x * (x * (x * R(5. / 6.) - R(4.5)) + R(7.5)) - R(3.5);
        return R(4.5) - x * (x * (x * R(1. / 6.) - R(1.5)) + R(4.5));
    case 6:
        if (x < R(0.5)) { R y = x * x; return -y * (y * R(5. / 6.) - R(1.75)) - R(77. / 96.); }
        if (x < R(1.5)) return x * (x * (x * (x * R(0.625) - R(35. / 12.)) + R(63. / 16.)) - R(35. / 48.)) - R(91. / 128.);
        if (x < R(2.5)) return -(x * (x * (x * (x * R(0.25) - R(7. / 3.)) + R(63. / 8.)) - R(133. / 12.)) + R(329. / 64.));
        return x * (x * (x * (x * R(1. / 24.) - R(7. / 12.)) + R(49. / 16.)) - R(343. / 48.)) + R(2401. / 384.);
    case 7:
        if (x < R(1)) { R y = x * x; return y * (y * (x * R(7. / 24.) - R(5. / 6.)) + R(4. / 3.)) - R(2. / 3.); }
        if (x < R(2)) return -(x * (x * (x * (x * (x * R(7. / 40.) - R(1.5)) + R(14. / 3.)) - R(6)) + R(7. / 3.)) + R(0.2));
        if (x < R(3)) return x * (x * (x * (x * (x * R(7. / 120.) - R(5. / 6.)) + R(14. / 3.)) - R(38. / 3.)) + R(49. / 3.)) - R(23. / 3.);
        return -(x * (x * (x * (x * (x * R(1. / 120.) - R(1. / 6.)) + R(4. / 3.)) - R(16. / 3.)) + R(32. / 3.)) - R(128. / 15.));
    }
    return R(0);
}

}  // namespace ib200
