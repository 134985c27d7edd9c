// ba_linesearch.cuh -- step-size selection of Ceres' projected Armijo line search, used by k_ba_solve for
// bound-constrained problems (estimate_flag == 2 landmarks, reference estimator.cpp:1293-1298).
//
// Third-party algorithm restated from Ceres Solver (internal/ceres/line_search.cc
// LineSearch::InterpolatingPolynomialMinimizingStepSize with CUBIC interpolation, polynomial.cc
// FindInterpolatingPolynomial / MinimizePolynomial); the CPU statement is
// oracle/ba_ref.c::ls_interpolating_step.  The samples are block-reduced values, identical in every thread; k_ba_solve lets
// one warp evaluate these functions (all of its lanes redundantly) and broadcasts the step through shared memory.
#pragma once
#include <math.h>

namespace vrf {

struct LsSample { double x, value, gradient; bool value_ok, grad_ok; };

__device__ __forceinline__ double ls_poly_eval(const double *c, int n, double x)
{
    double v = 0;
    for (int i = 0; i < n; ++i) v = v * x + c[i];
    return v;
}

// value and derivative of a polynomial (n coefficients, highest power first)
__device__ __forceinline__ void ls_poly_eval2(const double *c, int n, double x, double *f, double *df)
{
    double v = c[0], d = 0.0;
    for (int i = 1; i < n; ++i) { d = d * x + v; v = v * x + c[i]; }
    *f = v; *df = d;
}

// root of c inside the bracket [u, v] (sign change, fu = c(u)): Newton steps, bisection whenever Newton leaves the bracket or
// stops halving the step (Numerical Recipes' rtsafe), down to a few ulp
__device__ __forceinline__ double ls_root_bracketed(const double *c, int n, double u, double v, double fu)
{
    double xl = fu < 0 ? u : v, xh = fu < 0 ? v : u;          // c(xl) < 0 < c(xh)
    double x = 0.5 * (u + v), dxold = fabs(v - u), dx = dxold, f, df;
    ls_poly_eval2(c, n, x, &f, &df);
    for (int it = 0; it < 200; ++it) {
        if (((x - xh) * df - f) * ((x - xl) * df - f) > 0.0 || fabs(2.0 * f) > fabs(dxold * df)) {
            dxold = dx; dx = 0.5 * (xh - xl);
            const double xn = xl + dx;
            if (xn == xl) return xn;
            x = xn;
        } else {
            dxold = dx; dx = f / df;
            const double xn = x - dx;
            if (xn == x) return x;
            x = xn;
        }
        if (fabs(dx) <= 4.4e-16 * fabs(x)) return x;
        ls_poly_eval2(c, n, x, &f, &df);
        if (f == 0.0) return x;
        if (f < 0) xl = x; else xh = x;
    }
    return x;
}

// real roots of c (degree <= 2 after stripping leading zeros) inside [a, b], ascending
__device__ __forceinline__ int ls_roots_low(const double *c, int n, double a, double b, double *roots)
{
    while (n > 0 && c[0] == 0.0) { ++c; --n; }
    const int deg = n - 1;
    int nr = 0;
    if (deg < 1) return 0;
    if (deg == 1) { const double r = -c[1] / c[0]; if (r >= a && r <= b) roots[nr++] = r; return nr; }
    const double D_ = c[1] * c[1] - 4 * c[0] * c[2];
    if (D_ < 0) return 0;
    const double sD = sqrt(D_);
    double r0, r1;
    if (c[1] >= 0) { r0 = (-c[1] - sD) / (2.0 * c[0]); r1 = (2.0 * c[2]) / (-c[1] - sD); }
    else { r0 = (2.0 * c[2]) / (-c[1] + sD); r1 = (-c[1] + sD) / (2.0 * c[0]); }
    if (r0 > r1) { const double t = r0; r0 = r1; r1 = t; }
    if (r0 >= a && r0 <= b) roots[nr++] = r0;
    if (r1 >= a && r1 <= b && r1 != r0) roots[nr++] = r1;
    return nr;
}

// real roots inside [a, b] of c given the real roots `crit` of its derivative: they cut [a, b] into monotone pieces
__device__ __forceinline__ int ls_roots_from_crit(const double *c, int n, double a, double b, const double *crit, int ncrit, double *roots)
{
    int nr = 0;
    double u = a, fu = ls_poly_eval(c, n, a);
    for (int k = 0; k <= ncrit; ++k) {
        const double v = k < ncrit ? crit[k] : b;
        if (!(v > u)) continue;
        const double fv = ls_poly_eval(c, n, v);
        double r = 0;
        bool have = true;
        if (fu == 0.0) r = u;
        else if (fv == 0.0) r = v;
        else if ((fu < 0) != (fv < 0)) r = ls_root_bracketed(c, n, u, v, fu);
        else have = false;
        if (have && (nr == 0 || r != roots[nr - 1])) roots[nr++] = r;
        u = v; fu = fv;
    }
    return nr;
}

// real roots inside [a, b] of a polynomial of degree <= 4, ascending (degree 2 in closed form, 3 and 4 through the derivative)
__device__ __noinline__ int ls_roots_in(const double *c, int n, double a, double b, double *roots)
{
    while (n > 0 && c[0] == 0.0) { ++c; --n; }
    if (n - 1 <= 2) return ls_roots_low(c, n, a, b, roots);
    double d1[4], d2[3], crit2[4], crit1[4];
    for (int i = 0; i < n - 1; ++i) d1[i] = (n - 1 - i) * c[i];
    int n1;
    if (n - 1 == 3) n1 = ls_roots_low(d1, 3, a, b, crit1);
    else {
        for (int i = 0; i < 3; ++i) d2[i] = (3 - i) * d1[i];
        const int n2 = ls_roots_low(d2, 3, a, b, crit2);
        n1 = ls_roots_from_crit(d1, 4, a, b, crit2, n2, crit1);
    }
    return ls_roots_from_crit(c, n, a, b, crit1, n1, roots);
}

/* LineSearch::InterpolatingPolynomialMinimizingStepSize for CUBIC interpolation: the minimiser over [min_step, max_step] of the
 * polynomial through value and gradient of (lower bound at 0, current[, previous]).  FindInterpolatingPolynomial solves the
 * Vandermonde-type system of all (4 or 6) constraints with Eigen's fullPivLu; the same polynomial is obtained here in the
 * scaled variable u = x / current.x, q(u) = f0 + g0 xc u + b2 u^2 + ... (the two constraints at 0 fix the low coefficients),
 * from a 2 x 2 / 4 x 4 system with entries of order one.  MinimizePolynomial samples the midpoint, the ends and the real part
 * of EVERY root of the derivative (companion-matrix eigenvalues, complex ones included, "a bit of an overkill") inside the
 * interval; the minimum over an interval sits at an end or at a real critical point, so only the real roots inside the
 * interval can win, and those are found exactly (bracketed through the derivative's roots, polished by safeguarded Newton). */
__device__ __noinline__ double ls_interpolating_step(const LsSample &lower, const LsSample &prev, const LsSample &cur, double min_step,
                                                     double max_step)
{
    if (!cur.value_ok) return fmin(fmax(cur.x * 0.5, min_step), max_step);
    const double xc = cur.x, f0 = lower.value, g0 = lower.gradient * xc;
    double q[6];
    int nq;
    const double rv = cur.value - f0 - g0, rg = cur.gradient * xc - g0;
    if (!prev.value_ok) {
        /* cubic: b3 + b2 = rv, 3 b3 + 2 b2 = rg */
        const double b3 = rg - 2.0 * rv, b2 = 3.0 * rv - rg;
        q[0] = b3; q[1] = b2; q[2] = g0; q[3] = f0; nq = 4;
    } else {
        const double u = prev.x / xc, u2 = u * u, u3 = u2 * u, u4 = u3 * u, u5 = u4 * u;
        double A[4][5] = {{1, 1, 1, 1, rv}, {5, 4, 3, 2, rg},
                          {u5, u4, u3, u2, prev.value - f0 - g0 * u}, {5 * u4, 4 * u3, 3 * u2, 2 * u, prev.gradient * xc - g0}};
        for (int k = 0; k < 4; k++) {                      /* Gaussian elimination, partial pivoting */
            int pr = k;
            for (int r = k + 1; r < 4; r++) if (fabs(A[r][k]) > fabs(A[pr][k])) pr = r;
            if (pr != k) for (int c = 0; c < 5; c++) { const double t = A[k][c]; A[k][c] = A[pr][c]; A[pr][c] = t; }
            if (A[k][k] == 0.0) continue;
            for (int r = k + 1; r < 4; r++) {
                const double f = A[r][k] / A[k][k];
                for (int c = k; c < 5; c++) A[r][c] -= f * A[k][c];
            }
        }
        double b[4];
        for (int k = 3; k >= 0; k--) {
            double v = A[k][4];
            for (int c = k + 1; c < 4; c++) v -= A[k][c] * b[c];
            b[k] = A[k][k] != 0.0 ? v / A[k][k] : 0.0;
        }
        q[0] = b[0]; q[1] = b[1]; q[2] = b[2]; q[3] = b[3]; q[4] = g0; q[5] = f0; nq = 6;
    }
    /* MinimizePolynomial over [min_step, max_step] in u */
    const double ua = min_step / xc, ub = max_step / xc;
    double best_x = (min_step + max_step) / 2.0, best = ls_poly_eval(q, nq, (ua + ub) / 2.0);
    double v = ls_poly_eval(q, nq, ua);
    if (v < best) { best = v; best_x = min_step; }
    v = ls_poly_eval(q, nq, ub);
    if (v < best) { best = v; best_x = max_step; }
    double d[5], roots[4];
    for (int i = 0; i < nq - 1; i++) d[i] = (nq - 1 - i) * q[i];
    const int nr = ls_roots_in(d, nq - 1, ua, ub, roots);
    for (int i = 0; i < nr; i++) {
        v = ls_poly_eval(q, nq, roots[i]);
        if (v < best) { best = v; best_x = roots[i] * xc; }
    }
    return best_x;
}

}  // namespace vrf
