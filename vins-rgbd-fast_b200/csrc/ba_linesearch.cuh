// ba_linesearch.cuh -- step-size selection of Ceres' projected Armijo line search, used by k_ba_solve for
// bound-constrained problems (estimate_flag == 2 landmarks, reference estimator.cpp:1293-1298).
//
// Third-party algorithm restated from Ceres Solver (internal/ceres/line_search.cc
// LineSearch::InterpolatingPolynomialMinimizingStepSize with CUBIC interpolation, polynomial.cc
// FindInterpolatingPolynomial / MinimizePolynomial / FindPolynomialRoots); the CPU statement is
// oracle/ba_ref.c::ls_interpolating_step.  Every thread of the CTA evaluates these functions redundantly on identical
// (block-reduced) sample values: no shared memory, no barrier, the same result in every thread.
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

__device__ __forceinline__ double ls_ipow(double x, int k)
{
    double v = 1.0;
    for (int i = 0; i < k; ++i) v *= x;
    return v;
}

// polynomial of degree (#constraints - 1) through the samples' values and gradients (coefficients highest power first);
// Ceres solves the system with Eigen's fullPivLu: Gaussian elimination with full pivoting
__device__ __noinline__ int ls_poly_interpolate(const LsSample *smp, int ns, double *coef)
{
    int nc = 0;
    for (int i = 0; i < ns; ++i) nc += (smp[i].value_ok ? 1 : 0) + (smp[i].grad_ok ? 1 : 0);
    const int degree = nc - 1;
    double A[6][7];
    int row = 0;
    for (int i = 0; i < ns; ++i) {
        if (smp[i].value_ok) {
            for (int j = 0; j <= degree; ++j) A[row][j] = ls_ipow(smp[i].x, degree - j);
            A[row][nc] = smp[i].value; ++row;
        }
        if (smp[i].grad_ok) {
            for (int j = 0; j < degree; ++j) A[row][j] = (degree - j) * ls_ipow(smp[i].x, degree - j - 1);
            A[row][degree] = 0.0;
            A[row][nc] = smp[i].gradient; ++row;
        }
    }
    int perm[6];
    for (int j = 0; j < nc; ++j) perm[j] = j;
    for (int k = 0; k < nc; ++k) {
        int pr = k, pc = k;
        double best = -1;
        for (int r = k; r < nc; ++r)
            for (int c = k; c < nc; ++c) if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); pr = r; pc = c; }
        if (best <= 0) { for (int r = k; r < nc; ++r) A[r][nc] = 0; break; }
        if (pr != k) for (int c = 0; c <= nc; ++c) { const double t = A[k][c]; A[k][c] = A[pr][c]; A[pr][c] = t; }
        if (pc != k) {
            for (int r = 0; r < nc; ++r) { const double t = A[r][k]; A[r][k] = A[r][pc]; A[r][pc] = t; }
            const int t = perm[k]; perm[k] = perm[pc]; perm[pc] = t;
        }
        for (int r = k + 1; r < nc; ++r) {
            const double f = A[r][k] / A[k][k];
            for (int c = k; c <= nc; ++c) A[r][c] -= f * A[k][c];
        }
    }
    double y[6];
    for (int k = nc - 1; k >= 0; --k) {
        double v = A[k][nc];
        for (int c = k + 1; c < nc; ++c) v -= A[k][c] * y[c];
        y[k] = A[k][k] != 0.0 ? v / A[k][k] : 0.0;
    }
    for (int k = 0; k < nc; ++k) coef[perm[k]] = y[k];
    return nc;
}

// real parts of all roots (complex ones included: MinimizePolynomial tests them too).  Degrees 1, 2 in closed form like
// FindPolynomialRoots; higher degrees by the Aberth-Ehrlich iteration instead of the companion-matrix eigenvalues.
__device__ __noinline__ int ls_root_real_parts(const double *c_in, int n, double *re)
{
    while (n > 0 && c_in[0] == 0.0) { ++c_in; --n; }
    const int deg = n - 1;
    if (deg < 1) return 0;
    if (deg == 1) { re[0] = -c_in[1] / c_in[0]; return 1; }
    if (deg == 2) {
        const double a = c_in[0], b = c_in[1], c = c_in[2];
        const double D = b * b - 4 * a * c, sD = sqrt(fabs(D));
        if (D >= 0) {
            if (b >= 0) { re[0] = (-b - sD) / (2.0 * a); re[1] = (2.0 * c) / (-b - sD); }
            else { re[0] = (2.0 * c) / (-b + sD); re[1] = (-b + sD) / (2.0 * a); }
        } else { re[0] = -b / (2.0 * a); re[1] = -b / (2.0 * a); }
        return 2;
    }
    double a[8];
    for (int i = 0; i <= deg; ++i) a[i] = c_in[i] / c_in[0];
    double rad = 0;
    for (int i = 1; i <= deg; ++i) rad = fmax(rad, fabs(a[i]));
    rad = 1.0 + rad;
    double zr[8], zi[8];
    for (int k = 0; k < deg; ++k) { const double ang = 2.0 * 3.14159265358979323846 * k / deg + 0.4; zr[k] = 0.5 * rad * cos(ang); zi[k] = 0.5 * rad * sin(ang); }
    for (int it = 0; it < 500; ++it) {
        double change = 0;
        for (int k = 0; k < deg; ++k) {
            double pr = 1.0, pi = 0.0, dr = 0.0, di = 0.0;
            for (int i = 1; i <= deg; ++i) {
                const double ndr = dr * zr[k] - di * zi[k] + pr, ndi = dr * zi[k] + di * zr[k] + pi;
                dr = ndr; di = ndi;
                const double npr = pr * zr[k] - pi * zi[k] + a[i], npi = pr * zi[k] + pi * zr[k];
                pr = npr; pi = npi;
            }
            const double dd = dr * dr + di * di;
            if (dd == 0.0) continue;
            const double wr = (pr * dr + pi * di) / dd, wi = (pi * dr - pr * di) / dd;
            double sr = 0, si = 0;
            for (int j = 0; j < deg; ++j) {
                if (j == k) continue;
                const double er = zr[k] - zr[j], ei = zi[k] - zi[j], ee = er * er + ei * ei;
                if (ee == 0.0) continue;
                sr += er / ee; si -= ei / ee;
            }
            const double qr = 1.0 - (wr * sr - wi * si), qi = -(wr * si + wi * sr), qq = qr * qr + qi * qi;
            if (qq == 0.0) continue;
            const double ur = (wr * qr + wi * qi) / qq, ui = (wi * qr - wr * qi) / qq;
            zr[k] -= ur; zi[k] -= ui;
            change = fmax(change, fabs(ur) + fabs(ui));
        }
        if (change <= 1e-15 * rad) break;
    }
    for (int k = 0; k < deg; ++k) re[k] = zr[k];
    return deg;
}

__device__ __noinline__ double ls_poly_minimize(const double *c, int n, double x_min, double x_max)
{
    double best_x = (x_min + x_max) / 2.0, best = ls_poly_eval(c, n, best_x);
    double v = ls_poly_eval(c, n, x_min);
    if (v < best) { best = v; best_x = x_min; }
    v = ls_poly_eval(c, n, x_max);
    if (v < best) { best = v; best_x = x_max; }
    if (n <= 2) return best_x;
    double d[8], re[8];
    for (int i = 0; i < n - 1; ++i) d[i] = (n - 1 - i) * c[i];
    const int nr = ls_root_real_parts(d, n - 1, re);
    for (int i = 0; i < nr; ++i) {
        if (re[i] < x_min || re[i] > x_max) continue;
        v = ls_poly_eval(c, n, re[i]);
        if (v < best) { best = v; best_x = re[i]; }
    }
    return best_x;
}

// next step size: minimiser on [min_step, max_step] of the polynomial through (lower bound, current, previous), values
// and gradients; bisection when the current sample is not finite
__device__ __noinline__ double ls_interpolating_step(const LsSample &lower, const LsSample &prev, const LsSample &cur, double min_step,
                                                     double max_step)
{
    if (!cur.value_ok) return fmin(fmax(cur.x * 0.5, min_step), max_step);
    LsSample smp[3];
    int ns = 0;
    smp[ns++] = lower;
    smp[ns++] = cur;
    if (prev.value_ok) smp[ns++] = prev;
    double coef[6];
    const int nc = ls_poly_interpolate(smp, ns, coef);
    return ls_poly_minimize(coef, nc, min_step, max_step);
}

}  // namespace vrf
