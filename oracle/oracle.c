/*
 * CPU oracle (plain C restatement) of ramannoodle's MD-Raman hot path.
 *
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library, and only as the
 * checker or as the timed CPU baseline.  The product library (libramannoodle_b200.so) does
 * not link, load or call it.
 *
 * Citations are file:line relative to /root/reference (ramannoodle v0.5.0).  The B-spline
 * evaluator restates scipy's published de Boor routine (scipy.interpolate.BSpline.__call__
 * -> _dierckx evaluate_spline / find_interval / _deBoor_D; scipy is an un-vendored,
 * only lower-bounded dependency of the reference: pyproject.toml:16-27; installed here:
 * scipy 1.18.1) and is pinned bit-for-bit against scipy in tests/test_oracle_c.py.
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC oracle.c -o _build/liboracle.so -lm
 * (-ffp-contract=off keeps every multiply and add separately rounded, like numpy/scipy).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_DEGREE 15

/* numpy's float `x % 1` (npy_divmod with b = 1): fmod, then shift negatives up by 1. */
static double np_mod1(double x) {
    double m = fmod(x, 1.0);
    if (m != 0.0) {
        if (m < 0.0) m += 1.0;
    } else {
        m = 0.0; /* copysign(0, b) with b = +1 */
    }
    return m;
}

/* numpy's float `x // 1` is floor(x) for finite x (npy_floor_divide). */
static double np_floordiv1(double x) { return floor(x); }

/* ramannoodle/structure/utils.py:13-29  apply_pbc: positions - positions // 1 */
static double apply_pbc(double p) { return p - np_floordiv1(p); }

/* ramannoodle/structure/utils.py:32-48  where(d % 1 > 0.5, d % 1 - 1, d % 1) */
static double apply_pbc_displacement(double d) {
    double m = np_mod1(d);
    return (m > 0.5) ? (m - 1.0) : m;
}

void orc_apply_pbc(const double *in, double *out, int64_t n) {
    for (int64_t i = 0; i < n; i++) out[i] = apply_pbc(in[i]);
}

void orc_apply_pbc_displacement(const double *in, double *out, int64_t n) {
    for (int64_t i = 0; i < n; i++) out[i] = apply_pbc_displacement(in[i]);
}

/*
 * Cartesian displacements of S frames: ramannoodle/pmodel/_interpolation.py:217-223
 *   calc_displacement (structure/utils.py:110-135) -> get_cart_displacement
 *   (structure/_reference.py:268-285: wrap again, then `@ lattice`, lattice rows = vectors).
 * out is (S, 3N) row-major.
 */
void orc_cart_displacements(const double *ref_pos, const double *lattice, const double *pos,
                            int64_t S, int64_t N, double *out) {
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < S; s++) {
        for (int64_t a = 0; a < N; a++) {
            double w[3];
            for (int c = 0; c < 3; c++) {
                double p1 = apply_pbc(ref_pos[a * 3 + c]);
                double p2 = apply_pbc(pos[(s * N + a) * 3 + c]);
                double d = apply_pbc_displacement(p2 - p1);
                w[c] = apply_pbc_displacement(d);
            }
            for (int c = 0; c < 3; c++) {
                out[(s * N + a) * 3 + c] =
                    w[0] * lattice[0 * 3 + c] + w[1] * lattice[1 * 3 + c] + w[2] * lattice[2 * 3 + c];
            }
        }
    }
}

/* scipy find_interval(t, k, xval, prev_l, extrapolate=True): index l with
 * t[l] <= x < t[l+1], clamped to [k, n-1]; returns -1 for NaN. */
static int find_interval(const double *t, int nt, int k, double x, int prev_l) {
    int n = nt - k - 1;
    if (x != x) return -1;
    int l = (k < prev_l && prev_l < n) ? prev_l : k;
    while (x < t[l] && l != k) l--;
    l++;
    while (x >= t[l] && l != n) l++;
    return l - 1;
}

/* scipy _deBoor_D(t, x, k, ell, m=0, result): the k+1 non-zero B-splines at x. */
static void deboor(const double *t, double x, int k, int ell, double *result) {
    double *hh = result + k + 1;
    double *h = result;
    h[0] = 1.0;
    for (int j = 1; j <= k; j++) {
        memcpy(hh, h, (size_t)j * sizeof(double));
        h[0] = 0.0;
        for (int n = 1; n <= j; n++) {
            int ind = ell + n;
            double xb = t[ind];
            double xa = t[ind - j];
            if (xb == xa) {
                h[n] = 0.0;
                continue;
            }
            double w = hh[n - 1] / (xb - xa);
            h[n - 1] += w * (xb - x);
            h[n] = w * (x - xa);
        }
    }
}

/*
 * BSpline(t, c, k, extrapolate=True)(x) with vector-valued coefficients c (n, m) row-major;
 * out is (nx, m).  This is the call at ramannoodle/pmodel/_interpolation.py:243.
 * Returns 0 on success, -1 if k is too large for the work buffer.
 */
int orc_eval_bspline(const double *t, int nt, const double *c, int m, int k, const double *x,
                     int64_t nx, double *out) {
    if (k > ORC_MAX_DEGREE || k < 0) return -1;
    double work[2 * ORC_MAX_DEGREE + 2];
    int interval = k;
    for (int64_t ip = 0; ip < nx; ip++) {
        double xv = x[ip];
        interval = find_interval(t, nt, k, xv, interval);
        if (interval < 0) {
            for (int jp = 0; jp < m; jp++) out[ip * m + jp] = NAN;
            interval = k;
            continue;
        }
        deboor(t, xv, k, interval, work);
        for (int jp = 0; jp < m; jp++) {
            double acc = 0.0;
            for (int a = 0; a <= k; a++) acc = acc + c[(int64_t)(interval + a - k) * m + jp] * work[a];
            out[ip * m + jp] = acc;
        }
    }
    return 0;
}

/*
 * get_polarizability: ramannoodle/pmodel/_interpolation.py:233-252 from precomputed (S,3N)
 * Cartesian displacements.  Ragged spline tables: DOF j has degree k[j], knots
 * t[t_off[j] .. t_off[j+1]) and coefficients c[9*c_off[j] .. 9*c_off[j+1]) as (n_j, 9).
 * weight[j] = 1 - mask[j].  alpha is (S, 9).
 */
int orc_get_polarizability(const double *cart, int64_t S, int64_t K3N, const double *basis,
                           int64_t J, const int32_t *k, const int64_t *t_off, const double *t,
                           const int64_t *c_off, const double *c, const double *weight,
                           const double *ref_pol, double *alpha) {
    for (int64_t j = 0; j < J; j++)
        if (k[j] > ORC_MAX_DEGREE || k[j] < 0) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < S; s++) {
        double delta[9] = {0};
        const double *d = cart + s * K3N;
        for (int64_t j = 0; j < J; j++) {
            const double *v = basis + j * K3N;
            double amp = 0.0;
            for (int64_t i = 0; i < K3N; i++) amp += v[i] * d[i]; /* einsum("i,ji") :239-241 */
            double val[9];
            orc_eval_bspline(t + t_off[j], (int)(t_off[j + 1] - t_off[j]), c + 9 * c_off[j], 9,
                             k[j], &amp, 1, val);
            for (int q = 0; q < 9; q++) delta[q] += weight[j] * val[q]; /* :242-244 */
        }
        for (int q = 0; q < 9; q++) alpha[s * 9 + q] = delta[q] + ref_pol[q]; /* :252 */
    }
    return 0;
}

/* calc_polarizabilities: ramannoodle/pmodel/_interpolation.py:191-252 (fractional positions). */
int orc_calc_polarizabilities(const double *ref_pos, const double *lattice, int64_t N,
                              const double *pos, int64_t S, const double *basis, int64_t J,
                              const int32_t *k, const int64_t *t_off, const double *t,
                              const int64_t *c_off, const double *c, const double *weight,
                              const double *ref_pol, double *alpha) {
    double *cart = (double *)malloc(sizeof(double) * (size_t)S * (size_t)N * 3);
    if (!cart) return -2;
    orc_cart_displacements(ref_pos, lattice, pos, S, N, cart);
    int rc = orc_get_polarizability(cart, S, 3 * N, basis, J, k, t_off, t, c_off, c, weight,
                                    ref_pol, alpha);
    free(cart);
    return rc;
}

/*
 * calc_signal_spectrum by its definition (ramannoodle/spectrum/utils.py:76-124), O(M^2):
 * ac[tau] = sum_t x[t] x[t+tau] (positive lags of correlate(x, x, "full")), then
 * I[k] = Re sum_tau ac[tau] exp(-2 pi i k tau / M) for the first ceil(M/2) bins, and
 * wn[k] = k / (M dt) * 33.35640951981521 * 1e3.  Small M only.
 */
void orc_signal_spectrum_direct(const double *x, int64_t M, double dt, double *wn, double *inten) {
    double *ac = (double *)malloc(sizeof(double) * (size_t)M);
    for (int64_t tau = 0; tau < M; tau++) {
        long double acc = 0.0L;
        for (int64_t i = 0; i + tau < M; i++) acc += (long double)x[i] * (long double)x[i + tau];
        ac[tau] = (double)acc;
    }
    int64_t nk = (M + 1) / 2;
#pragma omp parallel for schedule(static)
    for (int64_t kk = 0; kk < nk; kk++) {
        long double acc = 0.0L;
        for (int64_t tau = 0; tau < M; tau++) {
            int64_t ph = (kk * tau) % M;
            acc += (long double)ac[tau] * cosl(2.0L * 3.14159265358979323846264338327950288L * (long double)ph / (long double)M);
        }
        inten[kk] = (double)acc;
        wn[kk] = ((double)kk / ((double)M * dt)) * 33.35640951981521 * 1e3;
    }
    free(ac);
}

/* convolve_spectrum: ramannoodle/spectrum/utils.py:57-72.  kind 0 = gaussian, 1 = lorentzian. */
int orc_convolve_spectrum(const double *wn, const double *inten, int64_t K, int kind, double width,
                          const double *out_wn, int64_t L, double *out_inten) {
    if (kind != 0 && kind != 1) return -1;
    const double pi = 3.141592653589793;
    for (int64_t l = 0; l < L; l++) out_inten[l] = out_wn[l] * 0;
    for (int64_t i = 0; i < K; i++) {
#pragma omp parallel for schedule(static)
        for (int64_t l = 0; l < L; l++) {
            double dx = wn[i] - out_wn[l];
            double factor;
            if (kind == 0)
                factor = (1 / width) * (1 / sqrt(2 * pi)) * exp(-(dx * dx) / (2 * (width * width)));
            else
                factor = (1 / pi) * (0.5 * width / (dx * dx + (0.5 * width) * (0.5 * width)));
            out_inten[l] += factor * inten[i];
        }
    }
    return 0;
}
