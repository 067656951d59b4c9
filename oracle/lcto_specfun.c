/*
 * lcto_specfun.c -- ORACLE (test infrastructure): special functions and depth tables.
 *
 * Restates statrs 0.19 (not vendored in /root/reference) ln_gamma / beta_reg / StudentsT::cdf
 * from the published algorithms, plus the reference's own Ln helpers and the Bayesian
 * Negative-Binomial depth table:
 *   src/math/mod.rs:10-95           Ln::{add,sum,sum_init}
 *   src/math/distr/nbinom.rs:35-42,68-70,127-132   NBinom::{new,mul,ln_pmf}
 *   src/math/distr/bayes.rs:26-35   BayesCalc::ln_pmf
 *   src/model/distr_cache.rs:61-75  DistrCache::new
 * Pinned against scipy fixtures in tests/golden/specfun.json (close, not bit-exact: different
 * algorithms); the product never computes these on device -- tables are inputs.
 */
#include "lcto.h"
#include <math.h>
#include <float.h>

/* statrs::function::gamma::ln_gamma: Lanczos, g = 10.900511, 11 coefficients. */
static const double GAMMA_R = 10.900511;
static const double GAMMA_DK[11] = {
    2.48574089138753565546e-5,
    1.05142378581721974210,
    -3.45687097222016235469,
    4.51227709466894823700,
    -2.98285225323576655721,
    1.05639711577126713077,
    -1.95428773191645869583e-1,
    1.70970543404441224307e-2,
    -5.71926117404305781283e-4,
    4.63399473359905636708e-6,
    -2.71994908488607703910e-9,
};
static const double LN_PI = 1.1447298858494001741434273513530587116472948129153;
static const double LN_2_SQRT_E_OVER_PI = 0.6207822376352452223455184457816472122518527279025978;

double lcto_ln_gamma(double x) {
    if (x < 0.5) {
        double s = GAMMA_DK[0];
        for (int i = 1; i < 11; i++) s += GAMMA_DK[i] / ((double)i - x);
        return LN_PI - log(sin(M_PI * x)) - log(s) - LN_2_SQRT_E_OVER_PI
            - (0.5 - x) * log((0.5 - x + GAMMA_R) / M_E);
    } else {
        double s = GAMMA_DK[0];
        for (int i = 1; i < 11; i++) s += GAMMA_DK[i] / (x + (double)i - 1.0);
        return log(s) + LN_2_SQRT_E_OVER_PI + (x - 0.5) * log((x - 0.5 + GAMMA_R) / M_E);
    }
}

/* statrs::function::beta::beta_reg: continued fraction (Numerical Recipes betacf), 140 terms. */
double lcto_beta_reg(double a, double b, double x) {
    if (!(a > 0.0) || !(b > 0.0) || !(x >= 0.0 && x <= 1.0)) return NAN;
    double bt;
    if (fabs(x) < 1.1102230246251565e-15 || fabs(x - 1.0) <= 4.0 * DBL_EPSILON) {
        bt = 0.0;
    } else {
        bt = exp(lcto_ln_gamma(a + b) - lcto_ln_gamma(a) - lcto_ln_gamma(b) + a * log(x) + b * log(1.0 - x));
    }
    int symm = x >= (a + 1.0) / (a + b + 2.0);
    const double eps = 1.1102230246251565e-16;
    const double fpmin = DBL_MIN / eps;
    if (symm) { double sw = a; x = 1.0 - x; a = b; b = sw; }
    double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0;
    double d = 1.0 - qab * x / qap;
    if (fabs(d) < fpmin) d = fpmin;
    d = 1.0 / d;
    double h = d;
    for (int mi = 1; mi < 141; mi++) {
        double m = (double)mi;
        double m2 = m * 2.0;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < fpmin) d = fpmin;
        c = 1.0 + aa / c;
        if (fabs(c) < fpmin) c = fpmin;
        d = 1.0 / d;
        h = h * d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d;
        if (fabs(d) < fpmin) d = fpmin;
        c = 1.0 + aa / c;
        if (fabs(c) < fpmin) c = fpmin;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (fabs(del - 1.0) <= eps) break;
    }
    return symm ? 1.0 - bt * h / a : bt * h / a;
}

/* statrs StudentsT::new(0, 1, freedom).cdf(x); src/math/mod.rs:195,218 */
double lcto_students_t_cdf(double x, double freedom) {
    if (isinf(freedom)) return 0.5 * erfc(-x / M_SQRT2);
    double h = freedom / (freedom + x * x);
    double ib = 0.5 * lcto_beta_reg(freedom / 2.0, 0.5, h);
    return x <= 0.0 ? ib : 1.0 - ib;
}

/* src/math/mod.rs:28-34 */
double lcto_ln_add(double a, double b) {
    if (a >= b) {
        return (b == -INFINITY) ? a : b + log1p(exp(a - b));
    } else {
        return (a == -INFINITY) ? b : a + log1p(exp(b - a));
    }
}

/* src/math/mod.rs:50-75 (Ln::sum = map_sum with identity) */
double lcto_ln_sum(const double *v, size_t n) {
    if (n == 0) return -INFINITY;
    if (n == 1) return v[0];
    double m = -INFINITY;
    for (size_t i = 0; i < n; i++) m = fmax(m, v[i]);
    if (isinf(m)) return m;
    double s = 0.0;
    for (size_t i = 0; i < n; i++) s = s + exp(v[i] - m);
    return m + log(s);
}

/* src/math/mod.rs:56-94 (Ln::sum_init = map_sum_init with identity) */
double lcto_ln_sum_init(const double *v, size_t n, double init) {
    if (n == 0) return init;
    if (n == 1) return lcto_ln_add(init, v[0]);
    double m = init;
    for (size_t i = 0; i < n; i++) m = fmax(m, v[i]);
    if (isinf(m)) return m;
    double s = exp(init - m);
    for (size_t i = 0; i < n; i++) s = s + exp(v[i] - m);
    return m + log(s);
}

typedef struct { double n, p, lnq, lnpmf_const; } nbinom;

/* src/math/distr/nbinom.rs:35-42 */
static nbinom nbinom_new(double n, double p) {
    nbinom d;
    d.n = n; d.p = p;
    d.lnq = log1p(-p);
    d.lnpmf_const = n * log(p) - lcto_ln_gamma(n);
    return d;
}

/* src/math/distr/nbinom.rs:127-132 */
static double nbinom_ln_pmf(const nbinom *d, uint32_t k) {
    double x = (double)k;
    return d->lnpmf_const + lcto_ln_gamma(d->n + x) - lcto_ln_gamma(x + 1.0) + x * d->lnq;
}

/* src/model/distr_cache.rs:61-75 with BayesCalc::ln_pmf (src/math/distr/bayes.rs:26-35).
 * The reference caches k < 256 lazily and recomputes k >= 256 with the same function, so one
 * dense table with k_cols columns reproduces every value it can return. */
void lcto_build_depth_table(const double *nb_n, const double *nb_p, int is_paired,
                            const double *alt_cn, size_t n_alt, uint32_t k_cols, double *out) {
    double mul_coef = is_paired ? 2.0 : 1.0;
    for (int gc = 0; gc < LCTO_GC_BINS; gc++) {
        nbinom cn1 = nbinom_new(nb_n[gc] * mul_coef, nb_p[gc]);     /* .mul(mul_coef), nbinom.rs:68-70 */
        nbinom alts[16];
        for (size_t a = 0; a < n_alt && a < 16; a++) alts[a] = nbinom_new(cn1.n * alt_cn[a], cn1.p);
        for (uint32_t k = 0; k < k_cols; k++) {
            double null_prob = nbinom_ln_pmf(&cn1, k);
            double probs[16];
            for (size_t a = 0; a < n_alt && a < 16; a++) probs[a] = nbinom_ln_pmf(&alts[a], k);
            double sum_prob = lcto_ln_sum_init(probs, n_alt, null_prob);
            out[(size_t)gc * k_cols + k] = null_prob - sum_prob;
        }
    }
}
