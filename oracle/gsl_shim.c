/*
 * gsl_shim.c — the 26 GSL 1.16 symbols the reference sources reference, so that they link here.
 *
 * TEST INFRASTRUCTURE ONLY (see garlic_oracle.c).  /root/reference ships GSL's headers
 * (include/gsl/) but not lib/linux/libgsl.a (.MISSING_LARGE_BLOBS:5), so oracle/_ref/ref_driver
 * — the reference's own src/*.cpp compiled where they lie — links against this file instead.
 * Each function is written from the published definition of the GSL routine (GSL 1.16 manual /
 * the usual textbook algorithm), not from GSL sources; the conventions that matter for visible
 * outputs are the ones SURVEY.md Appendix A verified against the reference binary.
 * ref_driver only exercises the RNG (gsl_ran_choose for --ld-subsample); the statistics / root
 * routines are here so that garlic-kde.o, gmm.o and BoundFinder.o resolve.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <gsl/gsl_fit.h>
#include <gsl/gsl_math.h>
#include <gsl/gsl_randist.h>
#include <gsl/gsl_rng.h>
#include <gsl/gsl_roots.h>
#include <gsl/gsl_sf_log.h>
#include <gsl/gsl_sort.h>
#include <gsl/gsl_statistics.h>

/* ---- RNG: MT19937 (Matsumoto & Nishimura 1998, 2002 initialisation), gsl_rng_default ------ */
#define MT_N 624
#define MT_M 397
typedef struct { unsigned long mt[MT_N]; int mti; } mt_state_t;

static void mt_set(void *vstate, unsigned long int s)
{
    mt_state_t *st = (mt_state_t *)vstate;
    if (s == 0) s = 4357;
    st->mt[0] = s & 0xffffffffUL;
    for (int i = 1; i < MT_N; i++)
        st->mt[i] = (1812433253UL * (st->mt[i - 1] ^ (st->mt[i - 1] >> 30)) + (unsigned long)i) & 0xffffffffUL;
    st->mti = MT_N;
}

static unsigned long int mt_get(void *vstate)
{
    mt_state_t *st = (mt_state_t *)vstate;
    unsigned long *mt = st->mt;
    if (st->mti >= MT_N) {
        int kk;
        for (kk = 0; kk < MT_N - MT_M; kk++) {
            unsigned long y = (mt[kk] & 0x80000000UL) | (mt[kk + 1] & 0x7fffffffUL);
            mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
        }
        for (; kk < MT_N - 1; kk++) {
            unsigned long y = (mt[kk] & 0x80000000UL) | (mt[kk + 1] & 0x7fffffffUL);
            mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
        }
        {
            unsigned long y = (mt[MT_N - 1] & 0x80000000UL) | (mt[0] & 0x7fffffffUL);
            mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ ((y & 1UL) ? 0x9908b0dfUL : 0UL);
        }
        st->mti = 0;
    }
    unsigned long k = mt[st->mti++];
    k ^= (k >> 11);
    k ^= (k << 7) & 0x9d2c5680UL;
    k ^= (k << 15) & 0xefc60000UL;
    k ^= (k >> 18);
    return k & 0xffffffffUL;
}

static double mt_get_double(void *vstate) { return mt_get(vstate) / 4294967296.0; }

static const gsl_rng_type mt_type = {"mt19937", 0xffffffffUL, 0, sizeof(mt_state_t), &mt_set, &mt_get, &mt_get_double};
const gsl_rng_type *gsl_rng_mt19937 = &mt_type;
const gsl_rng_type *gsl_rng_default = &mt_type;
unsigned long int gsl_rng_default_seed = 0;

gsl_rng *gsl_rng_alloc(const gsl_rng_type *T)
{
    gsl_rng *r = (gsl_rng *)malloc(sizeof(gsl_rng));
    r->state = calloc(1, T->size);
    r->type = T;
    T->set(r->state, gsl_rng_default_seed);
    return r;
}
void gsl_rng_set(const gsl_rng *r, unsigned long int seed) { r->type->set(r->state, seed); }
void gsl_rng_free(gsl_rng *r) { if (r) { free(r->state); free(r); } }
double gsl_rng_uniform(const gsl_rng *r) { return r->type->get_double(r->state); }

/* selection sampling (Knuth 3.4.2 Algorithm S): keeps source order */
int gsl_ran_choose(const gsl_rng *r, void *dest, size_t k, void *src, size_t n, size_t size)
{
    size_t i, j = 0;
    if (k > n) return GSL_EINVAL;
    for (i = 0; i < n && j < k; i++) {
        if ((n - i) * gsl_rng_uniform(r) < k - j) {
            memcpy((char *)dest + size * j, (char *)src + size * i, size);
            j++;
        }
    }
    return GSL_SUCCESS;
}

double gsl_ran_gaussian_pdf(const double x, const double sigma)
{
    double u = x / fabs(sigma);
    return (1 / (sqrt(2 * M_PI) * fabs(sigma))) * exp(-u * u / 2);
}

/* ---- statistics: running-mean forms in long double -------------------- */
double gsl_stats_mean(const double data[], const size_t stride, const size_t n)
{
    long double mean = 0;
    for (size_t i = 0; i < n; i++) mean += (data[i * stride] - mean) / (i + 1);
    return (double)mean;
}

static double stats_variance_m(const double data[], size_t stride, size_t n, double mean)
{
    long double variance = 0;
    for (size_t i = 0; i < n; i++) {
        const long double delta = (data[i * stride] - mean);
        variance += (delta * delta - variance) / (i + 1);
    }
    return (double)variance;
}

double gsl_stats_variance(const double data[], const size_t stride, const size_t n)
{
    const double mean = gsl_stats_mean(data, stride, n);
    return stats_variance_m(data, stride, n, mean) * ((double)n / (double)(n - 1));
}

double gsl_stats_sd(const double data[], const size_t stride, const size_t n)
{
    const double mean = gsl_stats_mean(data, stride, n);
    return sqrt(stats_variance_m(data, stride, n, mean) * ((double)n / (double)(n - 1)));
}

void gsl_stats_minmax(double *min_out, double *max_out, const double data[], const size_t stride, const size_t n)
{
    double mn = data[0], mx = data[0];
    for (size_t i = 0; i < n; i++) {
        const double x = data[i * stride];
        if (x < mn) mn = x;
        if (x > mx) mx = x;
        if (isnan(x)) { mn = x; mx = x; break; }
    }
    *min_out = mn;
    *max_out = mx;
}

double gsl_stats_quantile_from_sorted_data(const double sorted_data[], const size_t stride, const size_t n, const double f)
{
    const double index = f * (n - 1);
    const size_t lhs = (size_t)index;
    const double delta = index - lhs;
    if (n == 0) return 0.0;
    if (lhs == n - 1) return sorted_data[lhs * stride];
    return (1 - delta) * sorted_data[lhs * stride] + delta * sorted_data[(lhs + 1) * stride];
}

/* ---- sorting ------------------------------------------------------------------------------- */
static int cmp_double(const void *a, const void *b)
{
    const double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}
void gsl_sort(double *data, const size_t stride, const size_t n)
{
    if (stride == 1) { qsort(data, n, sizeof(double), cmp_double); return; }
    double *tmp = (double *)malloc(n * sizeof(double));
    for (size_t i = 0; i < n; i++) tmp[i] = data[i * stride];
    qsort(tmp, n, sizeof(double), cmp_double);
    for (size_t i = 0; i < n; i++) data[i * stride] = tmp[i];
    free(tmp);
}

typedef struct { double v; size_t i; } idx_t;
static int cmp_idx(const void *a, const void *b)
{
    const idx_t *x = (const idx_t *)a, *y = (const idx_t *)b;
    if (x->v != y->v) return (x->v > y->v) - (x->v < y->v);
    return (x->i > y->i) - (x->i < y->i);
}
void gsl_sort_index(size_t *p, const double *data, const size_t stride, const size_t n)
{
    idx_t *t = (idx_t *)malloc((n ? n : 1) * sizeof(idx_t));
    for (size_t i = 0; i < n; i++) { t[i].v = data[i * stride]; t[i].i = i; }
    qsort(t, n, sizeof(idx_t), cmp_idx);
    for (size_t i = 0; i < n; i++) p[i] = t[i].i;
    free(t);
}

/* ---- least squares y = c0 + c1 x (running-mean accumulation) ------------------------------- */
int gsl_fit_linear(const double *x, const size_t xstride, const double *y, const size_t ystride, const size_t n,
                   double *c0, double *c1, double *cov_00, double *cov_01, double *cov_11, double *sumsq)
{
    double m_x = 0, m_y = 0, m_dx2 = 0, m_dxdy = 0;
    size_t i;
    for (i = 0; i < n; i++) {
        m_x += (x[i * xstride] - m_x) / (i + 1.0);
        m_y += (y[i * ystride] - m_y) / (i + 1.0);
    }
    for (i = 0; i < n; i++) {
        const double dx = x[i * xstride] - m_x;
        const double dy = y[i * ystride] - m_y;
        m_dx2 += (dx * dx - m_dx2) / (i + 1.0);
        m_dxdy += (dx * dy - m_dxdy) / (i + 1.0);
    }
    {
        double s2 = 0, d2 = 0;
        const double b = m_dxdy / m_dx2;
        const double a = m_y - m_x * b;
        *c0 = a;
        *c1 = b;
        for (i = 0; i < n; i++) {
            const double dx = x[i * xstride] - m_x;
            const double dy = y[i * ystride] - m_y;
            const double d = dy - b * dx;
            d2 += d * d;
        }
        s2 = d2 / (n - 2.0);
        *cov_00 = s2 * (1.0 / n) * (1 + m_x * m_x / m_dx2);
        *cov_11 = s2 * 1.0 / (n * m_dx2);
        *cov_01 = s2 * (-m_x) / (n * m_dx2);
        *sumsq = d2;
    }
    return GSL_SUCCESS;
}

/* ---- special functions ---------------------------------------------------------------------- */
double gsl_sf_log(const double x)
{
    if (x <= 0.0) {   /* GSL's default error handler aborts (observed with the reference binary) */
        fprintf(stderr, "gsl: log.c: ERROR: domain error\nDefault GSL error handler invoked.\n");
        abort();
    }
    return log(x);
}
double gsl_pow_2(const double x) { return x * x; }

/* ---- root bracketing: Brent's method (Brent 1973, ch. 4) ------------------------------------ */
typedef struct { double a, b, c, d, e, fa, fb, fc; } brent_state_t;

static int brent_init(void *vstate, gsl_function *f, double *root, double x_lower, double x_upper)
{
    brent_state_t *s = (brent_state_t *)vstate;
    *root = 0.5 * (x_lower + x_upper);
    const double f_lower = GSL_FN_EVAL(f, x_lower), f_upper = GSL_FN_EVAL(f, x_upper);
    s->a = x_lower; s->fa = f_lower;
    s->b = x_upper; s->fb = f_upper;
    s->c = x_upper; s->fc = f_upper;
    s->d = x_upper - x_lower;
    s->e = x_upper - x_lower;
    if ((f_lower < 0.0 && f_upper < 0.0) || (f_lower > 0.0 && f_upper > 0.0)) {
        fprintf(stderr, "gsl: brent.c: ERROR: endpoints do not straddle y=0\nDefault GSL error handler invoked.\n");
        abort();
    }
    return GSL_SUCCESS;
}

static int brent_iterate(void *vstate, gsl_function *f, double *root, double *x_lower, double *x_upper)
{
    brent_state_t *s = (brent_state_t *)vstate;
    double tol, m;
    int ac_equal = 0;
    double a = s->a, b = s->b, c = s->c, fa = s->fa, fb = s->fb, fc = s->fc, d = s->d, e = s->e;
    if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) { ac_equal = 1; c = a; fc = fa; d = b - a; e = b - a; }
    if (fabs(fc) < fabs(fb)) { ac_equal = 1; a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
    tol = 0.5 * GSL_DBL_EPSILON * fabs(b);
    m = 0.5 * (c - b);
    if (fb == 0) {
        *root = b; *x_lower = b; *x_upper = b;
        return GSL_SUCCESS;
    }
    if (fabs(m) <= tol) {
        *root = b;
        if (b < c) { *x_lower = b; *x_upper = c; } else { *x_lower = c; *x_upper = b; }
        return GSL_SUCCESS;
    }
    if (fabs(e) < tol || fabs(fa) <= fabs(fb)) {
        d = m; e = m;                                  /* bisection */
    } else {
        double p, q, r;
        double sv = fb / fa;
        if (ac_equal) { p = 2 * m * sv; q = 1 - sv; }  /* secant */
        else {                                         /* inverse quadratic */
            q = fa / fc; r = fb / fc;
            p = sv * (2 * m * q * (q - r) - (b - a) * (r - 1));
            q = (q - 1) * (r - 1) * (sv - 1);
        }
        if (p > 0) q = -q; else p = -p;
        if (2 * p < GSL_MIN(3 * m * q - fabs(tol * q), fabs(e * q))) { e = d; d = p / q; }
        else { d = m; e = m; }
    }
    a = b; fa = fb;
    if (fabs(d) > tol) b += d;
    else b += (m > 0 ? +tol : -tol);
    fb = GSL_FN_EVAL(f, b);
    s->a = a; s->b = b; s->c = c; s->d = d; s->e = e; s->fa = fa; s->fb = fb; s->fc = fc;
    *root = b;
    if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) c = a;
    if (b < c) { *x_lower = b; *x_upper = c; } else { *x_lower = c; *x_upper = b; }
    return GSL_SUCCESS;
}

static const gsl_root_fsolver_type brent_type = {"brent", sizeof(brent_state_t), &brent_init, &brent_iterate};
const gsl_root_fsolver_type *gsl_root_fsolver_brent = &brent_type;

gsl_root_fsolver *gsl_root_fsolver_alloc(const gsl_root_fsolver_type *T)
{
    gsl_root_fsolver *s = (gsl_root_fsolver *)malloc(sizeof(gsl_root_fsolver));
    s->state = calloc(1, T->size);
    s->type = T;
    s->function = NULL;
    return s;
}
int gsl_root_fsolver_set(gsl_root_fsolver *s, gsl_function *f, double x_lower, double x_upper)
{
    s->function = f;
    s->root = 0.5 * (x_lower + x_upper);
    s->x_lower = x_lower;
    s->x_upper = x_upper;
    return (s->type->set)(s->state, s->function, &(s->root), x_lower, x_upper);
}
int gsl_root_fsolver_iterate(gsl_root_fsolver *s)
{
    return (s->type->iterate)(s->state, s->function, &(s->root), &(s->x_lower), &(s->x_upper));
}
void gsl_root_fsolver_free(gsl_root_fsolver *s) { if (s) { free(s->state); free(s); } }
double gsl_root_fsolver_root(const gsl_root_fsolver *s) { return s->root; }
double gsl_root_fsolver_x_lower(const gsl_root_fsolver *s) { return s->x_lower; }
double gsl_root_fsolver_x_upper(const gsl_root_fsolver *s) { return s->x_upper; }

int gsl_root_test_interval(double x_lower, double x_upper, double epsabs, double epsrel)
{
    const double abs_lower = fabs(x_lower), abs_upper = fabs(x_upper);
    double min_abs, tolerance;
    if ((x_lower > 0.0 && x_upper > 0.0) || (x_lower < 0.0 && x_upper < 0.0)) min_abs = GSL_MIN_DBL(abs_lower, abs_upper);
    else min_abs = 0;
    tolerance = epsabs + epsrel * min_abs;
    if (fabs(x_upper - x_lower) < tolerance) return GSL_SUCCESS;
    return GSL_CONTINUE;
}
