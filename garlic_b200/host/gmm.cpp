// gmm.cpp — ROH size classes on the host (north-star: "the GMM length classification stays on the host"):
// 1-D Gaussian mixture by EM with the reference's deterministic initialisation, accumulation order and stopping
// rule (src/gmm.cpp:276-331,385-441; src/garlic-roh.cpp:935-1003), class boundaries by Brent's method on the
// difference of adjacent weighted Gaussians (src/BoundFinder.cpp; GSL brent semantics).
#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>

#include "garlic_host.h"

namespace gh {

namespace {

double normal_log(double x, double mean, double var)
{
    static const double C = (-0.5 * std::log(2 * M_PI));
    return C - (0.5 * std::log(var)) - ((x - mean) * (x - mean)) / (2.0 * var);
}

struct Pair { double mu1, var1, a1, mu2, var2, a2; };
double gauss_pdf(double x, double sigma)
{
    const double u = x / std::fabs(sigma);
    return (1 / (std::sqrt(2 * M_PI) * std::fabs(sigma))) * std::exp(-u * u / 2);
}
double diff(double x, const Pair& p)
{
    return p.a1 * gauss_pdf(x - p.mu1, std::sqrt(p.var1)) - p.a2 * gauss_pdf(x - p.mu2, std::sqrt(p.var2));
}

// Brent (1973) bracketing root finder stepped the way GSL's brent solver reports (root = current b, bracket =
// [min(b,c), max(b,c)]); stop when |hi - lo| < epsrel * min(|lo|, |hi|) (gsl_root_test_interval, epsabs = 0).
bool brent_boundary(const Pair& P, int max_iter, double epsrel, double& root)
{
    double a = std::min(P.mu1, P.mu2), b = std::max(P.mu1, P.mu2);
    double fa = diff(a, P), fb = diff(b, P);
    double c = b, fc = fb, d = b - a, e = b - a;
    if ((fa < 0.0 && fb < 0.0) || (fa > 0.0 && fb > 0.0)) return false;
    const double eps = std::numeric_limits<double>::epsilon();
    for (int it = 0; it < max_iter; ++it) {
        bool ac_equal = false;
        if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) { ac_equal = true; c = a; fc = fa; d = b - a; e = b - a; }
        if (std::fabs(fc) < std::fabs(fb)) { ac_equal = true; a = b; b = c; c = a; fa = fb; fb = fc; fc = fa; }
        const double tol = 0.5 * eps * std::fabs(b);
        const double m = 0.5 * (c - b);
        double lo, hi;
        if (fb == 0) { root = b; return true; }
        if (std::fabs(m) <= tol) {
            root = b;
            lo = std::min(b, c); hi = std::max(b, c);
        } else {
            if (std::fabs(e) < tol || std::fabs(fa) <= std::fabs(fb)) { d = m; e = m; }
            else {
                double p, q, r;
                const double s = fb / fa;
                if (ac_equal) { p = 2 * m * s; q = 1 - s; }
                else {
                    q = fa / fc; r = fb / fc;
                    p = s * (2 * m * q * (q - r) - (b - a) * (r - 1));
                    q = (q - 1) * (r - 1) * (s - 1);
                }
                if (p > 0) q = -q; else p = -p;
                if (2 * p < std::min(3 * m * q - std::fabs(tol * q), std::fabs(e * q))) { e = d; d = p / q; }
                else { d = m; e = m; }
            }
            a = b; fa = fb;
            if (std::fabs(d) > tol) b += d;
            else b += (m > 0 ? +tol : -tol);
            fb = diff(b, P);
            root = b;
            if ((fb < 0 && fc < 0) || (fb > 0 && fc > 0)) c = a;
            lo = std::min(b, c); hi = std::max(b, c);
        }
        const double al = std::fabs(lo), ah = std::fabs(hi);
        const double min_abs = ((lo > 0 && hi > 0) || (lo < 0 && hi < 0)) ? std::min(al, ah) : 0;
        if (std::fabs(hi - lo) < epsrel * min_abs) return true;
    }
    return false;
}

}  // namespace

bool size_classes(const std::vector<double>& x, int K, std::vector<double>& bounds)
{
    const size_t N = x.size();
    bounds.clear();
    if (N < 2) { LOG.error("ERROR: too few ROH to fit a Gaussian mixture; provide --size-bounds."); return false; }
    // gsl_stats_mean / gsl_stats_variance: running means in long double, variance scaled by N/(N-1)
    long double meanl = 0;
    for (size_t i = 0; i < N; ++i) meanl += (x[i] - meanl) / (i + 1);
    const double mu = (double)meanl;
    long double varl = 0;
    for (size_t i = 0; i < N; ++i) { const long double dd = x[i] - mu; varl += (dd * dd - varl) / (i + 1); }
    const double var = (double)varl * ((double)N / (double)(N - 1));
    std::vector<double> a(K), mean(K), v(K), resp(K), sw(K), swx(K), swx2(K);
    for (int n = 0; n < K; ++n) {
        a[n] = 1.0 / double(K);
        mean[n] = mu * double(n + 1) / double(K + 1);
        v[n] = var * (n + 1) / double(K);
    }
    double last = -std::numeric_limits<double>::max(), L = last;
    for (int it = 1; it <= 1000; ++it) {                           // GMM::update / estimate
        std::fill(sw.begin(), sw.end(), 0.0); std::fill(swx.begin(), swx.end(), 0.0); std::fill(swx2.begin(), swx2.end(), 0.0);
        L = 0;
        for (size_t j = 0; j < N; ++j) {
            double lmax = -std::numeric_limits<double>::max();
            for (int i = 0; i < K; ++i) {
                if (!(a[i] > 0) || !(v[i] > 0)) { LOG.error("gsl: log.c:116: ERROR: domain error (GMM component collapsed); provide --size-bounds."); return false; }
                resp[i] = std::log(a[i]) + normal_log(x[j], mean[i], v[i]);
                if (resp[i] > lmax) lmax = resp[i];
            }
            double sum = 0;
            for (int i = 0; i < K; ++i) sum += std::exp(resp[i] - lmax);
            const double tmp = lmax + std::log(sum);
            L += tmp;
            double den = 0;
            for (int i = 0; i < K; ++i) { resp[i] = std::exp(resp[i] - tmp); den += resp[i]; }
            for (int k = 0; k < K; ++k) {
                sw[k] += resp[k] / den;
                swx[k] += x[j] * resp[k] / den;
                swx2[k] += x[j] * x[j] * resp[k] / den;
            }
        }
        for (int k = 0; k < K; ++k) {
            a[k] = sw[k] / double(N);
            mean[k] = swx[k] / sw[k];
            v[k] = swx2[k] / sw[k] - mean[k] * mean[k];
        }
        if (std::fabs(L - last) <= 1e-5) break;
        last = L;
    }
    std::vector<size_t> order(K);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](size_t p, size_t q) { return mean[p] < mean[q]; });
    char cls = 'A';
    for (int i = 0; i < K; ++i, ++cls)
        LOG.line(std::string("Gaussian class ") + cls + " ( mixture, mean, std ) = ( " + fmt_g(a[order[i]]) + ", " +
                 fmt_g(mean[order[i]]) + ", " + fmt_g(v[order[i]]) + " )");
    for (int i = 1; i < K; ++i) {
        Pair P{mean[order[i - 1]], v[order[i - 1]], a[order[i - 1]], mean[order[i]], v[order[i]], a[order[i]]};
        double r;
        if (!brent_boundary(P, 1000, 1e-4, r)) { LOG.error("Root finder failed to converge after 1000 iterations."); return false; }
        bounds.push_back(r);
    }
    return true;
}

}  // namespace gh
