// kde.cpp — cutoff selection on the host, as the north-star keeps it (src/garlic-kde.cpp): nrd0 bandwidth,
// 512-point Gauss transform through FIGTree (the reference's vendored library, linked — not re-implemented —
// when the build finds it), "minimum between modes" heuristic, wiggle statistic for window-size selection.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <limits>

#include "garlic_host.h"

#ifdef GARLIC_HAVE_FIGTREE
#include <figtree.h>   // from the reference's include/ (build-time -I; see host/Makefile)
#endif

namespace gh {

namespace {

// gsl_stats_sd / gsl_stats_quantile_from_sorted_data conventions: running means in
// long double, N-1 denominator; quantile at f·(N-1) with linear interpolation.
double stats_sd(const double* x, size_t n)
{
    long double mean = 0;
    for (size_t i = 0; i < n; ++i) mean += (x[i] - mean) / (i + 1);
    const double m = (double)mean;
    long double var = 0;
    for (size_t i = 0; i < n; ++i) {
        const long double d = x[i] - m;
        var += (d * d - var) / (i + 1);
    }
    return std::sqrt((double)var * ((double)n / (double)(n - 1)));
}
double quantile_sorted(const double* x, size_t n, double f)
{
    const double index = f * (n - 1);
    const size_t lhs = (size_t)index;
    const double delta = index - lhs;
    if (n == 0) return 0.0;
    if (lhs == n - 1) return x[lhs];
    return (1 - delta) * x[lhs] + delta * x[lhs + 1];
}
double nrd0(std::vector<double>& x)   // garlic-kde.cpp:130-140
{
    std::sort(x.begin(), x.end());
    const size_t n = x.size();
    const double hi = stats_sd(x.data(), n);
    const double iqr = quantile_sorted(x.data(), n, 0.75) - quantile_sorted(x.data(), n, 0.25);
    const double lo = std::min(hi, iqr / 1.34);
    return 0.9 * lo * std::pow((double)n, -0.2);
}
int arg_max(const double* v, int n)   // seed numeric_limits<double>::min() as the reference (garlic-kde.cpp:241)
{
    double mx = std::numeric_limits<double>::min();
    int a = -1;
    for (int i = 0; i < n; ++i) if (mx < v[i]) { mx = v[i]; a = i; }
    return a;
}
int arg_min(const double* v, int n)
{
    double mn = std::numeric_limits<double>::max();
    int a = -1;
    for (int i = 0; i < n; ++i) if (mn > v[i]) { mn = v[i]; a = i; }
    return a;
}

}  // namespace

void compute_kde(std::vector<double>& data, Kde& k, bool direct)   // computeKDE, garlic-kde.cpp:14-101
{
    const int n = (int)data.size();
    LOG.line("KDE with " + std::to_string(n) + " points.");
    const int M = 512;
    const double h = nrd0(data);
    double mn = data[0], mx = data[0];
    for (double v : data) { mn = std::min(mn, v); mx = std::max(mx, v); }
    mx += 3 * h;
    mn -= 3 * h;
    k.x.assign(M, 0.0);
    k.y.assign(M, 0.0);
    for (int i = 0; i < M; ++i) k.x[i] = (double(i + 1) / double(M)) * (mx - mn) + mn;
    const double spacing = k.x[1] - k.x[0];
    std::vector<double> q(n, 1.0 / double(n));
#ifdef GARLIC_HAVE_FIGTREE
    // NB: FIGTree's IFGT clustering is seeded from the clock inside the library, so this transform — in the
    // reference binary just as here — differs from run to run within its ε = 1e-2 (measured: ±0.3 % of the peak,
    // enough to move the selected cutoff by one grid step on tests/golden/auto_cutoff).
    // --kde-direct asks FIGTree for its exact evaluation instead (reproducible).
    figtree(1, n, M, 1, data.data(), h, q.data(), k.x.data(), 1e-2, k.y.data(), direct ? FIGTREE_EVAL_DIRECT : FIGTREE_EVAL_AUTO);
#else
    (void)direct;
    LOG.line("KDE: built without FIGTree (reference tree not present at build time): exact Gauss transform.");
    // FIGTree not available at build time: exact Gauss transform Σ q_j exp(-(t-x_j)²/h²) (FIGTree's kernel
    // convention); differs from the reference's ε = 1e-2 approximation in the last digits of the .kde
    for (int i = 0; i < M; ++i) {
        double s = 0;
        for (int j = 0; j < n; ++j) { const double d = (k.x[i] - data[j]) / h; s += q[j] * std::exp(-d * d); }
        k.y[i] = s;
    }
#endif
    double sum = 0;
    for (int i = 0; i < M; ++i) sum += k.y[i];
    for (int i = 0; i < M; ++i) k.y[i] /= (sum * spacing);
}

// get_min_btw_modes (garlic-kde.cpp:142-234): the two most persistent maxima of a 20-point sliding arg-max,
// then the arg-min between them; 0 if |x/W| >= 1.  Behaviour-for-behaviour, including the i == 1 case.
double min_between_modes(const Kde& k, int wsize)
{
    const int size = (int)k.x.size(), win = 20, m = size - win;
    const double* y = k.y.data();
    std::vector<double> umax(m, 0.0), ucnt(m, 0.0);
    int index = 0;
    for (int i = 0; i < m; ++i) {
        const double mx = y[arg_max(y + i, win) + i];
        if (i == 1) { umax[i] = mx; ucnt[i]++; }
        else if (umax[index] == mx) ucnt[index]++;
        else { index++; umax[index] = mx; ucnt[index]++; }
    }
    int c1 = (int)ucnt[0], c2 = 0;
    for (int i = 1; i < m; ++i) {
        if (c1 <= ucnt[i]) { c2 = c1; c1 = (int)ucnt[i]; }
        else if (c2 <= ucnt[i]) c2 = (int)ucnt[i];
    }
    std::vector<double> values;
    for (int i = 0; i < m; ++i)
        if (c1 == ucnt[i] || c2 == ucnt[i]) values.push_back(umax[i]);
    double first = -1, second = -1;
    for (double v : values) {
        if (first <= v) { second = first; first = v; }
        else if (second <= v) second = v;
    }
    int li = -1, ri = -1;
    for (int i = 0; i < size; ++i) {
        if (y[i] == first) li = i;
        if (y[i] == second) ri = i;
    }
    if (ri < li) std::swap(li, ri);
    if (li < 0) return 0;   // the reference would read out of bounds here; no mode pair → no cutoff
    const int mi = arg_min(y + li, ri - li + 1) + li;
    if (std::fabs(k.x[mi] / wsize) < 1) return k.x[mi];
    return 0;
}

// calculateWiggle (garlic-kde.cpp:3-12): y is scaled by 100 IN PLACE (the selected KDE is written scaled);
// sum over i of the residual sum of squares of a 20-point least-squares line, divided by 20.
double wiggle(Kde& k, int fit)
{
    const int size = (int)k.x.size();
    for (int i = 0; i < size; ++i) k.y[i] = k.y[i] * 100;
    double tot = 0;
    for (int i = 0; i < size - fit; ++i) {
        const double* x = &k.x[i];
        const double* y = &k.y[i];
        double mx = 0, my = 0, dx2 = 0, dxdy = 0;     // gsl_fit_linear's running-mean accumulation
        for (int j = 0; j < fit; ++j) { mx += (x[j] - mx) / (j + 1.0); my += (y[j] - my) / (j + 1.0); }
        for (int j = 0; j < fit; ++j) {
            const double dx = x[j] - mx, dy = y[j] - my;
            dx2 += (dx * dx - dx2) / (j + 1.0);
            dxdy += (dx * dy - dxdy) / (j + 1.0);
        }
        const double b = dxdy / dx2;
        double d2 = 0;
        for (int j = 0; j < fit; ++j) {
            const double dx = x[j] - mx, dy = y[j] - my, d = dy - b * dx;
            d2 += d * d;
        }
        tot += d2 / double(fit);
    }
    return tot;
}

}  // namespace gh
