/*
 * garlic_oracle.c — CPU restatement of GARLIC v1.1.6a's LOD/wLOD → ROH hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (garlic_b200/, include/garlic_b200.h) never calls into this file.
 *
 * Every function restates one reference function and cites it (paths relative to
 * /root/reference).  The control flow follows the reference literally (skip-ahead of the
 * window loop, the MISSING sentinel test, the else-if ladder of the ROH state machine), on
 * purpose: the CUDA path uses closed forms (validity mask, segment table, bit-parallel
 * run-length) and the tests check that the two agree.
 *
 * Parity pinning: tests/test_oracle_vs_reference.py checks this file against outputs of the
 * reference binary (bin/linux/garlic) committed under tests/golden/ (.roh.bed exact, .freq,
 * --raw-lod to print precision) and, where /root/reference is present, against the reference
 * sources compiled into oracle/_ref (full-precision window dumps).
 *
 * Layouts (flat, row-major): geno is SNP-major int8 [L][N] with codes 0/1/2 and 3 = missing
 * (the reference's short -9); win is individual-major double [N][L] with MISSING = -9999.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MISSING (-9999.0)

/* ---- garlic-roh.cpp:11-16 inGap ------------------------------------------------------- */
int orc_in_gap(int qStart, int qEnd, int tStart, int tEnd)
{
    if (tStart <= qStart && tEnd >= qStart) return 1;
    if (tStart <= qEnd && tEnd >= qEnd) return 1;
    if (tStart >= qStart && tEnd <= qEnd) return 1;
    return 0;
}

/* ---- garlic-data.cpp:103-150 genotype coding + allele counting -------------------------
 * alleles: [L0][N][2] characters.  geno out: [L0][N].  one_allele: the "1" allele per SNP
 * (the missing char if every call is missing).  freq = nalleles/total or 0.               */
void orc_code_tped(const uint8_t *alleles, int L0, int N, int missing_char,
                   int8_t *geno, int32_t *nalleles, int32_t *total, uint8_t *one_allele,
                   double *freq)
{
    for (int l = 0; l < L0; l++) {
        int na = 0, tot = 0;
        uint8_t one = (uint8_t)missing_char;
        for (int i = 0; i < N; i++) {
            uint8_t a1 = alleles[((size_t)l * N + i) * 2];
            uint8_t a2 = alleles[((size_t)l * N + i) * 2 + 1];
            int d = 0;
            if (one == missing_char && a1 != missing_char) one = a1;
            if (one == missing_char && a2 != missing_char) one = a2;
            if (a1 == missing_char) d += -9;
            else if (a1 == one) { d += 1; na++; tot++; }
            else tot++;
            if (a2 == missing_char) d += -9;
            else if (a2 == one) { d += 1; na++; tot++; }
            else tot++;
            if (d < 0) d = 3; /* the reference stores -9 */
            geno[(size_t)l * N + i] = (int8_t)d;
        }
        nalleles[l] = na;
        total[l] = tot;
        one_allele[l] = one;
        freq[l] = (tot == 0) ? 0 : ((double)na / (double)tot);
    }
}

/* ---- garlic-data.cpp:656-676 calculateGenoFreq ----------------------------------------- */
void orc_hom_freq(const int8_t *geno, int L, int N, double *hom_freq)
{
    for (int l = 0; l < L; l++) {
        double total = 0, hom = 0;
        for (int i = 0; i < N; i++) {
            int g = geno[(size_t)l * N + i];
            if (g != 3) {
                if (g == 2 || g == 0) hom++;
                total++;
            }
        }
        hom /= total;
        hom_freq[l] = hom;
    }
}

/* ---- garlic-data.cpp:963-988 / 1066-1098 site filter predicate -------------------------
 * oob != 0 adds the map-scaffold / centromere conditions of filterMonomorphicAndOOBSites.  */
void orc_keep_mask(const double *freq, const int32_t *pos, int L, int oob,
                   int scaf_first, int scaf_last, int cen_start, int cen_end, uint8_t *keep)
{
    for (int i = 0; i < L; i++) {
        int k = (freq[i] > 0 && freq[i] < 1);
        if (oob) {
            k = k && !(pos[i] < scaf_first) && !(pos[i] > scaf_last) &&
                !(pos[i] > cen_start && pos[i] < cen_end);
        }
        keep[i] = (uint8_t)k;
    }
}

/* ---- garlic-data.cpp:1555-1577 per-genotype error from GQ(0)/GL(1)/PL(2) ---------------- */
double orc_gl_error(double gl, int type)
{
    if (type == 0) {
        gl /= (-10.0);
        gl = (gl > -10) ? gl : -10;
        gl = pow(10, gl);
    } else if (type == 1) {
        gl = (gl > -10) ? gl : -10;
        gl = 1 - pow(10, gl);
    } else {
        gl /= (-10.0);
        gl = (gl > -10) ? gl : -10;
        gl = 1 - pow(10, gl);
    }
    if (gl <= 0) gl = 0.0000000000000001;
    if (gl > 1) gl = 1;
    return gl;
}

void orc_gl_error_array(const double *in, size_t n, int type, double *out)
{
    for (size_t i = 0; i < n; i++) out[i] = orc_gl_error(in[i], type);
}

/* ---- garlic-roh.cpp:355-386 lod() ------------------------------------------------------ */
double orc_lod(int g, double freq, double error)
{
    double a, na;
    if (freq == 0 || freq == 1) { a = 1; na = 1; }
    else if (g == 0) {
        na = (1 - freq) * (1 - freq);
        a = (1 - error) * (1 - freq) + error * na;
    } else if (g == 1) {
        na = 2 * (freq) * (1 - freq);
        a = error * na;
    } else if (g == 2) {
        na = (freq) * (freq);
        a = (1 - error) * (freq) + error * na;
    } else { a = 1; na = 1; }
    return log10(a / na);
}

/* per-genotype lod() for per-genotype error rates: out[l][i] = lod(geno[l][i], freq[l], err[l][i]) */
void orc_lod_matrix(const int8_t *geno, const double *freq, const double *err, int L, int N, double *out)
{
    for (int l = 0; l < L; l++)
        for (int i = 0; i < N; i++) {
            const size_t k = (size_t)l * N + i;
            out[k] = orc_lod(geno[k], freq[l], err[k]);
        }
}

/* per-SNP table lut[L][4] for a global error rate (what K4 builds on the device) */
void orc_lod_lut(const double *freq, int L, double error, double *lut)
{
    for (int l = 0; l < L; l++)
        for (int g = 0; g < 4; g++) lut[(size_t)l * 4 + g] = orc_lod(g, freq[l], error);
}

/* ---- garlic-roh.cpp:18-132 calcLOD (one chromosome) -------------------------------------
 * gl: per-genotype error [L][N] or NULL.  win: [N][L], fully initialised here to MISSING
 * (initWinData, garlic-data.cpp:1608-1638).                                                */
void orc_calc_lod(const int8_t *geno, const double *freq, const int32_t *pos, int L, int N,
                  int W, double error, int max_gap, int cStart, int cEnd, const double *gl,
                  double *win)
{
    int start = 0, stop = L;
    for (size_t k = 0; k < (size_t)N * L; k++) win[k] = ORC_MISSING;
    if (L - stop < W) stop = L - W + 1;
#define G(l, i) geno[(size_t)(l) * N + (i)]
#define E(l, i) (gl ? gl[(size_t)(l) * N + (i)] : error)
    for (int ind = 0; ind < N; ind++) {
        double *w = win + (size_t)ind * L;
        for (int locus = start; locus < stop; locus++) {
            w[locus] = 0;
            int fresh = (locus == start) || (w[locus - 1] == ORC_MISSING);
            if (fresh) {
                int prevI = locus;
                for (int i = locus; i < locus + W; i++) {
                    if (pos[i] - pos[prevI] > max_gap || orc_in_gap(pos[prevI], pos[i], cStart, cEnd)) {
                        w[locus] = ORC_MISSING;
                        locus = prevI;
                        break;
                    }
                    w[locus] += orc_lod(G(i, ind), freq[i], E(i, ind));
                    prevI = i;
                }
            } else {
                int a = locus + W - 2, b = locus + W - 1;
                if (pos[b] - pos[a] > max_gap || orc_in_gap(pos[a], pos[b], cStart, cEnd)) {
                    w[locus] = ORC_MISSING;
                    locus = locus + W - 2;
                } else {
                    w[locus] = w[locus - 1] - orc_lod(G(locus - 1, ind), freq[locus - 1], E(locus - 1, ind))
                               + orc_lod(G(b, ind), freq[b], E(b, ind));
                }
            }
        }
    }
}

/* ---- garlic-data.cpp:558-583 hr2 ------------------------------------------------------- */
static double orc_hr2(const int8_t *geno, const double *hom_freq, int N, int i, int j,
                      const int32_t *ind_index, int nsub)
{
    double HA = hom_freq[i], HB = hom_freq[j];
    if (HA > 0 && HA < 1 && HB > 0 && HB < 1) {
        double HAB = 0, total = 0;
        for (int k = 0; k < nsub; k++) {
            int ind = ind_index[k];
            int gi = G(i, ind), gj = G(j, ind);
            if (gi != 3 && gj != 3) {
                total++;
                if (gi != 1 && gj != 1) HAB++;
            }
        }
        HAB /= total;
        double H = HAB - HA * HB;
        double HR2 = H * H / (HA * (1 - HA) * HB * (1 - HB));
        if (HR2 > 1) return 1;
        return HR2;
    }
    return 0;
}

/* ---- garlic-data.cpp:377-424,474-495,521-527,619-631 calcHR2LD (one chromosome) ---------
 * LD: [L][W], zero-initialised here; rows >= L-W+1 stay zero.                             */
void orc_calc_hr2_ld(const int8_t *geno, const double *hom_freq, int L, int N, int W,
                     const int32_t *ind_index, int nsub, double *LD)
{
    memset(LD, 0, sizeof(double) * (size_t)L * W);
    int stop = L;
    if (L - stop < W) stop = L - W + 1;
    for (int locus = 0; locus < stop; locus++) {
        for (int site = locus; site < locus + W; site++) {
            double *cell = &LD[(size_t)locus * W + (site - locus)];
            for (int i = locus; i <= locus + W - 1; i++) {
                if (i != site) *cell += orc_hr2(geno, hom_freq, N, i, site, ind_index, nsub);
                else *cell += 1;
            }
        }
    }
}

/* ---- garlic-data.cpp:585-617 r2 (--phased): first_copy[l*N+ind] = (first allele character == the "1" allele) ---- */
static double orc_r2(const int8_t *geno, const uint8_t *first_copy, const double *freq, int N, int i, int j,
                     const int32_t *ind_index, int nsub)
{
    double pi = freq[i], pj = freq[j];
    if (pi > 0 && pi < 1 && pj > 0 && pj < 1) {
        double x11 = 0, total = 0;
        for (int k = 0; k < nsub; k++) {
            int ind = ind_index[k];
            int gi = G(i, ind), gj = G(j, ind);
            if (gi != 3 && gj != 3) {
                total += 2;
                if (gi == 2 && gj == 2) x11 += 2;
                else if (gi == 1 && gj == 2) x11++;
                else if (gi == 2 && gj == 1) x11++;
                else if (gi == 1 && gj == 1 &&
                         first_copy[(size_t)j * N + ind] == first_copy[(size_t)i * N + ind]) x11++;
            }
        }
        x11 /= total;
        double D = x11 - pi * pj;
        double R2 = D * D / (pi * (1 - pi) * pj * (1 - pj));
        if (R2 > 1) return 1;
        return R2;
    }
    return 0;
}

/* ---- garlic-data.cpp:426-471,497-519,529-535 calcR2LD (one chromosome), layout as orc_calc_hr2_ld ---- */
void orc_calc_r2_ld(const int8_t *geno, const uint8_t *first_copy, const double *freq, int L, int N, int W,
                    const int32_t *ind_index, int nsub, double *LD)
{
    memset(LD, 0, sizeof(double) * (size_t)L * W);
    int stop = L;
    if (L - stop < W) stop = L - W + 1;
    for (int locus = 0; locus < stop; locus++) {
        for (int site = locus; site < locus + W; site++) {
            double *cell = &LD[(size_t)locus * W + (site - locus)];
            for (int i = locus; i <= locus + W - 1; i++) {
                if (i != site) *cell += orc_r2(geno, first_copy, freq, N, i, site, ind_index, nsub);
                else *cell += 1;
            }
        }
    }
}

/* ---- garlic-roh.cpp:134-140 nomut / norec ---------------------------------------------- */
static double orc_nomut(double M, double mu, double interval) { return exp(-2.0 * M * mu * interval); }
static double orc_norec(double M, double interval) { return orc_nomut(M, 1, interval); }

/* per-SNP weight nomut*norec is not a single product in the reference: score = lod*nomut*norec
 * evaluated left to right (garlic-roh.cpp:249); expose the two factors separately.          */
void orc_wlod_weights(const int32_t *pos, const double *gpos, int L, double mu, int M,
                      double *nomut, double *norec)
{
    for (int l = 0; l < L; l++) {
        double pi = (l > 0) ? (pos[l] - pos[l - 1]) : pos[l];
        double gi = (l > 0) ? (gpos[l] - gpos[l - 1]) : gpos[l];
        nomut[l] = orc_nomut(M, mu, pi);
        norec[l] = orc_norec(M, gi);
    }
}

/* ---- garlic-roh.cpp:144-277 calcwLOD / parallelwLOD (one chromosome; thread partition is
 * result-neutral) ---------------------------------------------------------- */
void orc_calc_wlod(const int8_t *geno, const double *freq, const int32_t *pos, const double *gpos,
                   int L, int N, int W, double error, int max_gap, int cStart, int cEnd,
                   const double *gl, const double *LD, double mu, int M, double *win)
{
    int start = 0, stop = L;
    for (size_t k = 0; k < (size_t)N * L; k++) win[k] = ORC_MISSING;
    if (L - stop < W) stop = L - W + 1;
    double *score = (double *)malloc(sizeof(double) * (size_t)(L > 0 ? L : 1));
    for (int ind = 0; ind < N; ind++) {
        double *w = win + (size_t)ind * L;
        int lim = (stop + W + 1 > L) ? L : stop + W + 1;
        for (int locus = start; locus < lim; locus++) {
            double e = E(locus, ind);
            double pi = (locus > 0) ? (pos[locus] - pos[locus - 1]) : pos[locus];
            double gi = (locus > 0) ? (gpos[locus] - gpos[locus - 1]) : gpos[locus];
            score[locus - start] = orc_lod(G(locus, ind), freq[locus], e) * orc_nomut(M, mu, pi) * orc_norec(M, gi);
        }
        for (int locus = start; locus < stop; locus++) {
            w[locus] = 0;
            int prevI = locus;
            for (int i = locus; i < locus + W; i++) {
                if (pos[i] - pos[prevI] > max_gap || orc_in_gap(pos[prevI], pos[i], cStart, cEnd)) {
                    w[locus] = ORC_MISSING;
                    locus = prevI;
                    break;
                }
                w[locus] += score[i - start] * (1.0 / LD[(size_t)locus * W + (i - locus)]);
                prevI = i;
            }
        }
    }
    free(score);
}
#undef G
#undef E

/* ---- garlic-data.cpp:2026-2069 / 2071-2150 thinning for the KDE -------------------------
 * One chromosome: appends to out[] the windows of the listed individuals at locus 0,step,…
 * that are neither MISSING nor NaN.  Returns the number appended.                          */
size_t orc_thin(const double *win, int L, const int32_t *ind_list, int nsub, int step, double *out)
{
    size_t n = 0;
    for (int k = 0; k < nsub; k++) {
        const double *w = win + (size_t)ind_list[k] * L;
        for (int locus = 0; locus < L; locus += step) {
            double x = w[locus];
            if (x != ORC_MISSING && !(isnan(x))) out[n++] = x;
        }
    }
    return n;
}

/* ---- garlic-roh.cpp:409-546 assembleROHWindows (one individual, one chromosome) ---------
 * Emits (start index, stop index, length) triples; returns the count.  cap = capacity.     */
int orc_assemble(const double *w /*[L] windows of this individual*/, const int32_t *pos,
                 const double *gpos, int L, double cutoff, int W, int max_gap,
                 double overlap_frac, int cm, int cStart, int cEnd,
                 int32_t *out_start_idx, int32_t *out_stop_idx, double *out_len, int cap)
{
    double thr = overlap_frac * W;
    thr = (thr >= 1) ? thr : 1;
    thr = (thr <= W) ? thr : W;
    /* the reference allocates L entries (:437) and writes inWin[k+i] only for non-MISSING windows,
     * which end at L-1; a cutoff <= MISSING would make MISSING windows pass and overrun the array
     * there (undefined behaviour) — the product rejects such cutoffs, the oracle pads. */
    short *inWin = (short *)calloc((size_t)L + (size_t)W, sizeof(short));
    int n = 0;
    for (int k = 0; k < L; k++)
        if (w[k] >= cutoff)
            for (int i = 0; i < W; i++) inWin[k + i]++;
    double gStart = -1, gStop = -1;
    int winStart = -1, startIdx = -1, winStop = -1, stopIdx = -1;
#define EMIT()                                                                   \
    do {                                                                         \
        if (stopIdx - startIdx + 1 >= thr) {                                     \
            double size = cm ? gStop - gStart : winStop - winStart + 1;          \
            if (n < cap) { out_start_idx[n] = startIdx; out_stop_idx[n] = stopIdx; out_len[n] = size; } \
            n++;                                                                 \
        }                                                                        \
    } while (0)
    for (int k = 0; k < L; k++) {
        if (winStart < 0 && inWin[k] >= thr) {
            gStart = gpos ? gpos[k] : 0; winStart = pos[k]; startIdx = k;
        } else if (inWin[k] >= thr && (pos[k] - pos[k - 1] > max_gap ||
                                       orc_in_gap(pos[k - 1], pos[k], cStart, cEnd))) {
            gStop = gpos ? gpos[k - 1] : 0; winStop = pos[k - 1]; stopIdx = k - 1;
            EMIT();
            gStart = gpos ? gpos[k] : 0; winStart = pos[k]; startIdx = k;
        } else if (winStart > 0 && !(inWin[k] >= thr)) {
            gStop = gpos ? gpos[k - 1] : 0; winStop = pos[k - 1]; stopIdx = k - 1;
            EMIT();
            gStart = -1; winStart = -1; startIdx = -1;
        } else if (winStart > 0 && k + 1 >= L) {
            gStop = gpos ? gpos[k] : 0; winStop = pos[k]; stopIdx = k;
            EMIT();
            gStart = -1; winStart = -1; startIdx = -1;
        }
    }
#undef EMIT
    free(inWin);
    return n;
}

/* ---- garlic-data.cpp:754-757 interpolate, :718-752 getMapInfo (scaffold cursor) ---------- */
double orc_interpolate(double x0, double y0, double x1, double y1, double q)
{
    return (((y1 - y0) / (x1 - x0)) * q + (y0 - ((y1 - y0) / (x1 - x0)) * x0));
}

/* returns number interpolated, or -1 if a query is outside the scaffold */
int orc_interpolate_map(const int32_t *pos, int L, const int32_t *spos, const double *sgen, int S,
                        double *gpos)
{
    int cur = 0, count = 0;
    for (int i = 0; i < L; i++) {
        int q = pos[i];
        if (q < spos[0] || q > spos[S - 1]) return -1;
        /* exact hit: the reference looks the position up in a map<int,int> (last one wins) */
        int hit = -1;
        {
            int lo = 0, hi = S - 1;
            while (lo <= hi) { int mid = (lo + hi) / 2; if (spos[mid] < q) lo = mid + 1; else if (spos[mid] > q) hi = mid - 1; else { hit = mid; break; } }
            while (hit >= 0 && hit + 1 < S && spos[hit + 1] == q) hit++;
        }
        if (hit >= 0) { gpos[i] = sgen[hit]; continue; }
        int s = -1, e = -1;
        for (; cur < S - 1; cur++) {
            if (q > spos[cur] && q < spos[cur + 1]) { s = cur; e = cur + 1; break; }
        }
        if (s < 0) return -1;
        count++;
        gpos[i] = orc_interpolate(spos[s], sgen[s], spos[e], sgen[e], q);
    }
    return count;
}

/* ---- garlic-data.cpp:3-8, garlic-roh.cpp:3-9 heuristics -------------------------------- */
double orc_select_overlap_frac(double density, int winsize)
{
    double frac = (6.375 * log(density) + 63.888) / 100.0;
    if (frac > 1) frac = 1.0;
    if (frac <= 0) frac = 1.0 / (double)winsize;
    return frac;
}

int orc_select_winsize_weighted(double density)
{
    int size = (int)(8.3235 * log(density) + 138.0521 + 0.5);
    return (size >= 10 ? size : 10);
}
