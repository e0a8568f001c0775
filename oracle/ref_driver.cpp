// ref_driver.cpp — harness that calls the REFERENCE's own hot-path functions on binary inputs.
//
// TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile (`make ref`) into oracle/_ref/ref_driver from
// the reference sources where they lie (/root/reference/src/*.cpp, never copied) plus gsl_shim.c.
// It feeds already-coded, already-filtered per-chromosome arrays to the reference's
//   calcLODWindows / calcwLODWindows (src/garlic-roh.cpp:279,311), calculateGenoFreq + calcHR2LD
//   (src/garlic-data.cpp:648,377) and assembleROHWindows (src/garlic-roh.cpp:409)
// and dumps their results as raw fp64, which is what pins the restatement (garlic_oracle.c) and the
// CUDA path to 1e-9 (the reference binary's --raw-lod prints 6 digits only).  It also times those
// calls: bench.py's `--impl reference` arm and cpu_baseline run this binary on the host cores.
//
// usage: ref_driver <in.bin> <out.bin>     (format: see oracle/refdrv.py)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "garlic-data.h"
#include "garlic-roh.h"

using namespace std;

static void rd(FILE *f, void *p, size_t n)
{
    if (n && fread(p, 1, n, f) != n) { fprintf(stderr, "ref_driver: short read\n"); exit(2); }
}
static double now()
{
    return chrono::duration<double>(chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: ref_driver in.bin out.bin\n"); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    char magic[4];
    rd(f, magic, 4);
    if (memcmp(magic, "GRLF", 4)) { fprintf(stderr, "bad magic\n"); return 2; }
    int32_t hdr[13];
    rd(f, hdr, sizeof(hdr));
    const int n_chr = hdr[0], N = hdr[1], W = hdr[2], max_gap = hdr[3], use_gl = hdr[4], weighted = hdr[5],
              cm = hdr[6], M = hdr[7], threads = hdr[8], n_ld = hdr[9], dump_win = hdr[10], dump_ld = hdr[11],
              do_roh = hdr[12];
    double dh[4];
    rd(f, dh, sizeof(dh));
    const double error = dh[0], cutoff = dh[1], overlap_frac = dh[2], mu = dh[3];

    LOG.init(string(argv[2]) + ".reflog");
    string cenfile = string(argv[2]) + ".cen";
    FILE *cf = fopen(cenfile.c_str(), "w");

    vector<HapData *> *hapByChr = new vector<HapData *>;
    vector<MapData *> *mapByChr = new vector<MapData *>;
    vector<FreqData *> *freqByChr = new vector<FreqData *>;
    vector<GenoLikeData *> *glByChr = use_gl ? new vector<GenoLikeData *> : NULL;
    for (int c = 0; c < n_chr; c++) {
        int32_t L, cen[2];
        char name[16];
        rd(f, &L, 4);
        rd(f, name, 16);
        rd(f, cen, 8);
        name[15] = 0;
        fprintf(cf, "%s %d %d\n", name, cen[0], cen[1]);
        MapData *m = new MapData;
        m->nloci = L;
        m->chr = name;
        m->physicalPos = new int[L];
        m->geneticPos = new double[L];
        m->locusName = new string[L];
        m->allele = new char[L];
        rd(f, m->physicalPos, sizeof(int) * L);
        rd(f, m->geneticPos, sizeof(double) * L);
        FreqData *fr = new FreqData;
        fr->nloci = L;
        fr->freq = new double[L];
        rd(f, fr->freq, sizeof(double) * L);
        HapData *h = new HapData;
        h->nind = N;
        h->nloci = L;
        h->firstCopy = NULL;
        h->data = new short *[L];
        vector<int8_t> row(N);
        for (int l = 0; l < L; l++) {
            h->data[l] = new short[N];
            rd(f, row.data(), N);
            for (int i = 0; i < N; i++) h->data[l][i] = (row[i] == 3) ? -9 : row[i];
        }
        if (use_gl) {
            GenoLikeData *g = new GenoLikeData;
            g->nind = N;
            g->nloci = L;
            g->data = new double *[L];
            for (int l = 0; l < L; l++) {
                g->data[l] = new double[N];
                rd(f, g->data[l], sizeof(double) * N);
            }
            glByChr->push_back(g);
        }
        hapByChr->push_back(h);
        mapByChr->push_back(m);
        freqByChr->push_back(fr);
    }
    fclose(cf);
    vector<int> ld_ind(n_ld > 0 ? n_ld : N);
    if (n_ld > 0) rd(f, ld_ind.data(), sizeof(int) * n_ld);
    else for (int i = 0; i < N; i++) ld_ind[i] = i;
    fclose(f);

    centromere *centro = new centromere("custom", cenfile, "none");
    IndData ind;
    ind.pop = "POP";
    ind.nind = N;
    ind.indID = new string[N];
    for (int i = 0; i < N; i++) ind.indID[i] = "i" + to_string(i);

    double t_ld = 0, t_win = 0, t_roh = 0;
    vector<LDData *> *ldByChr = NULL;
    vector<WinData *> *win = NULL;
    if (weighted) {
        double t0 = now();
        vector<GenoFreqData *> *gf = calculateGenoFreq(hapByChr);
        ldByChr = new vector<LDData *>;
        const int nsub = (int)ld_ind.size();
        for (int c = 0; c < n_chr; c++)   // what calcLDData does per chromosome, with our (not time-seeded) list
            ldByChr->push_back(calcHR2LD(hapByChr->at(c), gf->at(c), W, threads, ld_ind.data(), nsub));
        t_ld = now() - t0;
        t0 = now();
        win = calcwLODWindows(hapByChr, freqByChr, mapByChr, glByChr, ldByChr, centro, W, error, max_gap, use_gl != 0, M, mu, threads);
        t_win = now() - t0;
    } else {
        double t0 = now();
        win = calcLODWindows(hapByChr, freqByChr, mapByChr, glByChr, centro, W, error, max_gap, use_gl != 0);
        t_win = now() - t0;
    }
    vector<ROHData *> *roh = NULL;
    ROHLength *lens = NULL;
    if (do_roh) {
        double t0 = now();
        roh = assembleROHWindows(win, mapByChr, &ind, centro, cutoff, &lens, W, max_gap, overlap_frac, cm != 0);
        t_roh = now() - t0;
    }

    FILE *o = fopen(argv[2], "wb");
    if (!o) { perror(argv[2]); return 2; }
    double t[3] = {t_win, t_ld, t_roh};
    fwrite(t, sizeof(double), 3, o);
    int64_t n_roh = 0;
    if (roh) for (int i = 0; i < N; i++) n_roh += (int64_t)roh->at(i)->chr.size();
    fwrite(&n_roh, 8, 1, o);
    if (roh)
        for (int i = 0; i < N; i++) {
            ROHData *r = roh->at(i);
            for (size_t k = 0; k < r->chr.size(); k++) {
                int32_t ic[2] = {i, r->chr[k]};
                double v[3] = {r->start[k], r->stop[k], r->length[k]};
                fwrite(ic, 4, 2, o);
                fwrite(v, 8, 3, o);
            }
        }
    if (dump_win)
        for (int c = 0; c < n_chr; c++)
            for (int i = 0; i < N; i++) fwrite(win->at(c)->data[i], sizeof(double), win->at(c)->nloci, o);
    if (dump_ld && ldByChr)
        for (int c = 0; c < n_chr; c++)
            for (int l = 0; l < ldByChr->at(c)->nloci; l++) fwrite(ldByChr->at(c)->LD[l], sizeof(double), W, o);
    fclose(o);
    remove(cenfile.c_str());
    return 0;
}
