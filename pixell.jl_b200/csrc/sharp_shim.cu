// sharp_shim.cu -- libsharp2-symbol-compatible front end over the pixsht C ABI (see include/pixsht_sharp_shim.h).
// Plain host code: it only calls the public functions of include/pixsht.h.
#include "../../include/pixsht.h"
#include "../../include/pixsht_sharp_shim.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

struct sharp_geom_info {
    int nrings = 0, nphi = 0;
    double phi0 = 0;
    bool supported = true;
    std::string why;
    std::vector<double> theta, wgt;
    struct Cached { int lmax, mmax, dtype; pixsht_plan* plan; };
    std::vector<Cached> plans;
    std::mutex mu;
};
struct sharp_alm_info { int lmax = 0, mmax = 0, stride = 1; };

static thread_local int g_shim_status = 0;
static void shim_fail(int code, const char* what)
{
    g_shim_status = code;
    fprintf(stderr, "pixsht sharp shim: %s (%s)\n", what, pixsht_last_error());
}

extern "C" int pixsht_shim_status(void) { return g_shim_status; }

extern "C" void sharp_make_geom_info(int nrings, const int* nph, const ptrdiff_t* ofs, const int* stride, const double* phi0,
                                     const double* theta, const double* wgt, sharp_geom_info** geom_info)
{
    g_shim_status = 0;
    if (!geom_info) return;
    sharp_geom_info* G = new sharp_geom_info();
    G->nrings = nrings;
    if (nrings < 1 || !nph || !ofs || !stride || !phi0 || !theta) { G->supported = false; G->why = "empty or null geometry"; *geom_info = G; return; }
    G->nphi = nph[0]; G->phi0 = phi0[0];
    G->theta.assign(theta, theta + nrings);
    if (wgt) G->wgt.assign(wgt, wgt + nrings); else G->wgt.assign(nrings, 0.0);
    for (int i = 0; i < nrings; ++i) {
        if (nph[i] != G->nphi || stride[i] != 1 || ofs[i] != (ptrdiff_t)i * G->nphi || phi0[i] != G->phi0) {
            G->supported = false; G->why = "only equal-length contiguous rings with a common phi0 are supported"; break;
        }
        if (i > 0 && !(theta[i] > theta[i - 1])) { G->supported = false; G->why = "rings must ascend in colatitude"; break; }
    }
    *geom_info = G;
}

extern "C" void sharp_destroy_geom_info(sharp_geom_info* G)
{
    if (!G) return;
    for (auto& c : G->plans) pixsht_plan_destroy(c.plan);
    delete G;
}

extern "C" ptrdiff_t sharp_map_size(const sharp_geom_info* G) { return G ? (ptrdiff_t)G->nrings * G->nphi : 0; }

extern "C" void sharp_make_triangular_alm_info(int lmax, int mmax, int stride, sharp_alm_info** alm_info)
{
    g_shim_status = 0;
    if (!alm_info) return;
    sharp_alm_info* A = new sharp_alm_info();
    A->lmax = lmax; A->mmax = mmax; A->stride = stride;
    *alm_info = A;
}
extern "C" void sharp_destroy_alm_info(sharp_alm_info* A) { delete A; }
extern "C" ptrdiff_t sharp_alm_count(const sharp_alm_info* A) { return A ? (ptrdiff_t)pixsht_nalm(A->lmax, A->mmax) : 0; }

extern "C" void sharp_execute(sharp_jobtype type, int spin, void* alm, void* map, const sharp_geom_info* geom_info,
                              const sharp_alm_info* A, int flags, double* time, unsigned long long* opcnt)
{
    g_shim_status = 0;
    const auto t0 = std::chrono::steady_clock::now();
    if (time) *time = 0;
    if (opcnt) *opcnt = 0;
    sharp_geom_info* G = const_cast<sharp_geom_info*>(geom_info);
    if (!G || !A || !alm || !map) { shim_fail(PIXSHT_ERR_ARG, "sharp_execute: null argument"); return; }
    if (!G->supported) { shim_fail(PIXSHT_ERR_UNSUPPORTED, ("sharp_execute: " + G->why).c_str()); return; }
    if (A->stride != 1 || A->mmax > A->lmax) { shim_fail(PIXSHT_ERR_UNSUPPORTED, "sharp_execute: alm stride must be 1 and mmax <= lmax"); return; }
    if (type != SHARP_MAP2ALM && type != SHARP_ALM2MAP) { shim_fail(PIXSHT_ERR_UNSUPPORTED, "sharp_execute: only SHARP_MAP2ALM / SHARP_ALM2MAP"); return; }
    if (spin != 0 && spin != 2) { shim_fail(PIXSHT_ERR_UNSUPPORTED, "sharp_execute: only spin 0 and spin 2"); return; }
    if (flags & (SHARP_ADD | SHARP_NO_FFT)) { shim_fail(PIXSHT_ERR_UNSUPPORTED, "sharp_execute: SHARP_ADD / SHARP_NO_FFT are not supported"); return; }
    const int dtype = (flags & SHARP_DP) ? PIXSHT_F64 : PIXSHT_F32;
    const int ncomp = spin == 0 ? 1 : 2;

    pixsht_plan* plan = nullptr;
    {
        std::lock_guard<std::mutex> lock(G->mu);
        for (auto& c : G->plans) if (c.lmax == A->lmax && c.mmax == A->mmax && c.dtype == dtype) plan = c.plan;
        if (!plan) {
            int rc = pixsht_plan_create_rings(&plan, G->nrings, G->theta.data(), G->wgt.data(), G->nphi, G->phi0, A->lmax, A->mmax, dtype, 0);
            if (rc != PIXSHT_OK) { shim_fail(rc, "sharp_execute: plan creation failed"); return; }
            G->plans.push_back({A->lmax, A->mmax, dtype, plan});
        }
    }
    int rc = pixsht_execute(plan, type == SHARP_MAP2ALM ? PIXSHT_MAP2ALM : PIXSHT_ALM2MAP, ncomp, (void* const*)alm, (void* const*)map, PIXSHT_HOST);
    if (rc != PIXSHT_OK) { shim_fail(rc, "sharp_execute: transform failed"); return; }
    if (time) *time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (opcnt) {
        // nominal operation count of SURVEY.md 8(d): (4 | 12) FMA = (8 | 24) flop per (l, m, ring pair)
        const double nlm = (double)pixsht_nalm(A->lmax, A->mmax), nrp = std::ceil(G->nrings / 2.0);
        *opcnt = (unsigned long long)(nlm * nrp * (spin == 0 ? 8.0 : 24.0));
    }
}
