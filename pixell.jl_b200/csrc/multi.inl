// multi.inl -- one process, several GPUs behind the same blocking call (included at the end of pixsht.cu).
//
// The reference's transform is ONE blocking call on caller-owned host arrays (src/transforms.jl:88-108, 185-194: sharp_execute!).
// A multi-GPU plan keeps that shape: pixsht_plan_create_multi(..., ndev, devices) builds one ordinary plan per GPU (tables
// replicated, SURVEY.md 8e "nothing else is partitioned") and pixsht_execute on it drives all of them from the calling thread:
//
//   alm2map: every GPU copies ITS alm columns straight from the caller's arrays (m-major alm: a run of consecutive m is one
//            contiguous range), runs the Legendre stage on them for all rings into its own phase buffer, then -- after an
//            event barrier across the GPUs -- the ring FFTs of ITS slab of rings, whose row loads fetch every m from the
//            phase buffer of the GPU that owns it (peer memory over NVLink: the phase transpose is fused into the FFT
//            kernels' row I/O, fft.cuh phase_elem), and copies its rows straight into the caller's map.
//   map2alm: the same backwards (row stores put every m into its owner's buffer).
//
// There is no torch, no NCCL and no second process here: streams and events order the stages, cudaDeviceEnablePeerAccess
// makes the phase buffers mutually visible.  IQU is pipelined per spin family so that copies hide under the Legendre work
// of the other family (alm2map: polarisation first, the T rows are the only output left after the last kernel; map2alm:
// T first, the polarisation alm leave in m pieces under the remaining analysis launches).
//
// Partition: the m axis is cut into ndev * S contiguous segments of equal MEASURED Legendre work (executed steps from the
// activation tables, spin 2 weighted 3x), boundaries on multiples of 8 m (128-byte runs in the phase rows), dealt to the
// GPUs boustrophedon -- equal work by construction, S contiguous alm ranges per GPU and component (S = PIXSHT_MULTI_SEGS,
// default 4).  Rings are split into contiguous slabs = contiguous rows of the caller's map.
//
// The same GPU may appear several times in `devices` (each entry is a separate shard with its own streams and buffers):
// that is how the single-GPU test box exercises this path.
#pragma once

struct MultiDev {
    int device = 0;
    pixsht_plan* sub = nullptr;          // ordinary plan on this GPU: tables, streams, staging buffers
    std::vector<int> m_list;             // this shard's m values, ascending
    int r0 = 0, r1 = 0;                  // band rings [r0, r1) of this shard's slab
    long long row_len = 0;               // phase row length of this shard's buffer (its m count rounded up to 8)
    DevBuf<int> d_m_list;
    DevBuf<long long> d_mtab;            // per m: address of (ring 0, comp 0, m) in its owner's phase buffer, owner's row length
    DevBuf<double2> d_phase;             // [nrings][3][row_len]
    DevBuf<unsigned char> d_slab[3];     // map rows of the slab, per component (host-pointer calls)
    cudaEvent_t e_stage[2] = {nullptr, nullptr};   // first stage of spin family slot 0 / 1 done on this shard
    cudaEvent_t e_t0 = nullptr, e_t1 = nullptr;
    std::vector<cudaEvent_t> pool; size_t npool = 0;
    cudaEvent_t next_ev()
    {
        if (npool == pool.size()) {
            cudaEvent_t e = nullptr;
            if (cudaEventCreateWithFlags(&e, trace ? cudaEventDefault : cudaEventDisableTiming) != cudaSuccess) return nullptr;
            pool.push_back(e);
        }
        return pool[npool++];
    }
    // PIXSHT_TRACE=1: a timeline of the shard's copies and stages (ms since the start of the call), printed to stderr after every
    // multi-GPU execute -- the tool that shows which copy a host-pointer call is waiting for
    bool trace = false;
    std::vector<std::pair<std::string, cudaEvent_t>> marks;
    void mark(cudaStream_t st, const char* fmt, int a = 0, int b = 0)
    {
        if (!trace) return;
        cudaEvent_t e = next_ev();
        if (!e || cudaEventRecord(e, st) != cudaSuccess) return;
        char buf[64];
        snprintf(buf, sizeof(buf), fmt, a, b);
        marks.emplace_back(buf, e);
    }
};

struct pixsht_multi {
    int ndev = 0;
    std::vector<MultiDev> dev;
    std::vector<double> work_m;          // Legendre work per m (partition weight)
    bool trace = false;                  // PIXSHT_TRACE
    int pieces = 5;                      // pieces of the first input / last output of a call (PIXSHT_MULTI_PIECES)
    int ring_pieces = 3;                 // alm2map: ring-range pieces of the last family's Legendre launches (PIXSHT_MULTI_RING_PIECES)
    int ring_pieces_anal = 2;            // map2alm: ring-range pieces of the last family's analysis (PIXSHT_MULTI_RING_PIECES_M2A): 2 = the equatorial
                                         // third of the rows for all m, then the rest in m pieces (3 equal pieces measured slower at 2 GPUs: the alm
                                         // can only leave after the last ring piece)
    int ring_pieces_first = 5;           // map2alm: ring-range pieces of a family that is not the last (T of IQU; PIXSHT_MULTI_RING_PIECES_T)
    double last_ms_device = 0;
};

#define MCU(call)                                                                                                     \
    do {                                                                                                              \
        cudaError_t e__ = (call);                                                                                     \
        if (e__ != cudaSuccess)                                                                                       \
            return fail(PIXSHT_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));                        \
    } while (0)

static pixsht_plan* multi_first_sub(pixsht_multi* M) { return M->dev[0].sub; }
static int multi_ndev(const pixsht_multi* M) { return M->ndev; }
static void multi_set_polconv(pixsht_multi* M, int iau) { for (auto& D : M->dev) if (D.sub) D.sub->polconv_iau = iau; }

// ---- partition -------------------------------------------------------------------------------------------------
// ndev * S contiguous segments of equal work, boundaries multiples of 8, dealt boustrophedon
static void multi_partition_m(const std::vector<double>& w, int ndev, int S, std::vector<std::vector<int>>& lists)
{
    const int n = (int)w.size();
    lists.assign(ndev, {});
    std::vector<double> cum(n + 1, 0.0);
    for (int m = 0; m < n; ++m) cum[m + 1] = cum[m] + std::max(w[m], 1e-30);
    const int nseg = std::max(1, std::min(ndev * S, (n + 7) / 8));
    std::vector<int> edge(nseg + 1, 0);
    for (int k = 1; k < nseg; ++k) {
        const double target = cum[n] * k / nseg;
        int m = (int)(std::lower_bound(cum.begin(), cum.end(), target) - cum.begin());
        m = (m + 4) / 8 * 8;
        edge[k] = std::min(n, std::max(edge[k - 1], m));
    }
    edge[nseg] = n;
    for (int k = 0; k < nseg; ++k) {
        const int round = k / ndev, pos = k % ndev;
        const int d = (round & 1) ? (ndev - 1 - pos) : pos;
        for (int m = edge[k]; m < edge[k + 1]; ++m) lists[d].push_back(m);
    }
    for (auto& l : lists) std::sort(l.begin(), l.end());
}

// contiguous alm index ranges [i0, i1) of rows [j0, j1) of a shard's m list (consecutive m are adjacent in the m-major alm)
template <class F>
static void multi_for_runs(const pixsht_plan* P, const MultiDev& D, int j0, int j1, F fn)
{
    int j = j0;
    while (j < j1) {
        int k = j;
        while (k + 1 < j1 && D.m_list[k + 1] == D.m_list[k] + 1) ++k;
        const int ma = D.m_list[j], mb = D.m_list[k];
        const long long i0 = alm_index(P->lmax, ma, ma), i1 = alm_index(P->lmax, P->lmax, mb) + 1;
        fn(i0, i1);
        j = k + 1;
    }
}

// rows [j0, j1) of a shard's m list in K pieces of about equal Legendre work
static std::vector<std::pair<int, int>> multi_m_pieces(const pixsht_multi* M, const MultiDev& D, int K)
{
    std::vector<std::pair<int, int>> out;
    const int nm = (int)D.m_list.size();
    if (nm == 0) return out;
    K = std::max(1, std::min(K, nm));
    std::vector<double> cum(nm + 1, 0.0);
    for (int j = 0; j < nm; ++j) cum[j + 1] = cum[j] + std::max(M->work_m[D.m_list[j]], 1e-30);
    int prev = 0;
    for (int k = 1; k <= K; ++k) {
        int e = (k == K) ? nm : (int)(std::lower_bound(cum.begin(), cum.end(), cum[nm] * k / K) - cum.begin());
        e = std::min(nm, std::max(prev, e));
        if (e > prev) out.push_back({prev, e});
        prev = e;
    }
    return out;
}

static void multi_destroy(pixsht_multi* M)
{
    if (!M) return;
    for (auto& D : M->dev) {
        (void)cudaSetDevice(D.device);
        D.d_m_list.release(); D.d_mtab.release(); D.d_phase.release();
        for (auto& b : D.d_slab) b.release();
        for (auto& e : D.e_stage) if (e) cudaEventDestroy(e);
        if (D.e_t0) cudaEventDestroy(D.e_t0);
        if (D.e_t1) cudaEventDestroy(D.e_t1);
        for (auto& e : D.pool) if (e) cudaEventDestroy(e);
        if (D.sub) pixsht_plan_destroy(D.sub);
    }
    (void)cudaGetLastError();
    delete M;
}

extern "C" int pixsht_plan_create_multi(pixsht_plan** out, const pixsht_geom* g, int lmax, int mmax, int dtype, int ndev, const int* devices)
{
    if (!out || !g || !devices) return fail(PIXSHT_ERR_ARG, "null argument");
    *out = nullptr;
    if (ndev < 1 || ndev > 64) return fail(PIXSHT_ERR_ARG, "need 1 <= ndev <= 64");
    pixsht_multi* M = new pixsht_multi();
    M->ndev = ndev;
    M->dev.resize(ndev);
    { const int v = env_int("PIXSHT_MULTI_PIECES", 5); M->pieces = (v >= 1 && v <= 16) ? v : 5; }
    M->trace = env_int("PIXSHT_TRACE", 0) != 0;
    { const int v = env_int("PIXSHT_MULTI_RING_PIECES", 3); M->ring_pieces = (v >= 1 && v <= 16) ? v : 3; }
    { const int v = env_int("PIXSHT_MULTI_RING_PIECES_M2A", 2); M->ring_pieces_anal = (v >= 1 && v <= 16) ? v : 2; }
    { const int v = env_int("PIXSHT_MULTI_RING_PIECES_T", 5); M->ring_pieces_first = (v >= 1 && v <= 16) ? v : 5; }
    auto bail = [&](int rc) { std::string keep = g_err; multi_destroy(M); g_err = keep; return rc; };
    for (int d = 0; d < ndev; ++d) {
        M->dev[d].device = devices[d];
        int rc = pixsht_plan_create(&M->dev[d].sub, g, lmax, mmax, dtype, devices[d]);
        if (rc) return bail(rc);
    }
    pixsht_plan* P0 = M->dev[0].sub;
    // ---- partition by measured work ----
    {
        std::vector<double> w0, w2;
        int rc = work_per_m(P0, 0, w0); if (rc) return bail(rc);
        rc = work_per_m(P0, 2, w2); if (rc) return bail(rc);
        M->work_m.resize(mmax + 1);
        for (int m = 0; m <= mmax; ++m) M->work_m[m] = w0[m] + 3.0 * w2[m];
        std::vector<std::vector<int>> lists;
        const int S = std::max(1, std::min(64, env_int("PIXSHT_MULTI_SEGS", 4)));
        multi_partition_m(M->work_m, ndev, S, lists);
        for (int d = 0; d < ndev; ++d) {
            MultiDev& D = M->dev[d];
            D.m_list = lists[d];
            D.row_len = std::max<long long>(8, ((long long)D.m_list.size() + 7) / 8 * 8);
            D.r0 = (int)((long long)P0->nrings * d / ndev);
            D.r1 = (int)((long long)P0->nrings * (d + 1) / ndev);
        }
    }
    // ---- peer access, phase buffers, tables ----
    for (int d = 0; d < ndev; ++d) {
        MultiDev& D = M->dev[d];
        if (cudaSetDevice(D.device) != cudaSuccess) return bail(fail(PIXSHT_ERR_CUDA, "cudaSetDevice failed"));
#ifndef PIXSHT_EMU
        for (int e = 0; e < ndev; ++e) {
            if (M->dev[e].device == D.device) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, D.device, M->dev[e].device) != cudaSuccess || !can)
                return bail(fail(PIXSHT_ERR_UNSUPPORTED, "the GPUs of a multi-GPU plan must have peer access to one another (NVLink / NVSwitch)"));
            const cudaError_t pe = cudaDeviceEnablePeerAccess(M->dev[e].device, 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) return bail(fail(PIXSHT_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(pe)));
            (void)cudaGetLastError();
        }
#endif
        if (D.d_phase.alloc((size_t)P0->nrings * 3 * D.row_len)) return bail(fail(PIXSHT_ERR_NOMEM, "phase buffer allocation failed"));
        if (D.d_m_list.upload(D.m_list)) return bail(fail(PIXSHT_ERR_NOMEM, "m list allocation failed"));
        for (auto& e : D.e_stage) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return bail(fail(PIXSHT_ERR_CUDA, "cudaEventCreate failed"));
        if (cudaEventCreate(&D.e_t0) != cudaSuccess || cudaEventCreate(&D.e_t1) != cudaSuccess) return bail(fail(PIXSHT_ERR_CUDA, "cudaEventCreate failed"));
        // both activation tables up front: the first transform should not pay for them
        int rc = ensure_seek(D.sub, 0, D.sub->stream); if (rc) return bail(rc);
        rc = ensure_seek(D.sub, 2, D.sub->stream); if (rc) return bail(rc);
    }
    {
        std::vector<long long> tab(2 * (size_t)(mmax + 1), 0);
        for (int d = 0; d < ndev; ++d) {
            const MultiDev& D = M->dev[d];
            for (size_t j = 0; j < D.m_list.size(); ++j) {
                tab[2 * (size_t)D.m_list[j]] = (long long)(uintptr_t)(D.d_phase.p + j);
                tab[2 * (size_t)D.m_list[j] + 1] = D.row_len;
            }
        }
        for (int d = 0; d < ndev; ++d) {
            MultiDev& D = M->dev[d];
            if (cudaSetDevice(D.device) != cudaSuccess) return bail(fail(PIXSHT_ERR_CUDA, "cudaSetDevice failed"));
            if (D.d_mtab.upload(tab)) return bail(fail(PIXSHT_ERR_NOMEM, "m table allocation failed"));
            if (cudaStreamSynchronize(D.sub->stream) != cudaSuccess) return bail(fail(PIXSHT_ERR_CUDA, "activation table build failed"));
        }
    }
    // the handle the caller sees: geometry only, no device state of its own
    pixsht_plan* P = new pixsht_plan();
    P->device = devices[0]; P->dtype = dtype; P->nphi = P0->nphi; P->nrings = P0->nrings; P->lmax = lmax; P->mmax = mmax;
    P->nx = P0->nx; P->ny = P0->ny; P->flipx = P0->flipx; P->flipy = P0->flipy; P->phi0 = P0->phi0; P->nalm = P0->nalm;
    P->npairs = P0->npairs; P->sm_count = P0->sm_count; P->nfft = P0->nfft; P->MP = P0->MP;
    P->R0 = P0->R0; P->R2 = P0->R2; P->R0a = P0->R0a; P->R2a = P0->R2a;
    P->h_theta = P0->h_theta; P->h_wgt = P0->h_wgt;
    P->multi = M;
    *out = P;
    return PIXSHT_OK;
}

// ---- execution ---------------------------------------------------------------------------------------------------
struct MultiFam { int spin; int cb, cn; };   // spin family and its components [cb, cb + cn) of the ncomp set

// `sharded`: alms[d * ncomp + c] / maps[d * ncomp + c] are device pointers on shard d's GPU (full-length alm of which only the
// shard's columns are used; the shard's slab of map rows) and no copies are made.  Otherwise alms[c] / maps[c] are the caller's
// whole arrays (host memory, or any memory the GPUs can copy from: cudaMemcpyDefault).
static int execute_multi(pixsht_plan* P, int direction, int ncomp, void* const* alms, void* const* maps, bool sharded)
{
    pixsht_multi* M = P->multi;
    const int nd = M->ndev;
    const bool f32 = P->dtype == PIXSHT_F32;
    const size_t esz = f32 ? 4 : 8;
    const size_t alm_bytes = (size_t)P->nalm * 2 * esz;
    const auto t_begin = std::chrono::steady_clock::now();
    std::vector<MultiFam> fams;
    if (ncomp == 1) fams = {{0, 0, 1}};
    else if (ncomp == 2) fams = {{2, 0, 2}};
    else fams = {{0, 0, 1}, {2, 1, 2}};   // T first in both directions: its copies and stages are short, the polarisation pipelines behind it
    const int nf = (int)fams.size();

    // per shard: device-side views
    struct View { void* dalm[3]; double2* dalm64[3]; void* dmap_virtual[3]; void* dslab[3]; size_t row_off; size_t slab_bytes; };
    std::vector<View> V(nd);
    for (int d = 0; d < nd; ++d) {
        MultiDev& D = M->dev[d];
        pixsht_plan* S = D.sub;
        MCU(cudaSetDevice(D.device));
        D.npool = 0; S->launches = 0; D.trace = M->trace; D.marks.clear();
        View& v = V[d];
        size_t off, nb; ring_rows(P, D.r0, D.r1, esz, off, nb);
        v.row_off = off; v.slab_bytes = nb;
        for (int c = 0; c < ncomp; ++c) {
            if (sharded) { v.dalm[c] = alms[(size_t)d * ncomp + c]; v.dslab[c] = maps[(size_t)d * ncomp + c]; }
            else {
                if (S->d_alm[c].n < alm_bytes && S->d_alm[c].alloc(alm_bytes)) return fail(PIXSHT_ERR_NOMEM, "alm staging allocation failed");
                if (D.d_slab[c].n < std::max<size_t>(nb, 16) && D.d_slab[c].alloc(std::max<size_t>(nb, 16))) return fail(PIXSHT_ERR_NOMEM, "map staging allocation failed");
                v.dalm[c] = S->d_alm[c].p; v.dslab[c] = D.d_slab[c].p;
            }
            if (f32) {
                if (S->d_alm64[c].n < (size_t)P->nalm && S->d_alm64[c].alloc(P->nalm)) return fail(PIXSHT_ERR_NOMEM, "alm work buffer allocation failed");
                v.dalm64[c] = S->d_alm64[c].p;
            } else v.dalm64[c] = reinterpret_cast<double2*>(v.dalm[c]);
            // the FFT kernels address rows of the FULL map: hand them the address the full map would start at
            v.dmap_virtual[c] = (char*)v.dslab[c] - off;
        }
        MCU(cudaEventRecord(D.e_t0, S->stream));
        cudaEvent_t e = D.next_ev();
        MCU(cudaEventRecord(e, S->stream));
        MCU(cudaStreamWaitEvent(S->s_h2d, e, 0));
        MCU(cudaStreamWaitEvent(S->s_d2h, e, 0));
    }
    const int K = sharded ? 1 : M->pieces;
    int rc;
    bool pg_alm[3] = {false, false, false}, pg_map[3] = {false, false, false};   // pageable arrays: staged by each shard's copy threads
    if (!sharded) for (int c = 0; c < ncomp; ++c) { pg_alm[c] = host_is_pageable(alms[c]); pg_map[c] = host_is_pageable(maps[c]); }

    // Ring-range pieces of a family's Legendre launches (host-pointer calls): chunk range [c0, c1) of the spin family's kernel and
    // the band rings it covers, north [rn0, rn1) and south [rs0, rs1).  With them the rows of a family leave (alm2map) or arrive
    // (map2alm) piece by piece under the Legendre launches of the other pieces, at the price of one cross-GPU barrier per piece.
    struct RingPiece { int c0, c1, rn0, rn1, rs0, rs1; };
    auto ring_pieces = [&](int spin, bool anal, int Q) -> std::vector<RingPiece> {
        pixsht_plan* S0 = M->dev[0].sub;
        const int R = leg_R(S0, spin, anal), nch = leg_total_chunks(S0, R);
        Q = std::max(1, std::min(Q, nch));
        std::vector<int> cb(Q + 1);
        for (int k = 0; k <= Q; ++k) {
            // chunks are numbered from the pole to the equator.  Synthesis: equal Legendre work (~ sin theta per pair), polar side
            // first, so that the last (exposed) piece of the output has the fewest rows; analysis: equal row counts.
            // Analysis with two pieces: the equatorial third of the rows (half of the Legendre work) and the polar rest -- the first is
            // analysed for all m while the second is still on the wire, the second in m pieces whose alm columns leave one by one.
            cb[k] = anal ? (Q == 2 ? (k == 1 ? (int)std::lround(nch * 2.0 / 3.0) : (int)((long long)nch * k / Q)) : (int)((long long)nch * k / Q))
                         : (int)std::lround(nch * (2.0 / 3.14159265358979323846) * std::acos(1.0 - (double)k / Q));
        }
        cb[0] = 0; cb[Q] = nch;
        for (int k = 1; k <= Q; ++k) cb[k] = std::max(cb[k], cb[k - 1]);
        std::vector<RingPiece> out;
        for (int k = 0; k < Q; ++k) {
            if (cb[k + 1] <= cb[k]) continue;
            int rng[2][2];
            if (!pair_range_rings(S0, std::min(S0->npairs, cb[k] * 32 * R), std::min(S0->npairs, cb[k + 1] * 32 * R), rng))
                return {{0, nch, 0, P->nrings, 0, 0}};
            out.push_back({cb[k], cb[k + 1], rng[0][0], rng[0][1], rng[1][0], rng[1][1]});
        }
        if (anal) std::reverse(out.begin(), out.end());   // analysis: equator (most work per row) first, polar caps last
        if (out.empty()) out.push_back({0, nch, 0, P->nrings, 0, 0});
        return out;
    };
    // the part of band rings [ra, rb) that lies in shard d's slab
    auto clip = [&](const MultiDev& D, int ra, int rb, int& a, int& b) { a = std::max(ra, D.r0); b = std::min(rb, D.r1); return b > a; };
    const int Q = sharded ? 1 : (direction == PIXSHT_ALM2MAP ? M->ring_pieces : M->ring_pieces_anal);

    if (direction == PIXSHT_ALM2MAP) {
        // ---- inputs: each shard's alm columns, family by family (T first); the first family in m pieces of equal work ----
        std::vector<std::vector<std::vector<std::pair<int, int>>>> segs(nd, std::vector<std::vector<std::pair<int, int>>>(nf));
        std::vector<std::vector<std::vector<cudaEvent_t>>> ev_in(nd, std::vector<std::vector<cudaEvent_t>>(nf));
        for (int d = 0; d < nd; ++d) {
            MultiDev& D = M->dev[d];
            pixsht_plan* S = D.sub;
            MCU(cudaSetDevice(D.device));
            for (int fi = 0; fi < nf; ++fi) {
                segs[d][fi] = multi_m_pieces(M, D, fi == 0 ? K : 1);
                for (auto& sg : segs[d][fi]) {
                    if (!sharded)
                        for (int c = fams[fi].cb; c < fams[fi].cb + fams[fi].cn; ++c) {
                            cudaError_t ce = cudaSuccess;
                            multi_for_runs(P, D, sg.first, sg.second, [&](long long i0, long long i1) {
                                if (ce == cudaSuccess)
                                    ce = host_copy_in(S, pg_alm[c], (char*)V[d].dalm[c] + (size_t)i0 * 2 * esz, (const char*)alms[c] + (size_t)i0 * 2 * esz,
                                                      (size_t)(i1 - i0) * 2 * esz, S->s_h2d);
                            });
                            MCU(ce);
                        }
                    cudaEvent_t e = D.next_ev();
                    MCU(cudaEventRecord(e, S->s_h2d));
                    ev_in[d][fi].push_back(e);
                    D.mark(S->s_h2d, "h2d alm fam%d piece%d", fi, (int)ev_in[d][fi].size() - 1);
                }
            }
        }
        // rows [ra, rb) of family F on shard d: FFT (every m fetched from its owner's phase buffer), then the copy out
        auto emit_rows = [&](int d, const MultiFam& F, int ra, int rb) -> int {
            MultiDev& D = M->dev[d];
            pixsht_plan* S = D.sub;
            int a, b;
            if (!clip(D, ra, rb, a, b)) return PIXSHT_OK;
            int rc2 = stage_fft(S, PIXSHT_ALM2MAP, ncomp, F.cb, F.cn, nullptr, a, b - a, V[d].dmap_virtual, S->stream, D.d_mtab.p);
            if (rc2) return rc2;
            D.mark(S->stream, "fft rows %d..%d", a, b);
            if (!sharded) {
                cudaEvent_t e = D.next_ev();
                MCU(cudaEventRecord(e, S->stream));
                MCU(cudaStreamWaitEvent(S->s_d2h, e, 0));
                size_t off, nb; ring_rows(P, a, b, esz, off, nb);
                for (int c = F.cb; c < F.cb + F.cn; ++c)
                    MCU(host_copy_out(S, pg_map[c], (char*)maps[c] + off, (char*)V[d].dslab[c] + (off - V[d].row_off), nb, S->s_d2h));
                D.mark(S->s_d2h, "d2h rows %d..%d", a, b);
            }
            return PIXSHT_OK;
        };
        for (int fi = 0; fi < nf; ++fi) {
            const MultiFam& F = fams[fi];
            // the last family leaves in ring-range pieces (its rows go out under its own Legendre launches); the others in one
            const std::vector<RingPiece> rp = ring_pieces(F.spin, false, fi == nf - 1 ? Q : 1);
            // ---- inputs of the family on each shard: conversion + record preparation of its m values as they arrive ----
            for (int d = 0; d < nd; ++d) {
                MultiDev& D = M->dev[d];
                pixsht_plan* S = D.sub;
                MCU(cudaSetDevice(D.device));
                cudaStream_t sc = S->stream;
                for (size_t k = 0; k < segs[d][fi].size(); ++k) {
                    const int j0 = segs[d][fi][k].first, j1 = segs[d][fi][k].second;
                    MCU(cudaStreamWaitEvent(sc, ev_in[d][fi][k], 0));
                    const int* ml = D.d_m_list.p + j0;
                    if (f32) {
                        const dim3 grid((unsigned)std::max(1, std::min(8, (2 * P->lmax + 512) / 256)), (unsigned)(j1 - j0));
                        for (int c = F.cb; c < F.cb + F.cn; ++c) {
                            PIXSHT_LAUNCH((k_cvt_rows<float, double, false>), grid, 256, 0, sc, ml, P->lmax, (const float*)V[d].dalm[c], (double*)V[d].dalm64[c]);
                            S->launches++;
                        }
                    }
                    rc = synth_prep(S, F.spin, V[d].dalm64[F.cb], V[d].dalm64[F.cb + F.cn - 1], sc, 0, -1, ml, j1 - j0); if (rc) return rc;
                    if (rp.size() == 1) {
                        // Legendre stage of this m piece, all rings
                        const LegJob J = {F.spin, ncomp, F.cb, 0, j1 - j0, ml, 0, leg_total_chunks(S, leg_R(S, F.spin, false)),
                                          {D.d_phase.p + j0, D.row_len, 1}};
                        rc = synth_launch(S, J, sc); if (rc) return rc;
                        D.mark(sc, "legendre fam%d mpiece%d", fi, (int)k);
                    }
                }
            }
            for (size_t q = 0; q < rp.size(); ++q) {
                std::vector<cudaEvent_t> e_q(nd, nullptr);
                for (int d = 0; d < nd; ++d) {
                    MultiDev& D = M->dev[d];
                    pixsht_plan* S = D.sub;
                    MCU(cudaSetDevice(D.device));
                    if (rp.size() > 1 && !D.m_list.empty()) {
                        // Legendre stage of the shard's m values on the ring pairs of this piece
                        const LegJob J = {F.spin, ncomp, F.cb, 0, (int)D.m_list.size(), D.d_m_list.p, rp[q].c0, rp[q].c1 - rp[q].c0,
                                          {D.d_phase.p, D.row_len, 1}};
                        rc = synth_launch(S, J, S->stream); if (rc) return rc;
                        D.mark(S->stream, "legendre fam%d ringpiece%d", fi, (int)q);
                    }
                    e_q[d] = D.next_ev();
                    MCU(cudaEventRecord(e_q[d], S->stream));
                }
                // ---- barrier across the shards, then the ring FFTs of the piece's rows in each shard's slab ----
                for (int d = 0; d < nd; ++d) {
                    MultiDev& D = M->dev[d];
                    MCU(cudaSetDevice(D.device));
                    int a, b;
                    const bool any = clip(D, rp[q].rn0, rp[q].rn1, a, b) || clip(D, rp[q].rs0, rp[q].rs1, a, b);
                    if (!any) continue;
                    for (int e = 0; e < nd; ++e) if (e != d) MCU(cudaStreamWaitEvent(D.sub->stream, e_q[e], 0));
                    rc = emit_rows(d, F, rp[q].rn0, rp[q].rn1); if (rc) return rc;
                    rc = emit_rows(d, F, rp[q].rs0, rp[q].rs1); if (rc) return rc;
                }
            }
        }
    } else {
        // ---- inputs: each shard's rows, family by family (T first).  The first family arrives in plain pieces of its slab, the
        // last one in the ring-range pieces of its analysis launches (equator first) ----
        struct RingSeg { int ra, rb; cudaEvent_t ev; };
        std::vector<std::vector<RingPiece>> rps(nf);
        std::vector<std::vector<std::vector<std::vector<RingSeg>>>> rsegs(nd);   // [d][fi][piece] -> row ranges of the shard
        // the last family: Q pieces (default 2: equatorial third for all m, then the rest in m pieces); a family before it (T of an IQU
        // set): ring pieces too, so that its analysis starts when the first third of its rows is in (trace at 2 GPUs: the call waited
        // 17 ms for the T rows) -- its alm leave in one go under the next family's stages
        for (int fi = 0; fi < nf; ++fi) rps[fi] = ring_pieces(fams[fi].spin, true, sharded ? 1 : (fi == nf - 1 ? Q : M->ring_pieces_first));
        for (int d = 0; d < nd; ++d) {
            MultiDev& D = M->dev[d];
            pixsht_plan* S = D.sub;
            MCU(cudaSetDevice(D.device));
            rsegs[d].resize(nf);
            const int nloc = D.r1 - D.r0;
            for (int fi = 0; fi < nf; ++fi) {
                rsegs[d][fi].resize(rps[fi].size());
                for (size_t q = 0; q < rps[fi].size(); ++q) {
                    std::vector<std::pair<int, int>> ranges;
                    if (rps[fi].size() == 1) {
                        const int kk = std::max(1, std::min(fi == 0 ? K : 1, std::max(nloc, 1)));
                        for (int k = 0; k < kk; ++k) ranges.push_back({D.r0 + (int)((long long)nloc * k / kk), D.r0 + (int)((long long)nloc * (k + 1) / kk)});
                    } else {
                        int a, b;
                        if (clip(D, rps[fi][q].rn0, rps[fi][q].rn1, a, b)) ranges.push_back({a, b});
                        if (clip(D, rps[fi][q].rs0, rps[fi][q].rs1, a, b)) ranges.push_back({a, b});
                    }
                    for (auto& rg : ranges) {
                        if (rg.second <= rg.first) continue;
                        if (!sharded) {
                            size_t off, nb; ring_rows(P, rg.first, rg.second, esz, off, nb);
                            for (int c = fams[fi].cb; c < fams[fi].cb + fams[fi].cn; ++c)
                                MCU(host_copy_in(S, pg_map[c], (char*)V[d].dslab[c] + (off - V[d].row_off), (const char*)maps[c] + off, nb, S->s_h2d));
                        }
                        cudaEvent_t e = D.next_ev();
                        MCU(cudaEventRecord(e, S->s_h2d));
                        rsegs[d][fi][q].push_back({rg.first, rg.second, e});
                        D.mark(S->s_h2d, "h2d rows %d..%d", rg.first, rg.second);
                    }
                }
            }
        }
        for (int fi = 0; fi < nf; ++fi) {
            const MultiFam& F = fams[fi];
            const std::vector<RingPiece>& rp = rps[fi];
            for (size_t q = 0; q < rp.size(); ++q) {
                // ---- ring FFTs of the piece's rows on each shard; every m goes to its owner's phase buffer ----
                std::vector<cudaEvent_t> e_q(nd, nullptr);
                for (int d = 0; d < nd; ++d) {
                    MultiDev& D = M->dev[d];
                    pixsht_plan* S = D.sub;
                    MCU(cudaSetDevice(D.device));
                    cudaStream_t sc = S->stream;
                    if (q == 0 && !D.m_list.empty()) {
                        // the analysis kernels accumulate atomically (also across the ring pieces): zero the shard's columns once
                        const dim3 grid((unsigned)std::max(1, std::min(8, (2 * P->lmax + 512) / 256)), (unsigned)D.m_list.size());
                        for (int c = F.cb; c < F.cb + F.cn; ++c) {
                            PIXSHT_LAUNCH((k_cvt_rows<double, double, true>), grid, 256, 0, sc, D.d_m_list.p, P->lmax, (const double*)nullptr, (double*)V[d].dalm64[c]);
                            S->launches++;
                        }
                    }
                    for (auto& sg : rsegs[d][fi][q]) {
                        MCU(cudaStreamWaitEvent(sc, sg.ev, 0));
                        rc = stage_fft(S, PIXSHT_MAP2ALM, ncomp, F.cb, F.cn, nullptr, sg.ra, sg.rb - sg.ra, V[d].dmap_virtual, sc, D.d_mtab.p); if (rc) return rc;
                        D.mark(sc, "fft rows %d..%d", sg.ra, sg.rb);
                    }
                    e_q[d] = D.next_ev();
                    MCU(cudaEventRecord(e_q[d], sc));
                }
                // ---- barrier, then the Legendre analysis of the shard's own m values on the piece's ring pairs; on the last piece
                // of the last family the m values are done in pieces whose alm columns leave one by one ----
                const bool last = (q + 1 == rp.size());
                for (int d = 0; d < nd; ++d) {
                    MultiDev& D = M->dev[d];
                    pixsht_plan* S = D.sub;
                    MCU(cudaSetDevice(D.device));
                    cudaStream_t sc = S->stream;
                    const int nm = (int)D.m_list.size();
                    if (nm == 0) continue;
                    for (int e = 0; e < nd; ++e) if (e != d) MCU(cudaStreamWaitEvent(sc, e_q[e], 0));
                    const int nch_all = leg_total_chunks(S, leg_R(S, F.spin, true));
                    const int cq0 = rp.size() == 1 ? 0 : rp[q].c0, cq1 = rp.size() == 1 ? nch_all : rp[q].c1;
                    const auto pieces = multi_m_pieces(M, D, (last && fi == nf - 1) ? K : 1);
                    for (auto& pc : pieces) {
                        const int j0 = pc.first, j1 = pc.second;
                        const int* ml = D.d_m_list.p + j0;
                        const LegJob J = {F.spin, ncomp, F.cb, 0, j1 - j0, ml, cq0, cq1 - cq0, {D.d_phase.p + j0, D.row_len, 1}};
                        rc = anal_launch(S, J, V[d].dalm64[F.cb], F.cn == 2 ? V[d].dalm64[F.cb + 1] : nullptr, sc); if (rc) return rc;
                        D.mark(sc, "legendre fam%d ringpiece%d", fi, (int)q);
                        if (!last) continue;
                        if (f32) {
                            const dim3 grid((unsigned)std::max(1, std::min(8, (2 * P->lmax + 512) / 256)), (unsigned)(j1 - j0));
                            for (int c = F.cb; c < F.cb + F.cn; ++c) {
                                PIXSHT_LAUNCH((k_cvt_rows<double, float, false>), grid, 256, 0, sc, ml, P->lmax, (const double*)V[d].dalm64[c], (float*)V[d].dalm[c]);
                                S->launches++;
                            }
                        }
                        if (!sharded) {
                            cudaEvent_t e = D.next_ev();
                            MCU(cudaEventRecord(e, sc));
                            MCU(cudaStreamWaitEvent(S->s_d2h, e, 0));
                            for (int c = F.cb; c < F.cb + F.cn; ++c) {
                                cudaError_t ce = cudaSuccess;
                                multi_for_runs(P, D, j0, j1, [&](long long i0, long long i1) {
                                    if (ce == cudaSuccess)
                                        ce = host_copy_out(S, pg_alm[c], (char*)alms[c] + (size_t)i0 * 2 * esz, (const char*)V[d].dalm[c] + (size_t)i0 * 2 * esz,
                                                           (size_t)(i1 - i0) * 2 * esz, S->s_d2h);
                                });
                                MCU(ce);
                            }
                            D.mark(S->s_d2h, "d2h alm fam%d m %d..", fi, j0);
                        }
                    }
                }
            }
        }
    }
    // ---- completion: the call returns when every shard's output is in place ----
    P->launches = 0;
    double dev_ms = 0;
    int status = PIXSHT_OK;
    for (int d = 0; d < nd; ++d) {
        MultiDev& D = M->dev[d];
        pixsht_plan* S = D.sub;
        if (cudaSetDevice(D.device) != cudaSuccess) { status = fail(PIXSHT_ERR_CUDA, "cudaSetDevice failed"); continue; }
        cudaError_t e1 = cudaEventRecord(D.e_t1, S->stream);
        cudaError_t e2 = cudaStreamSynchronize(S->s_h2d), e3 = cudaStreamSynchronize(S->stream), e4 = cudaStreamSynchronize(S->s_d2h);
        cudaError_t e5 = cudaGetLastError();
        for (cudaError_t e : {e1, e2, e3, e4, e5})
            if (e != cudaSuccess && status == PIXSHT_OK) status = fail(PIXSHT_ERR_CUDA, std::string("multi-GPU transform, shard ") + std::to_string(d) + ": " + cudaGetErrorString(e));
        if (status == PIXSHT_OK) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, D.e_t0, D.e_t1) == cudaSuccess) dev_ms = std::max(dev_ms, (double)ms);
            (void)cudaGetLastError();
        }
        host_copies_done(S);
        P->launches += S->launches;
        if (D.trace && status == PIXSHT_OK) {
            std::vector<std::pair<float, std::string>> tl;
            for (auto& mk : D.marks) { float ms = 0; if (cudaEventElapsedTime(&ms, D.e_t0, mk.second) == cudaSuccess) tl.emplace_back(ms, mk.first); }
            (void)cudaGetLastError();
            std::sort(tl.begin(), tl.end());
            fprintf(stderr, "[pixsht trace] %s shard %d (GPU %d):", direction == PIXSHT_ALM2MAP ? "alm2map" : "map2alm", d, D.device);
            for (auto& t : tl) fprintf(stderr, " | %.1f %s", t.first, t.second.c_str());
            fprintf(stderr, "\n");
        }
    }
    for (auto& t : P->timings) t = 0;
    P->timings[5] = dev_ms;
    P->timings[4] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    return status;
}

// every stream of every shard idle (error paths: nothing may still reference the caller's buffers when the call returns)
static void multi_quiesce(pixsht_multi* M)
{
    for (auto& D : M->dev) {
        if (cudaSetDevice(D.device) != cudaSuccess) continue;
        (void)cudaStreamSynchronize(D.sub->s_h2d); (void)cudaStreamSynchronize(D.sub->stream); (void)cudaStreamSynchronize(D.sub->s_d2h);
        host_copies_done(D.sub);
    }
    (void)cudaGetLastError();
}

extern "C" int pixsht_execute_sharded(pixsht_plan* P, int direction, int ncomp, void* const* alms, void* const* maps)
{
    if (!P || !alms || !maps) return fail(PIXSHT_ERR_ARG, "null argument");
    if (!P->multi) return fail(PIXSHT_ERR_ARG, "pixsht_execute_sharded needs a plan made by pixsht_plan_create_multi");
    if (direction != PIXSHT_MAP2ALM && direction != PIXSHT_ALM2MAP) return fail(PIXSHT_ERR_ARG, "bad direction");
    if (ncomp < 1 || ncomp > 3) return fail(PIXSHT_ERR_ARG, "SHTs require 1 <= ncomp <= 3, for I, QU, and IQU.");
    for (int i = 0; i < ncomp * P->multi->ndev; ++i) if (!alms[i] || !maps[i]) return fail(PIXSHT_ERR_ARG, "null component pointer");
    std::lock_guard<std::mutex> lock(P->mu);
    const int rc = execute_multi(P, direction, ncomp, alms, maps, true);
    if (rc) { std::string keep = g_err; multi_quiesce(P->multi); g_err = keep; }
    return rc;
}

extern "C" int pixsht_multi_shard(const pixsht_plan* P, int shard, int32_t info[4], int32_t* m_list)
{
    if (!P || !P->multi || !info) return fail(PIXSHT_ERR_ARG, "not a multi-GPU plan");
    if (shard < 0 || shard >= P->multi->ndev) return fail(PIXSHT_ERR_ARG, "shard index out of range");
    const MultiDev& D = P->multi->dev[shard];
    info[0] = D.device; info[1] = D.r0; info[2] = D.r1 - D.r0; info[3] = (int32_t)D.m_list.size();
    if (m_list) for (size_t j = 0; j < D.m_list.size(); ++j) m_list[j] = D.m_list[j];
    return PIXSHT_OK;
}

// A batch on a multi-GPU plan: whole transforms are independent, so the members are dealt to the GPUs in contiguous shares
// and every GPU runs its share through its own single-GPU plan (one host thread per GPU: the host-pointer path blocks).
static int execute_batch_multi(pixsht_plan* P, int direction, int nbatch, void* const* alms, void* const* maps, int location)
{
    pixsht_multi* M = P->multi;
    if (location != PIXSHT_HOST) return fail(PIXSHT_ERR_UNSUPPORTED, "batches on a multi-GPU plan take host pointers");
    const int nd = M->ndev;
    std::vector<int> rcs(nd, PIXSHT_OK);
    std::vector<std::string> errs(nd);
    std::vector<std::thread> th;
    const auto t_begin = std::chrono::steady_clock::now();
    for (int d = 0; d < nd; ++d) {
        const int b0 = (int)((long long)nbatch * d / nd), b1 = (int)((long long)nbatch * (d + 1) / nd);
        if (b1 <= b0) continue;
        auto job = [&, d, b0, b1]() {
            rcs[d] = pixsht_execute_batch(M->dev[d].sub, direction, b1 - b0, alms + b0, maps + b0, PIXSHT_HOST);
            if (rcs[d]) errs[d] = pixsht_last_error();
        };
#ifdef PIXSHT_EMU
        job();                   // the host emulation of the kernels is not re-entrant
#else
        th.emplace_back(job);
#endif
    }
    for (auto& t : th) t.join();
    P->launches = 0;
    for (int d = 0; d < nd; ++d) P->launches += M->dev[d].sub->launches;
    for (auto& t : P->timings) t = 0;
    P->timings[4] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    for (int d = 0; d < nd; ++d) if (rcs[d]) return fail(rcs[d], errs[d]);
    return PIXSHT_OK;
}
