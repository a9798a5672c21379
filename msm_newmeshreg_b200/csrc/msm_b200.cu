// msm_b200.cu — context + extern "C" layer (include/msm_b200.h) over the sm_100a kernels in kernels.cuh.
// Host code is C++; there is no CPU fallback: every compute entry point launches CUDA kernels or fails.
#include "../../include/msm_b200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <array>

#include "ico_hierarchy.h"
#include "kernels.cuh"
#include "meshwarp.cuh"
#include "triplet.cuh"
#include "patchfast.cuh"
#include "features.cuh"

using namespace msm;

namespace {

// cudaFuncAttributeMaxDynamicSharedMemorySize is per (function, device), not per context: contexts driven from different
// threads hold this lock from setting it to launching, so that a smaller request of another context cannot slip in between.
// Only launches that need more than the default 48 KB take it.
std::mutex g_big_smem_mtx;
struct BigSmemLock {
    std::unique_lock<std::mutex> lk;
    explicit BigSmemLock(size_t smem) : lk(g_big_smem_mtx, std::defer_lock) { if (smem > 48 * 1024) lk.lock(); }
};
constexpr int kUncCap = 1 << 20;   // boundary-tie pairs one patch build can hand to the host
thread_local std::string g_create_error;   // per thread: contexts may be created concurrently

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    ~PinBuf() { if (p) cudaFreeHost(p); }
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes + 256);
        if (e == cudaSuccess) cap = bytes + 256;
        return e;
    }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace

struct msm_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::string err;
    long long launches = 0;
    int timing_mode = 0;  // 0 off, 1 synchronous, 2 asynchronous (collected later)
    std::map<std::string, float> last_ms;
    struct PendingEv { const char* name; cudaEvent_t e0, e1; };
    std::vector<PendingEv> ev_pending;
    std::vector<cudaEvent_t> ev_free;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    msm_params par;

    // target
    IcoHierarchy ico;
    bool have_target = false;
    int Vt = 0, Ft = 0, Ct = 0;
    DevBuf d_ico_base, d_ntab, d_ntab_d, d_mtab, d_leaf_vid, d_leaf_user, d_leaf_perm, d_tfeat, d_face_feat;
    int face_feat_cq = -1;  // layout the face-packed features were built for (-1 none, 0 univariate, q quads)
    int hint_level = -1;    // level the hint tables below were built for
    DevBuf d_hdual, d_hnbr, d_hT;
    double max_label_disp = 0.0;  // max |label - centre| (sizes the hint level)

    // source
    int Vs = 0, Cs = 0, wrows = 0;
    bool multivariate = false;
    PinBuf h_src;            // pinned host copy of the source coordinates (upload staging + exact tie resolution)
    double src_bound = 0.0;  // max |coordinate| of the source mesh (sizes the patch voxel grid)
    DevBuf d_src, d_sxf, d_syf, d_szf, d_sfeat, d_swgt;

    // control-point grid
    int N = 0, Tcp = 0;
    std::vector<double> h_cp, h_maxsep;
    DevBuf d_cp, d_ocp, d_orig, d_maxsep, d_absw, d_tris, d_vt_off, d_vt_idx;
    int k_begin = 0, k_end = 0;

    // labels
    int L = 0;
    double centre[3] = {0, 0, 0};
    DevBuf d_labels, d_disp, d_Kf, d_Kd;
    bool rot_dirty = true;

    // cliques
    int P = 0, Tn = 0;
    DevBuf d_pairs, d_trip;
    int S = 1, N_subj = 0, T_subj = 0;
    long long n_maps = 0;            // group patch maps (set_patch_data)
    DevBuf d_moff, d_mkey, d_mval;

    // patches
    bool have_patches = false;
    long long total = 0;
    int maxP = 0;
    DevBuf d_counts, d_off, d_idx, d_unc, d_unc_count, d_acc_off, d_acc_idx, d_totmax;
    DevBuf d_px, d_py, d_pz, d_pa, d_pw;
    DevBuf d_scan_sums;       // chunk sums of the voxel-grid scan
    DevBuf d_slab;            // [Ns][slab_cap] members stashed by the count pass (one-pass patch construction)
    int slab_cap = 0;         // capacity per control point for the NEXT build (from the last build's largest patch); 0 = off
    DevBuf d_vkey;            // [Vs] spatial sort key of every source vertex (k_vertex_key)
    bool vkey_valid = false;
    DevBuf d_cell_counts, d_cell_start, d_cell_fill, d_cell_of, d_sorted;  // voxel grid of the source vertices
    DevBuf d_pcx, d_pcy, d_pcz, d_hint_face, d_lab12, d_tref;  // k_unary_setup outputs
    int pack_cp = 0;
    bool patches_padded = false;   // packed streams are fixed-capacity rows [Ns][pad_cap] (one-pass build, patchfast.cuh), not CSR
    int pad_cap = 0;
    DevBuf d_pflags, d_end;
    long long fast_builds = 0, exact_builds = 0;
    // deferred check of the one-pass build (msm_build_patches_async): the flag word travels to `h_pflags` behind the
    // kernel, `ev_patch` marks its arrival; until msm_patches_resolve has read it, maxP / total are the row capacity's bounds
    bool patch_pending = false;
    cudaEvent_t ev_patch = nullptr;
    PinBuf h_pflags;
    BinGrid pend_bg{};
    bool pend_binned = false;

    DevBuf d_w_start, d_w_items, d_w_items2, d_w_counts, d_w_fill, d_w_from, d_w_to, d_w_tris, d_w_pts, d_w_out;  // msm_mesh_interpolate
    // feature pre-processing (features.cuh): scratch of its own, so that a deferred patch build's voxel grid stays intact
    DevBuf d_f_in, d_f_ref, d_f_vals, d_f_mask, d_f_mask2, d_f_out, d_f_stats, d_f_cdf;
    DevBuf d_f_cell_counts, d_f_cell_start, d_f_cell_fill, d_f_cell_of, d_f_sorted, d_f_sorted2, d_f_scan;

    // outputs / scratch
    DevBuf d_unary, d_scratch, d_scratch2, d_tprep;
    PinBuf h_pin;
    std::vector<float> h_unary;   // last unary table, shard rows [L][Ns], for msm_total_cost
    int h_unary_k0 = 0, h_unary_ns = 0;
    bool have_unary = false;
};

#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                             \
            return 1;                                                                                  \
        }                                                                                              \
    } while (0)
#define FAIL(msg)          \
    do {                   \
        ctx->err = (msg);  \
        return 1;          \
    } while (0)
#define KLAUNCH_CHECK()                     \
    do {                                    \
        ctx->launches++;                    \
        CK(cudaGetLastError());             \
    } while (0)

namespace {

// Per-kernel-family timing with CUDA events on the launching stream.  mode 1: synchronous (event pair, sync, store the
// last duration).  mode 2: asynchronous — event pairs are only recorded; msm_timing_collect() reads them after the
// caller's own synchronisation, so a timed region is not perturbed by host syncs.
struct Timer {
    msm_ctx* ctx; const char* name;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    Timer(msm_ctx* c, const char* n) : ctx(c), name(n) {
        if (ctx->timing_mode != 0) {
            if (ctx->ev_free.size() < 2) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); ctx->ev_free.push_back(a); ctx->ev_free.push_back(b); }
            e0 = ctx->ev_free.back(); ctx->ev_free.pop_back();
            e1 = ctx->ev_free.back(); ctx->ev_free.pop_back();
            cudaEventRecord(e0, ctx->stream);
        }
    }
    ~Timer() {
        if (!e0) return;
        if (ctx->timing_mode == 1) {
            cudaEventRecord(e1, ctx->stream);
            cudaEventSynchronize(e1);
            float ms = -1.f;
            cudaEventElapsedTime(&ms, e0, e1);
            ctx->last_ms[name] = ms;
            ctx->ev_free.push_back(e0); ctx->ev_free.push_back(e1);
        } else {
            cudaEventRecord(e1, ctx->stream);
            ctx->ev_pending.push_back({name, e0, e1});
        }
    }
};

DevParams dev_params(const msm_ctx* ctx) {
    DevParams d;
    const msm_params& p = ctx->par;
    d.lambda = p.lambda; d.rexp = p.rexp; d.mu = p.mu; d.kappa = p.kappa; d.k_exp = p.k_exp; d.range = p.range;
    d.meanangle = p.meanangle; d.mvdmax = p.mvdmax; d.simmeasure = p.simmeasure; d.rmode = p.rmode; d.dweight = p.dweight;
    d.anorm = p.anorm; d.group = p.group;
    return d;
}

template <typename T>
int upload(msm_ctx* ctx, DevBuf& b, const T* h, size_t n) {
    CK(b.ensure(std::max<size_t>(n, 1) * sizeof(T)));
    if (n) CK(cudaMemcpyAsync(b.p, h, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

int upload_as_float(msm_ctx* ctx, DevBuf& b, const double* h, size_t n) {
    CK(ctx->h_pin.ensure(n * sizeof(float)));
    float* f = ctx->h_pin.as<float>();
    for (size_t i = 0; i < n; i++) f[i] = (float)h[i];
    CK(b.ensure(std::max<size_t>(n, 1) * 4));
    CK(cudaMemcpyAsync(b.p, f, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

IcoDev ico_dev(const msm_ctx* ctx) {
    IcoDev d;
    d.base = ctx->d_ico_base.as<IcoBase>();
    d.depth = ctx->ico.depth;
    d.nleaf_per_base = 1 << (2 * ctx->ico.depth);
    d.leaf_off = (d.nleaf_per_base - 1) / 3;
    d.ntab = ctx->d_ntab.as<float4>();
    d.ntab_d = ctx->d_ntab_d.as<double4>();
    d.mtab = ctx->d_mtab.as<float4>() + kMtabPad;
    return d;
}

int ensure_label_rot(msm_ctx* ctx) {
    if (!ctx->rot_dirty) return 0;
    if (ctx->N <= 0 || ctx->L <= 0) FAIL("set_cpgrid and set_labels must be called first");
    const size_t nl = (size_t)ctx->N * ctx->L;
    CK(ctx->d_disp.ensure(nl * 3 * sizeof(double)));
    CK(ctx->d_Kf.ensure(nl * 9 * sizeof(float)));
    CK(ctx->d_Kd.ensure(nl * 9 * sizeof(double)));
    Timer t(ctx, "label_rot");
    k_label_rot<<<(unsigned)((nl + 255) / 256), 256, 0, ctx->stream>>>(ctx->N, ctx->L, ctx->d_cp.as<double>(), ctx->d_labels.as<double>(),
                                                                      ctx->centre[0], ctx->centre[1], ctx->centre[2],
                                                                      ctx->d_disp.as<double>(), ctx->d_Kf.as<float>(), ctx->d_Kd.as<double>());
    KLAUNCH_CHECK();
    ctx->rot_dirty = false;
    return 0;
}

int ensure_face_feat(msm_ctx* ctx) {
    if (!ctx->have_target || ctx->Ct <= 0) FAIL("target features not set");
    const int want = ctx->multivariate ? (ctx->Ct + 3) / 4 : 0;
    if (!ctx->multivariate && ctx->Ct != 1) FAIL("univariate cost function needs C == 1");
    if (ctx->face_feat_cq == want) return 0;
    const int Cp = want == 0 ? 1 : 4 * want;   // any channel count: > 16 channels run the generic kernel variant
    const size_t per_face = want == 0 ? 4 : (size_t)3 * Cp;
    CK(ctx->d_face_feat.ensure((size_t)ctx->Ft * per_face * sizeof(float)));
    k_face_pack<<<(ctx->Ft + 255) / 256, 256, 0, ctx->stream>>>(ctx->Ft, ctx->Ct, Cp, ctx->Vt, want == 0 ? 1 : 0, ctx->d_leaf_vid.as<int>(),
                                                               ctx->d_tfeat.as<double>(), ctx->d_face_feat.as<float>());
    KLAUNCH_CHECK();
    ctx->face_feat_cq = want;
    return 0;
}

// Hint level: the deepest level whose faces are still comfortably larger than the disc a control point's patch
// sweeps under all labels (patch radius + largest label displacement); few edge walks, few levels left to descend.
int choose_hint_level(const msm_ctx* ctx) {
    const int depth = ctx->ico.depth;
    if (depth <= 0) return 0;
    if (const char* e = getenv("MSM_HINT_LEVEL")) {
        int h = atoi(e);
        return std::max(0, std::min(depth - 1, h));
    }
    double maxsep = 0.0;
    for (double v : ctx->h_maxsep) maxsep = std::max(maxsep, v);
    const double disc = ctx->par.range * maxsep + ctx->max_label_disp;
    const double edge0 = 110.71487177940904;  // icosahedron edge arc length at R = 100
    int h = 0;
    while (h + 1 < depth && edge0 / double(1 << (h + 1)) > 2.0 * disc) h++;
    return h;
}

int ensure_hint_level(msm_ctx* ctx) {
    const int h = choose_hint_level(ctx);
    if (h == ctx->hint_level) return 0;
    std::vector<double> hdual, hT;
    std::vector<int> hnbr;
    ctx->ico.build_hint_level(h, hdual, hnbr, hT);
    std::vector<float> hTf(hnbr.size() * 12, 0.f);   // per (face, edge): 3 rows of float4
    for (size_t fe = 0; fe < hnbr.size(); fe++) {
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) hTf[fe * 12 + 4 * r + c] = (float)hT[fe * 9 + 3 * r + c];
        int nb = hnbr[fe];
        std::memcpy(&hTf[fe * 12 + 3], &nb, 4);
    }
    if (upload(ctx, ctx->d_hdual, hdual.data(), hdual.size())) return 1;
    if (upload(ctx, ctx->d_hnbr, hnbr.data(), hnbr.size())) return 1;
    if (upload(ctx, ctx->d_hT, hTf.data(), hTf.size())) return 1;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->hint_level = h;
    return 0;
}

HintDev hint_dev(const msm_ctx* ctx) {
    HintDev d;
    d.h = ctx->hint_level;
    d.nf_per_base = 1 << (2 * d.h);
    d.node_off = (d.nf_per_base - 1) / 3;
    d.hdual = ctx->d_hdual.as<double>();
    d.hnbr = ctx->d_hnbr.as<int>();
    d.hT = ctx->d_hT.as<float4>();
    return d;
}

TripArgs trip_args(msm_ctx* ctx) {
    TripArgs a;
    a.p = dev_params(ctx);
    a.L = ctx->L;
    a.group_N = ctx->N_subj; a.group_T = ctx->T_subj;
    a.t_off = 0;
    a.trip = ctx->d_trip.as<int>();
    a.cp = ctx->d_cp.as<double>();
    a.orig = ctx->d_orig.as<double>();
    a.disp = ctx->d_disp.as<double>();
    return a;
}

int check_triplets(msm_ctx* ctx) {
    if (ctx->Tn <= 0) FAIL("set_triplets must be called first");
    if (ctx->par.group) {
        if (ctx->N_subj <= 0 || ctx->T_subj <= 0) FAIL("group mode needs msm_set_group");
    } else if (ctx->par.rmode != 2 && ctx->par.rmode != 3) {
        FAIL("DiscreteModel computeTripletCost regoption does not exist");  // src/DiscreteCostFunction.cpp:242
    }
    return ensure_label_rot(ctx);
}

PairArgs pair_args(msm_ctx* ctx) {
    PairArgs a;
    a.p = dev_params(ctx);
    a.L = ctx->L; a.P = ctx->P;
    a.pairs = ctx->d_pairs.as<int>();
    a.Kd = ctx->d_Kd.as<double>();
    a.disp = ctx->d_disp.as<double>();
    a.cp = ctx->d_cp.as<double>();
    a.ocp = ctx->d_ocp.as<double>();
    a.tris = ctx->d_tris.as<int>();
    a.vt_off = ctx->d_vt_off.as<int>();
    a.vt_idx = ctx->d_vt_idx.as<int>();
    return a;
}

}  // namespace

extern "C" {

static int patch_settle(msm_ctx* ctx, bool redo, int* redone);

void msm_default_params(msm_params* p) {
    p->lambda = 1.0; p->rexp = 2.0; p->mu = 0.4f; p->kappa = 1.6f; p->k_exp = 2.0; p->range = 1.0;
    p->simmeasure = 2; p->rmode = 1; p->dweight = 0; p->anorm = 0; p->meanangle = 0.0; p->mvdmax = 0.0; p->group = 0;
    p->bary_orthogonal = 0;
}

int msm_create(int device, msm_ctx** out) {
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_create_error = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); this library has no CPU fallback";
        return 1;
    }
    if (device < 0 || device >= n) { g_create_error = "device index out of range"; return 1; }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return 1; }
    msm_ctx* ctx = new msm_ctx();
    ctx->device = device;
    msm_default_params(&ctx->par);
    if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) {
        g_create_error = cudaGetErrorString(e);
        delete ctx;
        return 1;
    }
    ctx->stream = ctx->own_stream;
    *out = ctx;
    return 0;
}

void msm_destroy(msm_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_patch) cudaEventDestroy(ctx->ev_patch);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* msm_last_error(const msm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int msm_set_stream(msm_ctx* ctx, void* s) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return 0;
}
int msm_synchronize(msm_ctx* ctx) { CK(cudaSetDevice(ctx->device)); CK(cudaStreamSynchronize(ctx->stream)); return 0; }

int msm_set_params(msm_ctx* ctx, const msm_params* p) {
    if (p->bary_orthogonal) FAIL("bary_orthogonal: only the radial projection (A4 default) is implemented on the device");
    if (p->range != ctx->par.range) {   // patches were built for the old range
        if (patch_settle(ctx, false, nullptr)) return 1;
        ctx->have_patches = false; ctx->have_unary = false;
    }
    if (p->simmeasure != ctx->par.simmeasure) ctx->have_unary = false;                        // sign of the cached table
    ctx->par = *p;
    return 0;
}

int msm_set_target(msm_ctx* ctx, const double* xyz, int V, const int* tris, int F, const double* feat, int C) {
    CK(cudaSetDevice(ctx->device));
    ctx->have_target = false;
    ctx->face_feat_cq = -1;
    ctx->hint_level = -1;
    ctx->vkey_valid = false;
    if (!ctx->ico.build(xyz, V, tris, F)) FAIL("msm_set_target: " + ctx->ico.error);
    const IcoHierarchy& H = ctx->ico;
    ctx->Vt = V; ctx->Ft = F; ctx->Ct = feat ? C : 0;
    // node tables
    {
        size_t nn = std::max<size_t>(H.ntab.size(), 1);
        std::vector<float> nf(nn * 4, 0.f);
        std::vector<double> nd(nn * 4, 0.0);
        for (size_t i = 0; i < H.ntab.size(); i++)
            for (int j = 0; j < 3; j++) { nf[4 * i + j] = (float)H.ntab[i][j]; nd[4 * i + j] = H.ntab[i][j]; }
        if (upload(ctx, ctx->d_ntab, nf.data(), nf.size())) return 1;
        if (upload(ctx, ctx->d_ntab_d, nd.data(), nd.size())) return 1;
        // child-indexed multiplier table (all nodes of levels 0..depth; node 0 unused)
        const size_t nall = ((size_t(1) << (2 * (H.depth + 1))) - 1) / 3;
        // stored with kMtabPad leading entries so that the four children of a node (heap ids 4n+1..4n+4) sit in ONE aligned
        // 64-byte block: a warp's gather then never straddles two 128-byte lines for siblings
        std::vector<float> mt((nall + kMtabPad) * 4, 1.f);
        for (size_t n = 0; n < H.ntab.size(); n++) {
            const float nab = (float)H.ntab[n][0], nbc = (float)H.ntab[n][1], nca = (float)H.ntab[n][2];
            const float mul[4][3] = {{1.f, nab, nca}, {nab, 1.f, nbc}, {nca, nbc, 1.f}, {nbc, nca, nab}};
            for (int c = 0; c < 4; c++)
                for (int j = 0; j < 3; j++) mt[(kMtabPad + 4 * n + 1 + c) * 4 + j] = mul[c][j];
        }
        if (upload(ctx, ctx->d_mtab, mt.data(), mt.size())) return 1;
        CK(cudaStreamSynchronize(ctx->stream));
    }
    {
        IcoBase hb;  // ~3.4 KB, per context (a process-wide __constant__ would be shared by every context of the device)
        for (int b = 0; b < 20; b++) {
            for (int j = 0; j < 9; j++) { hb.dual[b][j] = (float)H.dual[b][j]; hb.dual_d[b][j] = H.dual[b][j]; }
            for (int j = 0; j < 3; j++) { hb.centre[b][j] = (float)H.centre[b][j]; hb.centre_d[b][j] = H.centre[b][j]; }
        }
        if (upload(ctx, ctx->d_ico_base, &hb, 1)) return 1;
        CK(cudaStreamSynchronize(ctx->stream));
    }
    {
        std::vector<int> lv((size_t)F * 3);
        std::vector<unsigned char> perm((size_t)F * 3);
        for (int f = 0; f < F; f++) {
            const int uf = H.leaf_user_face[f];
            for (int j = 0; j < 3; j++) {
                lv[3 * f + j] = H.leaf_vid[f][j];
                int pos = 0;
                for (int q = 0; q < 3; q++) if (tris[3 * uf + q] == H.leaf_vid[f][j]) pos = q;
                perm[3 * f + j] = (unsigned char)pos;
            }
        }
        if (upload(ctx, ctx->d_leaf_vid, lv.data(), lv.size())) return 1;
        if (upload(ctx, ctx->d_leaf_user, H.leaf_user_face.data(), H.leaf_user_face.size())) return 1;
        if (upload(ctx, ctx->d_leaf_perm, perm.data(), perm.size())) return 1;
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (feat) {
        if (C <= 0) FAIL("msm_set_target: C must be positive");
        if (upload(ctx, ctx->d_tfeat, feat, (size_t)C * V)) return 1;
        CK(cudaStreamSynchronize(ctx->stream));
    }
    ctx->have_target = true;
    return 0;
}

int msm_set_source(msm_ctx* ctx, const double* xyz, int V, const double* feat, int C, const double* cfweight, int wrows, int multivariate) {
    CK(cudaSetDevice(ctx->device));
    if (V <= 0 || C <= 0) FAIL("CostFunction::You must supply source and target meshes.");  // src/DiscreteCostFunction.cpp:112-113
    if (wrows != 0 && wrows != 1 && wrows != C)
        FAIL("DiscreteModel ERROR:: costfunction weighting has dimensions incompatible with data");  // :116-117
    ctx->Vs = V; ctx->Cs = C; ctx->wrows = wrows; ctx->multivariate = multivariate != 0;
    if (upload(ctx, ctx->d_sfeat, feat, (size_t)C * V)) return 1;
    CK(cudaStreamSynchronize(ctx->stream));
    if (wrows > 0 && upload_as_float(ctx, ctx->d_swgt, cfweight, (size_t)wrows * V)) return 1;
    if (patch_settle(ctx, false, nullptr)) return 1;
    ctx->have_patches = false;
    return msm_reset_source(ctx, xyz);
}

int msm_reset_source(msm_ctx* ctx, const double* xyz) {
    CK(cudaSetDevice(ctx->device));
    if (ctx->Vs <= 0) FAIL("msm_set_source must be called first");
    // one host pass into a persistent pinned buffer: it is the staging area of the (then truly asynchronous) upload, the
    // fp64 copy the exact-tie resolution of msm_build_patches reads, and yields the bounding radius of the voxel grid;
    // the fp32 SoA copy used by the patch pre-filter is derived on the device
    const size_t n3 = (size_t)ctx->Vs * 3;
    CK(cudaStreamSynchronize(ctx->stream));   // a previous upload may still be reading the staging buffer
    if (patch_settle(ctx, false, nullptr)) return 1;   // a deferred patch build of the old coordinates is retired, not repeated
    CK(ctx->h_src.ensure(n3 * sizeof(double)));
    double* hs = ctx->h_src.as<double>();
    double m0 = 0.0, m1 = 0.0, m2 = 0.0, m3 = 0.0;
    size_t i = 0;
    for (; i + 4 <= n3; i += 4) {   // four independent max chains
        const double a = xyz[i], b = xyz[i + 1], c = xyz[i + 2], d = xyz[i + 3];
        hs[i] = a; hs[i + 1] = b; hs[i + 2] = c; hs[i + 3] = d;
        m0 = std::max(m0, std::fabs(a)); m1 = std::max(m1, std::fabs(b)); m2 = std::max(m2, std::fabs(c)); m3 = std::max(m3, std::fabs(d));
    }
    for (; i < n3; i++) { hs[i] = xyz[i]; m0 = std::max(m0, std::fabs(xyz[i])); }
    ctx->src_bound = std::max(std::max(m0, m1), std::max(m2, m3));
    if (upload(ctx, ctx->d_src, hs, n3)) return 1;
    CK(ctx->d_sxf.ensure((size_t)ctx->Vs * 4)); CK(ctx->d_syf.ensure((size_t)ctx->Vs * 4)); CK(ctx->d_szf.ensure((size_t)ctx->Vs * 4));
    k_split_xyz<<<(ctx->Vs + 255) / 256, 256, 0, ctx->stream>>>(ctx->Vs, ctx->d_src.as<double>(), ctx->d_sxf.as<float>(), ctx->d_syf.as<float>(),
                                                               ctx->d_szf.as<float>());
    KLAUNCH_CHECK();
    ctx->have_patches = false;
    ctx->vkey_valid = false;
    return 0;
}

int msm_set_cpgrid(msm_ctx* ctx, const double* xyz, int N, const int* tris, int T, const double* maxsep, const double* orig_xyz,
                   const double* abs_weights) {
    CK(cudaSetDevice(ctx->device));
    if (N <= 0) FAIL("msm_set_cpgrid: empty control grid");
    if (patch_settle(ctx, false, nullptr)) return 1;
    ctx->N = N; ctx->Tcp = T;
    ctx->k_begin = 0; ctx->k_end = N;
    ctx->h_cp.assign(xyz, xyz + (size_t)N * 3);
    if (upload(ctx, ctx->d_cp, xyz, (size_t)N * 3)) return 1;
    if (upload(ctx, ctx->d_ocp, xyz, (size_t)N * 3)) return 1;
    if (upload(ctx, ctx->d_orig, orig_xyz ? orig_xyz : xyz, (size_t)N * 3)) return 1;
    if (maxsep) {
        ctx->h_maxsep.assign(maxsep, maxsep + N);
        if (upload(ctx, ctx->d_maxsep, maxsep, (size_t)N)) return 1;
    } else ctx->h_maxsep.clear();
    {
        std::vector<double> aw(N, 1.0);
        if (abs_weights) aw.assign(abs_weights, abs_weights + N);
        if (upload_as_float(ctx, ctx->d_absw, aw.data(), (size_t)N)) return 1;
    }
    if (T > 0) {
        for (int i = 0; i < 3 * T; i++)
            if (tris[i] < 0 || tris[i] >= N) FAIL("msm_set_cpgrid: triangle index out of range");
        if (upload(ctx, ctx->d_tris, tris, (size_t)T * 3)) return 1;
        std::vector<int> off(N + 1, 0), idx((size_t)T * 3);
        for (int i = 0; i < 3 * T; i++) off[tris[i] + 1]++;
        for (int i = 0; i < N; i++) off[i + 1] += off[i];
        std::vector<int> fill(off.begin(), off.end() - 1);
        for (int t = 0; t < T; t++)
            for (int j = 0; j < 3; j++) idx[fill[tris[3 * t + j]]++] = t;
        if (upload(ctx, ctx->d_vt_off, off.data(), off.size())) return 1;
        if (upload(ctx, ctx->d_vt_idx, idx.data(), idx.size())) return 1;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->rot_dirty = true;
    ctx->have_patches = false;
    ctx->have_unary = false;
    return 0;
}

int msm_set_abs_weights(msm_ctx* ctx, const double* abs_weights) {
    CK(cudaSetDevice(ctx->device));
    if (ctx->N <= 0) FAIL("msm_set_cpgrid must be called first");
    if (upload_as_float(ctx, ctx->d_absw, abs_weights, (size_t)ctx->N)) return 1;
    ctx->have_unary = false;
    return 0;
}

int msm_reset_cpgrid(msm_ctx* ctx, const double* xyz) {
    CK(cudaSetDevice(ctx->device));
    if (ctx->N <= 0) FAIL("msm_set_cpgrid must be called first");
    if (patch_settle(ctx, false, nullptr)) return 1;   // (the exact redo of a deferred build reads h_cp)
    ctx->h_cp.assign(xyz, xyz + (size_t)ctx->N * 3);
    if (upload(ctx, ctx->d_cp, xyz, (size_t)ctx->N * 3)) return 1;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->rot_dirty = true;
    ctx->have_patches = false;
    ctx->have_unary = false;
    return 0;
}

int msm_set_shard(msm_ctx* ctx, int k_begin, int k_end) {
    if (k_begin < 0 || k_end > ctx->N || k_begin > k_end) FAIL("msm_set_shard: bad range");
    if (patch_settle(ctx, false, nullptr)) return 1;
    ctx->k_begin = k_begin; ctx->k_end = k_end;
    ctx->have_patches = false;
    ctx->have_unary = false;
    return 0;
}

int msm_set_labels(msm_ctx* ctx, const double* labels, int L, const double* centre) {
    CK(cudaSetDevice(ctx->device));
    if (L <= 0) FAIL("msm_set_labels: empty label set");
    ctx->L = L;
    for (int j = 0; j < 3; j++) ctx->centre[j] = centre[j];
    ctx->max_label_disp = 0.0;
    for (int l = 0; l < L; l++) {
        const double dx = labels[3 * l] - centre[0], dy = labels[3 * l + 1] - centre[1], dz = labels[3 * l + 2] - centre[2];
        ctx->max_label_disp = std::max(ctx->max_label_disp, std::sqrt(dx * dx + dy * dy + dz * dz));
    }
    if (upload(ctx, ctx->d_labels, labels, (size_t)L * 3)) return 1;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->rot_dirty = true;
    ctx->have_unary = false;
    return 0;
}

int msm_set_pairs(msm_ctx* ctx, const int* pairs, int P) {
    CK(cudaSetDevice(ctx->device));
    for (int i = 0; i < 2 * P; i++)
        if (pairs[i] < 0 || pairs[i] >= ctx->N) FAIL("msm_set_pairs: node id out of range");
    ctx->P = P;
    if (upload(ctx, ctx->d_pairs, pairs, (size_t)P * 2)) return 1;
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int msm_set_triplets(msm_ctx* ctx, const int* triplets, int T) {
    CK(cudaSetDevice(ctx->device));
    for (int i = 0; i < 3 * T; i++)
        if (triplets[i] < 0 || triplets[i] >= ctx->N) FAIL("msm_set_triplets: node id out of range");
    ctx->Tn = T;
    if (upload(ctx, ctx->d_trip, triplets, (size_t)T * 3)) return 1;
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int msm_set_group(msm_ctx* ctx, int S, int N_subj, int T_subj) {
    if (S <= 0 || (long long)S * N_subj != ctx->N) FAIL("msm_set_group: S*N_subj must equal the control grid size");
    ctx->S = S; ctx->N_subj = N_subj; ctx->T_subj = T_subj;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
static PatchArgs patch_args(msm_ctx* ctx) {
    PatchArgs a;
    a.k_begin = ctx->k_begin; a.k_end = ctx->k_end; a.V = ctx->Vs;
    a.sxf = ctx->d_sxf.as<float>(); a.syf = ctx->d_syf.as<float>(); a.szf = ctx->d_szf.as<float>();
    a.src = ctx->d_src.as<double>(); a.cp = ctx->d_cp.as<double>(); a.maxsep = ctx->d_maxsep.as<double>();
    a.range = ctx->par.range;
    a.counts = ctx->d_counts.as<int>();
    a.unc = ctx->d_unc.as<int2>(); a.unc_cap = kUncCap; a.unc_count = ctx->d_unc_count.as<int>();
    a.acc_off = nullptr; a.acc_idx = nullptr; a.off = nullptr; a.idx = nullptr;
    a.slab = nullptr; a.slab_cap = 0;
    return a;
}

static int patches_exact(msm_ctx* ctx, const BinGrid& bg, bool binned);

// Reads the flag word of a deferred one-pass build.  redo = true: a build that saw boundary ties or an overflowing row is
// repeated through the exact path (the voxel grid of that build is still in place); redo = false: the inputs are about
// to change, the build is only retired.
static int patch_settle(msm_ctx* ctx, bool redo, int* redone) {
    if (redone) *redone = 0;
    if (!ctx->patch_pending) return 0;
    CK(cudaEventSynchronize(ctx->ev_patch));
    ctx->patch_pending = false;
    const int* hf = ctx->h_pflags.as<int>();
    if (hf[0] == 0 && hf[1] == 0) {
        ctx->total = hf[3]; ctx->maxP = hf[2];
        const long long want = (long long)ctx->maxP + ctx->maxP / 4 + 16;
        ctx->slab_cap = (int)std::max<long long>(want, 32);
        ctx->fast_builds++;
        return 0;
    }
    ctx->have_patches = false;
    ctx->have_unary = false;
    if (!redo) return 0;
    if (redone) *redone = 1;
    Timer tm(ctx, "patches");
    CK(cudaMemsetAsync(ctx->d_unc_count.p, 0, 4, ctx->stream));
    return patches_exact(ctx, ctx->pend_bg, ctx->pend_binned);
}

int msm_build_patches_async(msm_ctx* ctx) {
    CK(cudaSetDevice(ctx->device));
    if (ctx->Vs <= 0 || ctx->N <= 0) FAIL("CostFunction::You must supply source and target meshes.");
    if (ctx->h_maxsep.empty()) FAIL("msm_build_patches: set_spacings (maxsep) missing");
    if (patch_settle(ctx, false, nullptr)) return 1;   // retire the previous build (its largest patch sizes this one's rows)
    const int Ns = ctx->k_end - ctx->k_begin;
    ctx->have_patches = false;
    ctx->total = 0; ctx->maxP = 0;
    if (Ns == 0) { ctx->have_patches = true; return 0; }
    CK(ctx->d_counts.ensure((size_t)Ns * 4));
    CK(ctx->d_off.ensure((size_t)(Ns + 1) * 4));
    CK(ctx->d_unc.ensure((size_t)kUncCap * sizeof(int2)));
    CK(ctx->d_unc_count.ensure(4));
    CK(ctx->d_totmax.ensure(8));
    CK(cudaMemsetAsync(ctx->d_unc_count.p, 0, 4, ctx->stream));
    const PatchArgs a = patch_args(ctx);
    Timer tm(ctx, "patches");
    // voxel grid over the source vertices (cell >= largest patch chord); brute force only for degenerate sizes
    BinGrid bg{};
    bool binned = false;
    {
        double maxsep = 0.0;
        for (int k = ctx->k_begin; k < ctx->k_end; k++) maxsep = std::max(maxsep, ctx->h_maxsep[k]);
        const double Rb = ctx->src_bound * (1.0 + 1e-9) + 1e-6;
        const double half = ctx->par.range * maxsep / 200.0;
        const char* brute = getenv("MSM_PATCH_BRUTE");
        if (half < 1.5 && !(brute && atoi(brute))) {
            const double chord = 200.0 * std::sin(half) * (1.0 + 1e-6);
            const int G = (int)std::min(64.0, std::floor(2.0 * Rb / std::max(chord, 1e-9)));
            if (G >= 3) {
                const size_t nc = (size_t)G * G * G;
                CK(ctx->d_cell_counts.ensure(nc * 4)); CK(ctx->d_cell_start.ensure((nc + 1) * 4)); CK(ctx->d_cell_fill.ensure(nc * 4));
                CK(ctx->d_cell_of.ensure((size_t)ctx->Vs * 4)); CK(ctx->d_sorted.ensure((size_t)ctx->Vs * 4));
                CK(cudaMemsetAsync(ctx->d_cell_counts.p, 0, nc * 4, ctx->stream));
                CK(cudaMemsetAsync(ctx->d_cell_fill.p, 0, nc * 4, ctx->stream));
                bg.G = G; bg.lo = -Rb; bg.inv_h = G / (2.0 * Rb);
                bg.cell_start = ctx->d_cell_start.as<int>(); bg.cell_fill = ctx->d_cell_fill.as<int>(); bg.sorted = ctx->d_sorted.as<int>();
                k_bin_count<<<(ctx->Vs + 255) / 256, 256, 0, ctx->stream>>>(bg, ctx->Vs, ctx->d_src.as<double>(), ctx->d_cell_of.as<int>(),
                                                                           ctx->d_cell_counts.as<int>());
                KLAUNCH_CHECK();
                {   // exclusive scan of the cell counts: chunk sums, then per-chunk scans (one CTA per chunk)
                    const int nchunks = 148, chunk = ((int)nc + nchunks - 1) / nchunks;
                    CK(ctx->d_scan_sums.ensure((size_t)nchunks * 4));
                    k_scan_chunk_sums<<<nchunks, 256, 0, ctx->stream>>>((int)nc, chunk, ctx->d_cell_counts.as<int>(), ctx->d_scan_sums.as<int>());
                    KLAUNCH_CHECK();
                    k_scan_chunk_apply<<<nchunks, 256, 0, ctx->stream>>>((int)nc, chunk, ctx->d_cell_counts.as<int>(), ctx->d_scan_sums.as<int>(),
                                                                        ctx->d_cell_start.as<int>());
                    KLAUNCH_CHECK();
                }
                k_bin_fill<<<(ctx->Vs + 255) / 256, 256, 0, ctx->stream>>>(bg, ctx->Vs, ctx->d_cell_of.as<int>());
                KLAUNCH_CHECK();
                binned = true;
            }
        }
    }
    // One-pass build (patchfast.cuh): from the second build of a level on, when the previous build's largest patch bounds
    // this one's.  Falls through to the exact two-pass path below on boundary ties or an overflowing row.
    {
        const char* slow = getenv("MSM_PATCH_TWOPASS");
        const char* nosort = getenv("MSM_PATCH_NOSORT");
        const int cap = ctx->slab_cap;
        const int Cp = ctx->multivariate ? 4 * ((ctx->Cs + 3) / 4) : 1;
        const size_t rows = (size_t)Ns * (size_t)std::max(cap, 1);
        if (binned && cap > 0 && cap <= 4096 && ctx->have_target && !(slow && atoi(slow)) && !(nosort && atoi(nosort)) &&
            (ctx->multivariate || ctx->Cs == 1) && rows * 4 * (5 + 2 * (size_t)Cp) <= ((size_t)2 << 30)) {
            if (!ctx->vkey_valid) {
                CK(ctx->d_vkey.ensure((size_t)ctx->Vs * 4));
                k_vertex_key<<<(ctx->Vs + 127) / 128, 128, 0, ctx->stream>>>(ctx->Vs, ctx->d_src.as<double>(), ico_dev(ctx), ctx->d_vkey.as<int>());
                KLAUNCH_CHECK();
                ctx->vkey_valid = true;
            }
            CK(ctx->d_idx.ensure(rows * 4));
            CK(ctx->d_px.ensure(rows * 4)); CK(ctx->d_py.ensure(rows * 4)); CK(ctx->d_pz.ensure(rows * 4));
            CK(ctx->d_pa.ensure(rows * 4 * Cp)); CK(ctx->d_pw.ensure(rows * 4 * Cp));
            CK(ctx->d_pflags.ensure(16));
            CK(cudaMemsetAsync(ctx->d_pflags.p, 0, 16, ctx->stream));
            PatchFastArgs f;
            f.k_begin = ctx->k_begin; f.Ns = Ns; f.V = ctx->Vs; f.cap = cap; f.g = bg;
            f.sxf = a.sxf; f.syf = a.syf; f.szf = a.szf; f.src = a.src; f.cp = a.cp; f.maxsep = a.maxsep; f.range = a.range;
            f.vkey = ctx->d_vkey.as<int>();
            f.Cp = Cp; f.C = ctx->Cs; f.wrows = ctx->wrows; f.univariate = ctx->multivariate ? 0 : 1;
            f.feat = ctx->d_sfeat.as<double>(); f.wgt = ctx->d_swgt.as<float>();
            CK(ctx->d_end.ensure((size_t)Ns * 4));
            f.counts = ctx->d_counts.as<int>(); f.idx = ctx->d_idx.as<int>();
            f.off = ctx->d_off.as<int>(); f.end = ctx->d_end.as<int>();
            f.px = ctx->d_px.as<float>(); f.py = ctx->d_py.as<float>(); f.pz = ctx->d_pz.as<float>();
            f.pa = ctx->d_pa.as<float>(); f.pw = ctx->d_pw.as<float>();
            f.flags = ctx->d_pflags.as<int>();
            const size_t smem = (size_t)kPatchWarps * 2 * cap * sizeof(int);
            {
                BigSmemLock big(smem);
                const unsigned grid = (unsigned)((Ns + kPatchWarps - 1) / kPatchWarps);
                const char* er = getenv("MSM_PATCH_R");
                const int R = er ? atoi(er) : std::min(4, (cap + 31) / 32);
#define PATCH_FUSED(RR)                                                                                                        \
    do {                                                                                                                       \
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_patch_fused<RR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_patch_fused<RR><<<grid, kPatchWarps * 32, smem, ctx->stream>>>(f);                                                   \
    } while (0)
                if (R == 1) PATCH_FUSED(1); else if (R == 2) PATCH_FUSED(2); else if (R == 3) PATCH_FUSED(3); else PATCH_FUSED(4);
#undef PATCH_FUSED
            }
            KLAUNCH_CHECK();
            // the flag word follows the kernel to the host; nobody waits for it here (msm_patches_resolve / the host-output
            // entry points do).  Until then the row capacity bounds the sizes the launches below need.
            CK(ctx->h_pflags.ensure(16));
            if (!ctx->ev_patch) CK(cudaEventCreateWithFlags(&ctx->ev_patch, cudaEventDisableTiming));
            CK(cudaMemcpyAsync(ctx->h_pflags.p, ctx->d_pflags.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaEventRecord(ctx->ev_patch, ctx->stream));
            ctx->patch_pending = true;
            ctx->pend_bg = bg; ctx->pend_binned = binned;
            ctx->total = (long long)Ns * cap; ctx->maxP = cap;
            ctx->pack_cp = Cp;
            ctx->patches_padded = true; ctx->pad_cap = cap;
            ctx->have_patches = true;
            ctx->have_unary = false;
            return 0;
        }
    }
    return patches_exact(ctx, bg, binned);
}

int msm_patches_resolve(msm_ctx* ctx, int* redone) {
    CK(cudaSetDevice(ctx->device));
    return patch_settle(ctx, true, redone);
}

int msm_build_patches(msm_ctx* ctx) {
    if (msm_build_patches_async(ctx)) return 1;
    return patch_settle(ctx, true, nullptr);
}

// The exact two-pass path: first build of a level, boundary ties, rows that outgrew their capacity.
static int patches_exact(msm_ctx* ctx, const BinGrid& bg, bool binned) {
    const int Ns = ctx->k_end - ctx->k_begin;
    const int unc_cap = kUncCap;
    PatchArgs a = patch_args(ctx);
    ctx->patches_padded = false;
    ctx->exact_builds++;
    // Patches change little between outer iterations: from the second build on, the count pass stashes the members it finds
    // in a slab sized from the previous build's largest patch, and the fill pass only sorts them (no second search).  A
    // patch that outgrew the slab sends the whole build through the two-pass path again.
    const char* twopass = getenv("MSM_PATCH_TWOPASS");
    const bool use_slab = binned && ctx->slab_cap > 0 && !(twopass && atoi(twopass));
    if (use_slab) {
        CK(ctx->d_slab.ensure((size_t)Ns * ctx->slab_cap * 4));
        a.slab = ctx->d_slab.as<int>(); a.slab_cap = ctx->slab_cap;
    }
    if (binned) k_patch_binned<false><<<Ns, 128, 0, ctx->stream>>>(a, bg, 0);
    else k_patch<false><<<Ns, 256, 0, ctx->stream>>>(a);
    KLAUNCH_CHECK();
    // common case (no boundary ties): the offsets scan runs speculatively so that one host sync returns both the tie
    // count and the totals
    k_patch_scan<<<1, 1024, 0, ctx->stream>>>(Ns, ctx->d_counts.as<int>(), nullptr, ctx->d_off.as<int>(), ctx->d_totmax.as<int>());
    KLAUNCH_CHECK();
    CK(ctx->h_pin.ensure(16));
    int* hp = ctx->h_pin.as<int>();
    CK(cudaMemcpyAsync(hp, ctx->d_unc_count.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hp + 1, ctx->d_totmax.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int n_unc = hp[0];
    int totmax[2] = {hp[1], hp[2]};
    if (n_unc > unc_cap) FAIL("msm_build_patches: too many boundary ties");
    bool have_acc = false;
    if (n_unc > 0) {
        // narrow phase on the host with the reference's own formula (src/DiscreteCostFunction.cpp:104-109)
        std::vector<int2> unc(n_unc);
        CK(cudaMemcpy(unc.data(), ctx->d_unc.p, (size_t)n_unc * sizeof(int2), cudaMemcpyDeviceToHost));
        std::vector<std::vector<int>> acc(Ns);
        const double RAD = 100.0;
        for (const int2& u : unc) {
            const double* c = &ctx->h_cp[3 * (size_t)u.x];
            const double* s = ctx->h_src.as<double>() + 3 * (size_t)u.y;
            const double X = c[0] - s[0], Y = c[1] - s[1], Z = c[2] - s[2];
            const double norm = std::sqrt(X * X + Y * Y + Z * Z);
            if ((2 * RAD * std::asin(norm / (2 * RAD))) < ctx->par.range * ctx->h_maxsep[u.x]) acc[u.x - ctx->k_begin].push_back(u.y);
        }
        std::vector<int> aoff(Ns + 1, 0), aidx;
        for (int r = 0; r < Ns; r++) {
            std::sort(acc[r].begin(), acc[r].end());
            aoff[r + 1] = aoff[r] + (int)acc[r].size();
            aidx.insert(aidx.end(), acc[r].begin(), acc[r].end());
        }
        if (upload(ctx, ctx->d_acc_off, aoff.data(), aoff.size())) return 1;
        if (upload(ctx, ctx->d_acc_idx, aidx.data(), aidx.size())) return 1;
        CK(cudaStreamSynchronize(ctx->stream));
        have_acc = true;
        k_patch_scan<<<1, 1024, 0, ctx->stream>>>(Ns, ctx->d_counts.as<int>(), ctx->d_acc_off.as<int>(), ctx->d_off.as<int>(), ctx->d_totmax.as<int>());
        KLAUNCH_CHECK();
        CK(cudaMemcpyAsync(hp + 1, ctx->d_totmax.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        totmax[0] = hp[1]; totmax[1] = hp[2];
    }
    ctx->total = totmax[0]; ctx->maxP = totmax[1];
    const size_t tot = std::max<long long>(ctx->total, 1);
    CK(ctx->d_idx.ensure(tot * 4));
    a.acc_off = have_acc ? ctx->d_acc_off.as<int>() : nullptr;
    a.acc_idx = have_acc ? ctx->d_acc_idx.as<int>() : nullptr;
    a.off = ctx->d_off.as<int>(); a.idx = ctx->d_idx.as<int>();
    int cap = 1;
    while (cap < ctx->maxP) cap <<= 1;
    if (binned && cap <= 16384) {
        BigSmemLock big((size_t)cap * 4);
        if (use_slab && ctx->maxP <= ctx->slab_cap) {
            if (cap * 4 > 48 * 1024) CK(cudaFuncSetAttribute(k_patch_from_slab, cudaFuncAttributeMaxDynamicSharedMemorySize, cap * 4));
            k_patch_from_slab<<<Ns, 128, (size_t)cap * 4, ctx->stream>>>(a, cap);
        } else {
            if (cap * 4 > 48 * 1024) CK(cudaFuncSetAttribute(k_patch_binned<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap * 4));
            k_patch_binned<true><<<Ns, 128, (size_t)cap * 4, ctx->stream>>>(a, bg, cap);
        }
        const long long want = (long long)ctx->maxP + ctx->maxP / 4 + 16;
        ctx->slab_cap = ((size_t)Ns * (size_t)want * 4 <= ((size_t)512 << 20)) ? (int)want : 0;
    } else {
        k_patch<true><<<Ns, 256, 0, ctx->stream>>>(a);
        ctx->slab_cap = 0;
    }
    KLAUNCH_CHECK();
    // packed sample streams
    const int Cp = ctx->multivariate ? 4 * ((ctx->Cs + 3) / 4) : 1;
    if (!ctx->multivariate && ctx->Cs != 1) FAIL("univariate cost function needs C == 1 (src/DiscreteModel.cpp:22-36)");
    ctx->pack_cp = Cp;
    CK(ctx->d_px.ensure(tot * 4)); CK(ctx->d_py.ensure(tot * 4)); CK(ctx->d_pz.ensure(tot * 4));
    CK(ctx->d_pa.ensure(tot * 4 * Cp)); CK(ctx->d_pw.ensure(tot * 4 * Cp));
    if (ctx->total > 0) {
        // spatial packing order (see k_patch_pack): needs the target hierarchy; skipped for very large patches
        const int* vkey = nullptr;
        const char* nosort = getenv("MSM_PATCH_NOSORT");
        if (ctx->have_target && ctx->maxP <= 8192 && !(nosort && atoi(nosort))) {
            if (!ctx->vkey_valid) {
                CK(ctx->d_vkey.ensure((size_t)ctx->Vs * 4));
                k_vertex_key<<<(ctx->Vs + 127) / 128, 128, 0, ctx->stream>>>(ctx->Vs, ctx->d_src.as<double>(), ico_dev(ctx), ctx->d_vkey.as<int>());
                KLAUNCH_CHECK();
                ctx->vkey_valid = true;
            }
            vkey = ctx->d_vkey.as<int>();
        }
        k_patch_pack<<<Ns, 128, vkey ? (size_t)ctx->maxP * 4 : 0, ctx->stream>>>(ctx->k_begin, (int)ctx->total, Cp, ctx->Cs, ctx->wrows, ctx->Vs, ctx->multivariate ? 0 : 1,
                                                 ctx->d_off.as<int>(), ctx->d_idx.as<int>(), vkey, ctx->d_src.as<double>(), ctx->d_cp.as<double>(),
                                                 ctx->d_sfeat.as<double>(), ctx->d_swgt.as<float>(), ctx->d_px.as<float>(),
                                                 ctx->d_py.as<float>(), ctx->d_pz.as<float>(), ctx->d_pa.as<float>(), ctx->d_pw.as<float>());
        KLAUNCH_CHECK();
    }
    ctx->have_patches = true;
    ctx->have_unary = false;
    return 0;
}

int msm_patch_count(msm_ctx* ctx, long long* total) {
    if (patch_settle(ctx, true, nullptr)) return 1;
    if (!ctx->have_patches) FAIL("msm_build_patches must be called first");
    *total = ctx->total;
    return 0;
}

int msm_patch_build_stats(const msm_ctx* ctx, long long* one_pass, long long* two_pass) {
    *one_pass = ctx->fast_builds; *two_pass = ctx->exact_builds;
    return 0;
}

int msm_get_patches(msm_ctx* ctx, int* offsets, int* idx) {
    CK(cudaSetDevice(ctx->device));
    if (patch_settle(ctx, true, nullptr)) return 1;
    if (!ctx->have_patches) FAIL("msm_build_patches must be called first");
    const int Ns = ctx->k_end - ctx->k_begin;
    std::vector<int> off(Ns + 1, 0);
    if (ctx->patches_padded) {
        // one-pass build: fixed-capacity rows in packed order; the public lists are ascending (the reference's order)
        std::vector<int> cnt(Ns), slab((size_t)Ns * ctx->pad_cap);
        if (Ns > 0) {
            CK(cudaMemcpyAsync(cnt.data(), ctx->d_counts.p, (size_t)Ns * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(slab.data(), ctx->d_idx.p, slab.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
        }
        CK(cudaStreamSynchronize(ctx->stream));
        for (int r = 0; r < Ns; r++) {
            off[r + 1] = off[r] + cnt[r];
            int* dst = idx + off[r];
            std::copy(slab.begin() + (size_t)r * ctx->pad_cap, slab.begin() + (size_t)r * ctx->pad_cap + cnt[r], dst);
            std::sort(dst, dst + cnt[r]);
        }
    } else {
        if (Ns > 0) CK(cudaMemcpyAsync(off.data(), ctx->d_off.p, (size_t)(Ns + 1) * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (ctx->total > 0) CK(cudaMemcpyAsync(idx, ctx->d_idx.p, (size_t)ctx->total * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    for (int k = 0; k <= ctx->N; k++) {
        if (k <= ctx->k_begin) offsets[k] = 0;
        else if (k <= ctx->k_end) offsets[k] = off[k - ctx->k_begin];
        else offsets[k] = (int)ctx->total;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
static int make_peer_set(msm_ctx* ctx, void* const* peer_tables, int world, void* multicast_table, PeerSet& ps) {
    if (world < 1 || world > kMaxPeers) FAIL("fused exchange: world size must be 1..16");
    ps.world = world;
    ps.mc = static_cast<float*>(multicast_table);
    for (int r = 0; r < kMaxPeers; r++) ps.p[r] = r < world ? static_cast<float*>(peer_tables[r]) : nullptr;
    for (int r = 0; r < world; r++)
        if (!ps.p[r]) FAIL("fused exchange: null peer table");
    return 0;
}

static int unary_launch(msm_ctx* ctx, float* d_out, const PeerSet& peers);

int msm_unary_table_dev(msm_ctx* ctx, float* d_out) {
    PeerSet none{};
    return unary_launch(ctx, d_out, none);
}

int msm_unary_table_peers(msm_ctx* ctx, void* const* peer_tables, int world, void* multicast_table) {
    PeerSet ps{};
    if (make_peer_set(ctx, peer_tables, world, multicast_table, ps)) return 1;
    return unary_launch(ctx, nullptr, ps);
}

static int unary_launch(msm_ctx* ctx, float* d_out, const PeerSet& peers) {
    CK(cudaSetDevice(ctx->device));
    if (!ctx->have_target) FAIL("CostFunction::You must supply source and target meshes.");
    if (!ctx->have_patches) FAIL("get_source_data (msm_build_patches) must run before computeUnaryCosts");
    if (ctx->Ct != ctx->Cs) FAIL("source and target feature dimensions differ");
    if (ensure_label_rot(ctx)) return 1;
    if (ensure_face_feat(ctx)) return 1;
    if (ensure_hint_level(ctx)) return 1;
    const int Ns = ctx->k_end - ctx->k_begin;
    if (Ns == 0) return 0;
    const int cq_actual = ctx->face_feat_cq;                 // channel quads of the face-packed layout
    const int CQ = cq_actual <= 4 ? cq_actual : 5;           // kernel variant (5 = generic, features read from global)
    const int Cp = CQ == 0 ? 1 : (CQ <= 4 ? 4 * CQ : 0);     // channels staged in shared memory
    const int L = ctx->L, Pmax = std::max(ctx->maxP, 1);
    // label groups: fit shared memory (<= ~56 KB so several CTAs share an SM) and expose >= 4 CTAs per SM
    auto smem_for = [&](int Lg) { return (size_t)4 * ((size_t)(6 + 2 * Cp) * Pmax + (size_t)Lg * 12 + (size_t)Lg * Pmax); };
    int Lg = L;
    while (Lg > 1 && smem_for(Lg) > 56 * 1024) Lg--;
    const int want_ctas = 4 * 148;
    if ((long long)Ns * ((L + Lg - 1) / Lg) < want_ctas) {
        int G = (want_ctas + Ns - 1) / Ns;
        Lg = std::min(Lg, std::max(1, (L + G - 1) / G));
    }
    if (const char* e = getenv("MSM_UNARY_LG")) {  // testing: force the label-group size
        Lg = std::max(1, std::min(L, atoi(e)));
        while (Lg > 1 && smem_for(Lg) > 200 * 1024) Lg--;
    }
    const size_t smem = smem_for(Lg);
    if (smem > 200 * 1024) FAIL("msm_unary_table: control-point patch too large for one CTA's shared memory");
    UnaryArgs a;
    a.k_begin = ctx->k_begin; a.Ns = Ns; a.L = L; a.Lg = Lg; a.G = (L + Lg - 1) / Lg;
    a.off = ctx->d_off.as<int>();
    a.end = ctx->patches_padded ? ctx->d_end.as<int>() : ctx->d_off.as<int>() + 1;
    a.px = ctx->d_px.as<float>(); a.py = ctx->d_py.as<float>(); a.pz = ctx->d_pz.as<float>();
    a.pa = ctx->d_pa.as<float>(); a.pw = ctx->d_pw.as<float>();
    a.total = ctx->patches_padded ? Ns * ctx->pad_cap : (int)ctx->total;
    a.cq_rt = cq_actual;
    a.hint = hint_dev(ctx);
    a.absw = ctx->d_absw.as<float>();
    a.face_feat = ctx->d_face_feat.as<float>();
    a.ico = ico_dev(ctx);
    a.simmeasure = ctx->par.simmeasure;
    a.out = d_out;
    a.peers = peers; a.n_full = ctx->N;
    const size_t tot = ctx->patches_padded ? (size_t)Ns * ctx->pad_cap : (size_t)std::max<long long>(ctx->total, 1);
    CK(ctx->d_pcx.ensure(tot * 4)); CK(ctx->d_pcy.ensure(tot * 4)); CK(ctx->d_pcz.ensure(tot * 4));
    CK(ctx->d_hint_face.ensure((size_t)Ns * 4));
    CK(ctx->d_lab12.ensure((size_t)Ns * L * 12 * 4));
    CK(ctx->d_tref.ensure((size_t)Ns * 4));
    a.pcx = ctx->d_pcx.as<float>(); a.pcy = ctx->d_pcy.as<float>(); a.pcz = ctx->d_pcz.as<float>();
    a.tref = ctx->d_tref.as<float>();
    a.hint_face = ctx->d_hint_face.as<int>();
    a.lab12 = ctx->d_lab12.as<float>();
    const unsigned grid = (unsigned)((long long)Ns * a.G);
    Timer tm(ctx, "unary");
    {
        UnarySetupArgs s;
        s.k_begin = ctx->k_begin; s.L = L; s.off = a.off; s.end = a.end; s.px = a.px; s.py = a.py; s.pz = a.pz;
        s.Kd = ctx->d_Kd.as<double>(); s.cp = ctx->d_cp.as<double>(); s.ico = a.ico; s.hint = a.hint;
        s.hint_face = ctx->d_hint_face.as<int>(); s.lab12 = ctx->d_lab12.as<float>();
        s.tref = ctx->d_tref.as<float>();
        s.face_feat = CQ == 0 ? ctx->d_face_feat.as<float>() : nullptr;
        s.pcx = ctx->d_pcx.as<float>(); s.pcy = ctx->d_pcy.as<float>(); s.pcz = ctx->d_pcz.as<float>();
        k_cp_hint<<<(Ns + 127) / 128, 128, 0, ctx->stream>>>(s, Ns);
        KLAUNCH_CHECK();
        const size_t setup_smem = (size_t)4 * L * 15 * sizeof(double);   // per warp: K block + output block of one control point
        const int staged = setup_smem <= 48 * 1024 ? 1 : 0;
        k_unary_setup<<<(Ns + 3) / 4, 128, staged ? setup_smem : 0, ctx->stream>>>(s, Ns, staged);
        KLAUNCH_CHECK();
    }
#define LAUNCH_UNARY(Q)                                                                                             \
    do {                                                                                                            \
        if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_unary<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_unary<Q><<<grid, 256, smem, ctx->stream>>>(a);                                                            \
    } while (0)
    {
        BigSmemLock big(smem);
        Timer tmain(ctx, "unary_main");
        switch (CQ) {
            case 0: LAUNCH_UNARY(0); break;
            case 1: LAUNCH_UNARY(1); break;
            case 2: LAUNCH_UNARY(2); break;
            case 3: LAUNCH_UNARY(3); break;
            case 4: LAUNCH_UNARY(4); break;
            case 5: LAUNCH_UNARY(5); break;
            default: FAIL("unsupported channel count");
        }
    }
#undef LAUNCH_UNARY
    KLAUNCH_CHECK();
    return 0;
}

int msm_unary_table(msm_ctx* ctx, double* out) {
    CK(cudaSetDevice(ctx->device));
    const int Ns = ctx->k_end - ctx->k_begin, L = ctx->L, N = ctx->N;
    CK(ctx->d_unary.ensure(std::max<size_t>((size_t)L * Ns, 1) * 4));
    if (msm_unary_table_dev(ctx, ctx->d_unary.as<float>())) return 1;
    if (Ns == 0) return 0;
    CK(ctx->h_pin.ensure((size_t)L * Ns * 4));
    CK(cudaMemcpyAsync(ctx->h_pin.p, ctx->d_unary.p, (size_t)L * Ns * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    {   // a deferred patch build is checked here, where the host waits anyway; the rare repeat recomputes the table
        int redone = 0;
        if (patch_settle(ctx, true, &redone)) return 1;
        if (redone) {
            if (msm_unary_table_dev(ctx, ctx->d_unary.as<float>())) return 1;
            CK(cudaMemcpyAsync(ctx->h_pin.p, ctx->d_unary.p, (size_t)L * Ns * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
    }
    const float* h = ctx->h_pin.as<float>();
    for (int l = 0; l < L; l++) {
        double* o = out + (size_t)l * N + ctx->k_begin;
        const float* hl = h + (size_t)l * Ns;
        for (int r = 0; r < Ns; r++) o[r] = (double)hl[r];
    }
    ctx->h_unary.assign(h, h + (size_t)L * Ns);   // shard rows [L][Ns], as produced (msm_total_cost)
    ctx->h_unary_k0 = ctx->k_begin; ctx->h_unary_ns = Ns;
    ctx->have_unary = true;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
int msm_triplet_costs(msm_ctx* ctx, int n, const int* q, double* out) {
    CK(cudaSetDevice(ctx->device));
    if (check_triplets(ctx)) return 1;
    if (n <= 0) return 0;
    for (int i = 0; i < n; i++) {
        if (q[4 * i] < 0 || q[4 * i] >= ctx->Tn) FAIL("msm_triplet_costs: triplet index out of range");
        for (int j = 1; j < 4; j++)
            if (q[4 * i + j] < 0 || q[4 * i + j] >= ctx->L) FAIL("msm_triplet_costs: label out of range");
    }
    CK(ctx->d_scratch.ensure((size_t)n * 16));
    CK(ctx->d_scratch2.ensure((size_t)n * 8));
    CK(cudaMemcpyAsync(ctx->d_scratch.p, q, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream));
    {
        Timer tm(ctx, "triplet_queries");
        k_triplet_queries<<<(n + 127) / 128, 128, 0, ctx->stream>>>(trip_args(ctx), n, ctx->d_scratch.as<int>(), ctx->d_scratch2.as<double>());
        KLAUNCH_CHECK();
    }
    CK(cudaMemcpyAsync(out, ctx->d_scratch2.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int msm_triplet_table_dev(msm_ctx* ctx, int t0, int t1, float* d_out) {
    CK(cudaSetDevice(ctx->device));
    if (check_triplets(ctx)) return 1;
    if (t0 < 0 || t1 > ctx->Tn || t0 > t1) FAIL("msm_triplet_table: bad triplet range");
    if (t0 == t1) return 0;
    const int L = ctx->L;
    const DevParams dp = dev_params(ctx);
    if (trip_fast_ok(dp) && triplet_table_smem(L) <= 200 * 1024) {
        const size_t smem = triplet_table_smem(L);
        const int nt = t1 - t0;
        BigSmemLock big(smem);
        CK(ctx->d_tprep.ensure((size_t)nt * triplet_rec_bytes(L)));
        unsigned char* recs = ctx->d_tprep.as<unsigned char>();
        Timer tm(ctx, "triplet_table");
        k_triplet_prep<<<(nt + 3) / 4, 128, 0, ctx->stream>>>(trip_args(ctx), t0, nt, recs);
        KLAUNCH_CHECK();
        if (L == 19) {   // the standard label sets: compile-time label count; one CTA per triplet, 6 per SM
            const size_t sm1 = triplet_table_smem(L, 1);
            k_triplet_table_fast<19, 6><<<nt, kTabThreads, sm1, ctx->stream>>>(trip_args(ctx), strain_consts(dp), recs, nt, 1, d_out);
        } else {
            int nsm = 148;
            cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, ctx->device);
            const int per_sm = std::max(1, std::min(5, (int)((200 * 1024) / smem)));
            if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_triplet_table_fast<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_triplet_table_fast<0, 1><<<std::min(nt, nsm * per_sm), kTabThreads, smem, ctx->stream>>>(trip_args(ctx), strain_consts(dp), recs, nt, 2, d_out);
        }
        KLAUNCH_CHECK();
        return 0;
    }
    const size_t smem = (size_t)9 * L * sizeof(double) + (size_t)L * L * L * sizeof(float);
    if (smem > 200 * 1024) FAIL("msm_triplet_table: label set too large for the full table kernel");
    BigSmemLock big(smem);
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_triplet_table, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    Timer tm(ctx, "triplet_table");
    k_triplet_table<<<t1 - t0, 256, smem, ctx->stream>>>(trip_args(ctx), t0, d_out);
    KLAUNCH_CHECK();
    return 0;
}

int msm_triplet_table(msm_ctx* ctx, int t0, int t1, float* out) {
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)std::max(0, t1 - t0) * ctx->L * ctx->L * ctx->L;
    CK(ctx->d_scratch.ensure(std::max<size_t>(n, 1) * 4));
    if (msm_triplet_table_dev(ctx, t0, t1, ctx->d_scratch.as<float>())) return 1;
    if (n) CK(cudaMemcpyAsync(out, ctx->d_scratch.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// The default regulariser settings (strain, k = 2, exponent 2; or the group rules) run the shared-edge kernel of
// triplet.cuh (grid: triplets x labels); every other parameter set the generic one (kernels.cuh trip_cost).
#ifndef MSM_MOVES_AG
#define MSM_MOVES_AG 5
#endif
constexpr int kMovesAG = MSM_MOVES_AG;   // proposal labels per thread of k_triplet_moves_fast (measured best of 1..19)
constexpr int kMovesBlock = 64;
#define LAUNCH_TRIPLET_MOVES(OUT_T, a_, T_, lab_, alpha_, out_, peers_, tfull_)                                             \
    do {                                                                                                                 \
        const DevParams dp_ = dev_params(ctx);                                                                           \
        const int nal_ = (alpha_) < 0 ? ctx->L : 1;                                                                      \
        if (trip_fast_ok(dp_)) {                                                                                         \
            const dim3 grid_((unsigned)(((T_) + kMovesBlock - 1) / kMovesBlock), (unsigned)((nal_ + kMovesAG - 1) / kMovesAG)); \
            k_triplet_moves_fast<OUT_T, kMovesAG><<<grid_, kMovesBlock, 0, ctx->stream>>>(a_, strain_consts(dp_), T_, lab_, alpha_, out_, peers_, tfull_); \
        } else {                                                                                                         \
            const long long n_ = (long long)nal_ * (T_);                                                                 \
            k_triplet_moves<OUT_T, false><<<(unsigned)((n_ + 127) / 128), 128, 0, ctx->stream>>>(a_, T_, lab_, alpha_, out_, peers_, tfull_); \
        }                                                                                                                \
    } while (0)

int msm_triplet_moves_range_dev(msm_ctx* ctx, const int* d_labeling, int alpha, int t0, int t1, float* d_out) {
    CK(cudaSetDevice(ctx->device));
    if (check_triplets(ctx)) return 1;
    if (alpha >= ctx->L) FAIL("msm_triplet_moves: label out of range");
    if (t0 < 0 || t1 > ctx->Tn || t0 > t1) FAIL("msm_triplet_moves: bad triplet range");
    if (t0 == t1) return 0;
    TripArgs a = trip_args(ctx);
    a.t_off = t0;
    Timer tm(ctx, "triplet_moves");
    LAUNCH_TRIPLET_MOVES(float, a, t1 - t0, d_labeling, alpha, d_out, PeerSet{}, 0);
    KLAUNCH_CHECK();
    return 0;
}

int msm_triplet_moves_range_peers(msm_ctx* ctx, const int* d_labeling, int alpha, int t0, int t1, void* const* peer_tables, int world,
                                  void* multicast_table) {
    CK(cudaSetDevice(ctx->device));
    if (check_triplets(ctx)) return 1;
    if (alpha >= ctx->L) FAIL("msm_triplet_moves: label out of range");
    if (t0 < 0 || t1 > ctx->Tn || t0 > t1) FAIL("msm_triplet_moves: bad triplet range");
    PeerSet ps{};
    if (make_peer_set(ctx, peer_tables, world, multicast_table, ps)) return 1;
    if (t0 == t1) return 0;
    TripArgs a = trip_args(ctx);
    a.t_off = t0;
    Timer tm(ctx, "triplet_moves");
    LAUNCH_TRIPLET_MOVES(float, a, t1 - t0, d_labeling, alpha, (float*)nullptr, ps, ctx->Tn);
    KLAUNCH_CHECK();
    return 0;
}

int msm_triplet_moves_dev(msm_ctx* ctx, const int* d_labeling, int alpha, float* d_out) {
    return msm_triplet_moves_range_dev(ctx, d_labeling, alpha, 0, ctx->Tn, d_out);
}

int msm_triplet_moves(msm_ctx* ctx, const int* labeling, int alpha, double* out) {
    CK(cudaSetDevice(ctx->device));
    if (check_triplets(ctx)) return 1;
    if (alpha >= ctx->L) FAIL("msm_triplet_moves: label out of range");
    for (int i = 0; i < ctx->N; i++)
        if (labeling[i] < 0 || labeling[i] >= ctx->L) FAIL("msm_triplet_moves: labeling out of range");
    const long long n = (long long)(alpha < 0 ? ctx->L : 1) * ctx->Tn * 8;
    CK(ctx->d_scratch.ensure((size_t)ctx->N * 4));
    CK(ctx->d_scratch2.ensure((size_t)n * 8));
    CK(cudaMemcpyAsync(ctx->d_scratch.p, labeling, (size_t)ctx->N * 4, cudaMemcpyHostToDevice, ctx->stream));
    {
        Timer tm(ctx, "triplet_moves");
        LAUNCH_TRIPLET_MOVES(double, trip_args(ctx), ctx->Tn, ctx->d_scratch.as<int>(), alpha, ctx->d_scratch2.as<double>(), PeerSet{}, 0);
        KLAUNCH_CHECK();
    }
    CK(cudaMemcpyAsync(out, ctx->d_scratch2.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// the same table in fp32 (the stated tolerance of the cost tables is 1e-5 relative): half the read-back
int msm_triplet_moves_f32(msm_ctx* ctx, const int* labeling, int alpha, float* out) {
    CK(cudaSetDevice(ctx->device));
    if (check_triplets(ctx)) return 1;
    if (alpha >= ctx->L) FAIL("msm_triplet_moves: label out of range");
    for (int i = 0; i < ctx->N; i++)
        if (labeling[i] < 0 || labeling[i] >= ctx->L) FAIL("msm_triplet_moves: labeling out of range");
    const long long n = (long long)(alpha < 0 ? ctx->L : 1) * ctx->Tn * 8;
    CK(ctx->d_scratch.ensure((size_t)ctx->N * 4));
    CK(ctx->d_scratch2.ensure((size_t)n * 4));
    CK(cudaMemcpyAsync(ctx->d_scratch.p, labeling, (size_t)ctx->N * 4, cudaMemcpyHostToDevice, ctx->stream));
    {
        Timer tm(ctx, "triplet_moves");
        LAUNCH_TRIPLET_MOVES(float, trip_args(ctx), ctx->Tn, ctx->d_scratch.as<int>(), alpha, ctx->d_scratch2.as<float>(), PeerSet{}, 0);
        KLAUNCH_CHECK();
    }
    CK(cudaMemcpyAsync(out, ctx->d_scratch2.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
static int check_pairs(msm_ctx* ctx) {
    if (ctx->P <= 0) FAIL("set_pairs must be called first");
    if (ctx->Tcp <= 0) FAIL("pairwise regulariser needs the control grid faces (fold test)");
    if (ctx->par.mvdmax <= 0) FAIL("pairwise regulariser needs MVDmax (set_spacings)");
    return ensure_label_rot(ctx);
}

int msm_pair_table_dev(msm_ctx* ctx, float* d_out) {
    CK(cudaSetDevice(ctx->device));
    if (check_pairs(ctx)) return 1;
    const long long n = (long long)ctx->P * ctx->L * ctx->L;
    Timer tm(ctx, "pair_table");
    k_pair_table<float><<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(pair_args(ctx), d_out);
    KLAUNCH_CHECK();
    return 0;
}

int msm_pair_table(msm_ctx* ctx, float* out) {
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)ctx->P * ctx->L * ctx->L;
    CK(ctx->d_scratch.ensure(std::max<size_t>(n, 1) * 4));
    if (msm_pair_table_dev(ctx, ctx->d_scratch.as<float>())) return 1;
    CK(cudaMemcpyAsync(out, ctx->d_scratch.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
int msm_set_group_patches(msm_ctx* ctx, long long nmaps, const long long* off, const int* keys, const double* vals) {
    CK(cudaSetDevice(ctx->device));
    if (nmaps != (long long)ctx->N * ctx->L) FAIL("msm_set_group_patches: need one map per (global node, label)");
    for (long long m = 0; m < nmaps; m++)
        for (long long e = off[m] + 1; e < off[m + 1]; e++)
            if (keys[e - 1] >= keys[e]) FAIL("msm_set_group_patches: keys must be strictly ascending within a map");
    ctx->n_maps = nmaps;
    if (upload(ctx, ctx->d_moff, off, (size_t)nmaps + 1)) return 1;
    if (upload(ctx, ctx->d_mkey, keys, (size_t)off[nmaps])) return 1;
    if (upload(ctx, ctx->d_mval, vals, (size_t)off[nmaps])) return 1;
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

static int group_pair_args(msm_ctx* ctx, GroupPairArgs& a) {
    if (ctx->P <= 0) FAIL("set_pairs must be called first");
    if (ctx->n_maps != (long long)ctx->N * ctx->L) FAIL("set_patch_data (msm_set_group_patches) must be called first");
    a.L = ctx->L; a.N_subj = ctx->N_subj;
    a.pairs = ctx->d_pairs.as<int>();
    a.moff = ctx->d_moff.as<long long>(); a.mkey = ctx->d_mkey.as<int>(); a.mval = ctx->d_mval.as<double>();
    return 0;
}

int msm_group_pair_costs(msm_ctx* ctx, int n, const int* q, double* out) {
    CK(cudaSetDevice(ctx->device));
    GroupPairArgs a;
    if (group_pair_args(ctx, a)) return 1;
    if (n <= 0) return 0;
    for (int i = 0; i < n; i++)
        if (q[3 * i] < 0 || q[3 * i] >= ctx->P || q[3 * i + 1] < 0 || q[3 * i + 1] >= ctx->L || q[3 * i + 2] < 0 || q[3 * i + 2] >= ctx->L)
            FAIL("msm_group_pair_costs: query out of range");
    CK(ctx->d_scratch.ensure((size_t)n * 12));
    CK(ctx->d_scratch2.ensure((size_t)n * 8));
    CK(cudaMemcpyAsync(ctx->d_scratch.p, q, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream));
    {
        Timer tm(ctx, "group_pair");
        k_group_pair_queries<<<(n + 127) / 128, 128, 0, ctx->stream>>>(a, n, ctx->d_scratch.as<int>(), ctx->d_scratch2.as<double>());
        KLAUNCH_CHECK();
    }
    CK(cudaMemcpyAsync(out, ctx->d_scratch2.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int msm_group_pair_table(msm_ctx* ctx, float* out) {
    CK(cudaSetDevice(ctx->device));
    GroupPairArgs a;
    if (group_pair_args(ctx, a)) return 1;
    const long long n = (long long)ctx->P * ctx->L * ctx->L;
    CK(ctx->d_scratch2.ensure((size_t)n * 4));
    {
        Timer tm(ctx, "group_pair");
        k_group_pair_table<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(a, ctx->P, ctx->d_scratch2.as<float>());
        KLAUNCH_CHECK();
    }
    CK(cudaMemcpyAsync(out, ctx->d_scratch2.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int msm_group_pair_moves(msm_ctx* ctx, const int* labeling, int alpha, double* out) {
    CK(cudaSetDevice(ctx->device));
    GroupPairArgs a;
    if (group_pair_args(ctx, a)) return 1;
    if (alpha < 0 || alpha >= ctx->L) FAIL("msm_group_pair_moves: label out of range");
    for (int i = 0; i < ctx->N; i++)
        if (labeling[i] < 0 || labeling[i] >= ctx->L) FAIL("msm_group_pair_moves: labeling out of range");
    CK(ctx->d_scratch.ensure((size_t)ctx->N * 4));
    CK(ctx->d_scratch2.ensure((size_t)ctx->P * 4 * 8));
    CK(cudaMemcpyAsync(ctx->d_scratch.p, labeling, (size_t)ctx->N * 4, cudaMemcpyHostToDevice, ctx->stream));
    {
        Timer tm(ctx, "group_pair");
        const long long n = (long long)ctx->P * 4;
        k_group_pair_moves<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(a, ctx->P, ctx->d_scratch.as<int>(), alpha,
                                                                             ctx->d_scratch2.as<double>());
        KLAUNCH_CHECK();
    }
    CK(cudaMemcpyAsync(out, ctx->d_scratch2.p, (size_t)ctx->P * 4 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
int msm_locate(msm_ctx* ctx, const double* pts, int M, int* tri, double* w) {
    CK(cudaSetDevice(ctx->device));
    if (!ctx->have_target) FAIL("msm_set_target must be called first");
    if (M <= 0) return 0;
    CK(ctx->d_scratch.ensure((size_t)M * 24));
    CK(ctx->d_scratch2.ensure((size_t)M * (24 + 4)));
    double* d_w = ctx->d_scratch2.as<double>();
    int* d_tri = reinterpret_cast<int*>(d_w + (size_t)3 * M);
    CK(cudaMemcpyAsync(ctx->d_scratch.p, pts, (size_t)M * 24, cudaMemcpyHostToDevice, ctx->stream));
    {
        Timer tm(ctx, "locate");
        k_locate<<<(M + 255) / 256, 256, 0, ctx->stream>>>(ico_dev(ctx), M, ctx->d_scratch.as<double>(), ctx->d_leaf_user.as<int>(),
                                                          ctx->d_leaf_perm.as<unsigned char>(), d_tri, d_w);
        KLAUNCH_CHECK();
    }
    CK(cudaMemcpyAsync(tri, d_tri, (size_t)M * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (w) CK(cudaMemcpyAsync(w, d_w, (size_t)M * 24, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// barycentric_mesh_interpolation ([EXT] A8; src/mesh_registration.cpp:390,305-306; src/DiscreteModel.h:123)
int msm_mesh_interpolate(msm_ctx* ctx, const double* pts, int M, const double* from_xyz, int V, const int* tris, int F, const double* to_xyz,
                         double radius, double* out, int* tri_out, double* w_out) {
    CK(cudaSetDevice(ctx->device));
    if (M <= 0) return 0;
    if (V <= 0 || F <= 0) FAIL("msm_mesh_interpolate: empty mesh");
    // voxel grid over unit directions, cell ~ mean edge (its size and the mesh's orientation from one host pass over the faces)
    std::vector<double> u((size_t)V * 3);
    for (int i = 0; i < V; i++) {
        const double n = std::sqrt(from_xyz[3 * i] * from_xyz[3 * i] + from_xyz[3 * i + 1] * from_xyz[3 * i + 1] + from_xyz[3 * i + 2] * from_xyz[3 * i + 2]);
        if (!(n > 0)) FAIL("msm_mesh_interpolate: vertex at the origin");
        for (int d = 0; d < 3; d++) u[3 * i + d] = from_xyz[3 * i + d] / n;
    }
    double mean_edge = 0, signed_volume = 0;
    for (int t = 0; t < F; t++) {
        for (int j = 0; j < 3; j++)
            if (tris[3 * t + j] < 0 || tris[3 * t + j] >= V) FAIL("msm_mesh_interpolate: triangle index out of range");
        const double* a = &u[3 * tris[3 * t]]; const double* b = &u[3 * tris[3 * t + 1]]; const double* c = &u[3 * tris[3 * t + 2]];
        mean_edge += std::sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]));
        signed_volume += a[0] * (b[1] * c[2] - b[2] * c[1]) + a[1] * (b[2] * c[0] - b[0] * c[2]) + a[2] * (b[0] * c[1] - b[1] * c[0]);
    }
    mean_edge /= F;
    // overall orientation of the mesh: faces wound inward as a whole (negative enclosed volume) are read the other way round;
    // single flipped faces of a folded mesh do not change the sign of the sum
    const double orient = signed_volume < 0.0 ? -1.0 : 1.0;
    const int G = std::max(4, std::min(192, (int)std::floor(2.0 / std::max(mean_edge, 1e-6))));
    const double h = 2.0 / G;
    if (upload(ctx, ctx->d_w_from, from_xyz, (size_t)V * 3)) return 1;
    if (upload(ctx, ctx->d_w_to, to_xyz, (size_t)V * 3)) return 1;
    if (upload(ctx, ctx->d_w_tris, tris, (size_t)F * 3)) return 1;
    if (upload(ctx, ctx->d_w_pts, pts, (size_t)M * 3)) return 1;
    CK(ctx->d_w_out.ensure((size_t)M * (24 + 24 + 4)));
    {   // voxel grid of the triangles, on the device: count, scan, fill, cell lists into ascending triangle order
        const size_t nc = (size_t)G * G * G;
        CK(ctx->d_w_counts.ensure(nc * 4)); CK(ctx->d_w_start.ensure((nc + 1) * 4)); CK(ctx->d_w_fill.ensure(nc * 4));
        CK(cudaMemsetAsync(ctx->d_w_counts.p, 0, nc * 4, ctx->stream));
        CK(cudaMemsetAsync(ctx->d_w_fill.p, 0, nc * 4, ctx->stream));
        Timer tm(ctx, "mesh_grid");
        k_tri_cells<false><<<(F + 127) / 128, 128, 0, ctx->stream>>>(F, G, 1.0 / h, ctx->d_w_from.as<double>(), ctx->d_w_tris.as<int>(),
                                                                    ctx->d_w_counts.as<int>(), nullptr, nullptr, nullptr);
        KLAUNCH_CHECK();
        const int nchunks = 148, chunk = ((int)nc + nchunks - 1) / nchunks;
        CK(ctx->d_scan_sums.ensure((size_t)nchunks * 4));
        k_scan_chunk_sums<<<nchunks, 256, 0, ctx->stream>>>((int)nc, chunk, ctx->d_w_counts.as<int>(), ctx->d_scan_sums.as<int>());
        KLAUNCH_CHECK();
        k_scan_chunk_apply<<<nchunks, 256, 0, ctx->stream>>>((int)nc, chunk, ctx->d_w_counts.as<int>(), ctx->d_scan_sums.as<int>(), ctx->d_w_start.as<int>());
        KLAUNCH_CHECK();
        // the length of the cell lists sizes their buffer (one 4-byte read-back; the point set's results are awaited below anyway)
        int total = 0;
        CK(cudaMemcpyAsync(&total, ctx->d_w_start.as<int>() + nc, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        CK(ctx->d_w_items.ensure((size_t)std::max(total, 1) * 4)); CK(ctx->d_w_items2.ensure((size_t)std::max(total, 1) * 4));
        k_tri_cells<true><<<(F + 127) / 128, 128, 0, ctx->stream>>>(F, G, 1.0 / h, ctx->d_w_from.as<double>(), ctx->d_w_tris.as<int>(), nullptr,
                                                                   ctx->d_w_start.as<int>(), ctx->d_w_fill.as<int>(), ctx->d_w_items2.as<int>());
        KLAUNCH_CHECK();
        k_cell_rank_sort<<<(unsigned)((nc + 3) / 4), 128, 0, ctx->stream>>>((int)nc, ctx->d_w_start.as<int>(), ctx->d_w_items2.as<int>(), ctx->d_w_items.as<int>());
        KLAUNCH_CHECK();
    }
    WarpArgs a;
    a.M = M; a.V = V; a.F = F;
    a.pts = ctx->d_w_pts.as<double>(); a.from = ctx->d_w_from.as<double>(); a.to = ctx->d_w_to.as<double>(); a.tris = ctx->d_w_tris.as<int>();
    a.grid.G = G; a.grid.inv_h = 1.0 / h; a.grid.cell_start = ctx->d_w_start.as<int>(); a.grid.items = ctx->d_w_items.as<int>();
    a.radius = radius; a.orient = orient;
    a.out = ctx->d_w_out.as<double>();
    a.w_out = w_out ? a.out + (size_t)3 * M : nullptr;
    a.tri_out = tri_out ? reinterpret_cast<int*>(a.out + (size_t)6 * M) : nullptr;
    {
        Timer tm(ctx, "mesh_interpolate");
        k_mesh_interpolate<<<(M + 127) / 128, 128, 0, ctx->stream>>>(a);
        KLAUNCH_CHECK();
    }
    CK(cudaMemcpyAsync(out, a.out, (size_t)M * 24, cudaMemcpyDeviceToHost, ctx->stream));
    if (w_out) CK(cudaMemcpyAsync(w_out, a.w_out, (size_t)M * 24, cudaMemcpyDeviceToHost, ctx->stream));
    if (tri_out) CK(cudaMemcpyAsync(tri_out, a.tri_out, (size_t)M * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------------------
// Feature pre-processing (src/featurespace.cc:18-108, src/reg_tools.cpp:750-810) — features.cuh
int msm_smooth_data(msm_ctx* ctx, const double* in_xyz, int Vin, const double* values, int D, const double* mask, const double* ref_xyz, int Vref,
                    double sigma, double* out) {
    CK(cudaSetDevice(ctx->device));
    if (Vin <= 0 || Vref <= 0 || D <= 0) FAIL("msm_smooth_data: empty input");
    if (!(sigma > 0.0)) FAIL("msm_smooth_data: sigma must be positive (the caller skips the smoothing otherwise, src/featurespace.cc:48)");
    const double RAD = 100.0;
    const double chord_max = 2.0 * RAD * std::sin(std::asin(std::min(1.0, sigma / RAD)));   // angular support 2 asin(sigma / RAD)
    double bound = 0.0;
    for (size_t i = 0; i < (size_t)Vin * 3; i++) bound = std::max(bound, std::fabs(in_xyz[i]));
    bound = bound * (1.0 + 1e-9) + 1e-6;
    const int G = std::max(1, std::min(64, (int)std::floor(2.0 * bound / std::max(chord_max, 1e-9))));
    const size_t nc = (size_t)G * G * G;
    if (upload(ctx, ctx->d_f_in, in_xyz, (size_t)Vin * 3)) return 1;
    if (upload(ctx, ctx->d_f_ref, ref_xyz, (size_t)Vref * 3)) return 1;
    if (upload(ctx, ctx->d_f_vals, values, (size_t)D * Vin)) return 1;
    if (mask && upload(ctx, ctx->d_f_mask, mask, (size_t)Vin)) return 1;
    CK(ctx->d_f_out.ensure((size_t)D * Vref * 8));
    CK(ctx->d_f_cell_counts.ensure(nc * 4)); CK(ctx->d_f_cell_start.ensure((nc + 1) * 4)); CK(ctx->d_f_cell_fill.ensure(nc * 4));
    CK(ctx->d_f_cell_of.ensure((size_t)Vin * 4)); CK(ctx->d_f_sorted.ensure((size_t)Vin * 4)); CK(ctx->d_f_sorted2.ensure((size_t)Vin * 4));
    CK(cudaMemsetAsync(ctx->d_f_cell_counts.p, 0, nc * 4, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_f_cell_fill.p, 0, nc * 4, ctx->stream));
    {
    Timer tm(ctx, "smooth");
    BinGrid bg{};
    bg.G = G; bg.lo = -bound; bg.inv_h = G / (2.0 * bound);
    bg.cell_start = ctx->d_f_cell_start.as<int>(); bg.cell_fill = ctx->d_f_cell_fill.as<int>(); bg.sorted = ctx->d_f_sorted.as<int>();
    k_bin_count<<<(Vin + 255) / 256, 256, 0, ctx->stream>>>(bg, Vin, ctx->d_f_in.as<double>(), ctx->d_f_cell_of.as<int>(), ctx->d_f_cell_counts.as<int>());
    KLAUNCH_CHECK();
    {
        const int nchunks = 148, chunk = ((int)nc + nchunks - 1) / nchunks;
        CK(ctx->d_f_scan.ensure((size_t)nchunks * 4));
        k_scan_chunk_sums<<<nchunks, 256, 0, ctx->stream>>>((int)nc, chunk, ctx->d_f_cell_counts.as<int>(), ctx->d_f_scan.as<int>());
        KLAUNCH_CHECK();
        k_scan_chunk_apply<<<nchunks, 256, 0, ctx->stream>>>((int)nc, chunk, ctx->d_f_cell_counts.as<int>(), ctx->d_f_scan.as<int>(),
                                                            ctx->d_f_cell_start.as<int>());
        KLAUNCH_CHECK();
    }
    k_bin_fill<<<(Vin + 255) / 256, 256, 0, ctx->stream>>>(bg, Vin, ctx->d_f_cell_of.as<int>());
    KLAUNCH_CHECK();
    k_cell_rank_sort<<<(unsigned)((nc + 3) / 4), 128, 0, ctx->stream>>>((int)nc, bg.cell_start, ctx->d_f_sorted.as<int>(), ctx->d_f_sorted2.as<int>());
    KLAUNCH_CHECK();
    bg.sorted = ctx->d_f_sorted2.as<int>();
    SmoothArgs a;
    a.g = bg; a.Vin = Vin; a.Vref = Vref; a.D = D;
    a.in_xyz = ctx->d_f_in.as<double>(); a.ref_xyz = ctx->d_f_ref.as<double>(); a.values = ctx->d_f_vals.as<double>();
    a.mask = mask ? ctx->d_f_mask.as<double>() : nullptr;
    a.chord_max = chord_max; a.sigma = sigma; a.out = ctx->d_f_out.as<double>();
    const unsigned grid = (unsigned)((Vref + 3) / 4);
    if (D == 1) k_smooth<1><<<grid, 128, 0, ctx->stream>>>(a);
    else if (D == 2) k_smooth<2><<<grid, 128, 0, ctx->stream>>>(a);
    else k_smooth<4><<<grid, 128, 0, ctx->stream>>>(a);
    KLAUNCH_CHECK();
    }
    CK(cudaMemcpyAsync(out, ctx->d_f_out.p, (size_t)D * Vref * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int msm_histogram_match(msm_ctx* ctx, double* data, int V, const double* data_ref, int Vref, int D, const double* excl, int excl_rows,
                        const double* excl_ref, int excl_ref_rows, int bins) {
    CK(cudaSetDevice(ctx->device));
    if (V <= 0 || Vref <= 0 || D <= 0) FAIL("msm_histogram_match: empty input");
    if (bins < 2 || bins > 2048) FAIL("msm_histogram_match: bins must be in [2, 2048] (both CDFs are staged in 48 KB of shared memory; the reference uses 256)");
    if (V >= (1 << 24) || Vref >= (1 << 24)) FAIL("msm_histogram_match: more than 2^24 values per dimension (the reference counts them in a float)");
    if ((excl && excl_rows < 1) || (excl_ref && excl_ref_rows < 1)) FAIL("msm_histogram_match: mask rows");
    if (upload(ctx, ctx->d_f_vals, data, (size_t)D * V)) return 1;
    if (upload(ctx, ctx->d_f_ref, data_ref, (size_t)D * Vref)) return 1;
    if (excl && upload(ctx, ctx->d_f_mask, excl, (size_t)excl_rows * V)) return 1;
    if (excl_ref && upload(ctx, ctx->d_f_mask2, excl_ref, (size_t)excl_ref_rows * Vref)) return 1;
    CK(ctx->d_f_stats.ensure((size_t)2 * D * sizeof(HistStats)));
    CK(ctx->d_f_cdf.ensure((size_t)2 * D * bins * 8));
    {
    Timer tm(ctx, "histmatch");
    k_hist_stats<<<2 * D, 1024, (size_t)bins * 4, ctx->stream>>>(D, bins, ctx->d_f_vals.as<double>(), V, excl ? ctx->d_f_mask.as<double>() : nullptr,
                                                                excl_rows, ctx->d_f_ref.as<double>(), Vref,
                                                                excl_ref ? ctx->d_f_mask2.as<double>() : nullptr, excl_ref_rows,
                                                                ctx->d_f_stats.as<HistStats>(), ctx->d_f_cdf.as<double>());
    KLAUNCH_CHECK();
    k_hist_match<<<dim3((unsigned)((V + 255) / 256), (unsigned)D), 256, (size_t)2 * bins * 8, ctx->stream>>>(
        D, bins, ctx->d_f_vals.as<double>(), V, excl ? ctx->d_f_mask.as<double>() : nullptr, excl_rows, ctx->d_f_stats.as<HistStats>(),
        ctx->d_f_cdf.as<double>());
    KLAUNCH_CHECK();
    }
    CK(cudaMemcpyAsync(data, ctx->d_f_vals.p, (size_t)D * V * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int msm_variance_normalise(msm_ctx* ctx, double* data, int D, int V, const double* mask) {
    CK(cudaSetDevice(ctx->device));
    if (V <= 0 || D <= 0) FAIL("msm_variance_normalise: empty input");
    if (upload(ctx, ctx->d_f_vals, data, (size_t)D * V)) return 1;
    if (mask && upload(ctx, ctx->d_f_mask, mask, (size_t)V)) return 1;
    {
        Timer tm(ctx, "varnorm");
        k_varnorm<<<D, 1024, 0, ctx->stream>>>(V, ctx->d_f_vals.as<double>(), mask ? ctx->d_f_mask.as<double>() : nullptr);
        KLAUNCH_CHECK();
    }
    CK(cudaMemcpyAsync(data, ctx->d_f_vals.p, (size_t)D * V * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int msm_create_exclusion(msm_ctx* ctx, const double* data, int D, int V, float lower, float upper, double* mask_out) {
    CK(cudaSetDevice(ctx->device));
    if (V <= 0 || D <= 0) FAIL("msm_create_exclusion: empty input");
    if (upload(ctx, ctx->d_f_vals, data, (size_t)D * V)) return 1;
    CK(ctx->d_f_mask.ensure((size_t)V * 8));
    k_exclusion<<<(V + 255) / 256, 256, 0, ctx->stream>>>(D, V, ctx->d_f_vals.as<double>(), (double)lower, (double)upper, ctx->d_f_mask.as<double>());
    KLAUNCH_CHECK();
    CK(cudaMemcpyAsync(mask_out, ctx->d_f_mask.p, (size_t)V * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int msm_total_cost(msm_ctx* ctx, const int* labeling, double sums[3]) {
    CK(cudaSetDevice(ctx->device));
    sums[0] = sums[1] = sums[2] = 0.0;
    for (int i = 0; i < ctx->N; i++)
        if (labeling[i] < 0 || labeling[i] >= ctx->L) FAIL("msm_total_cost: labeling out of range");
    if (!ctx->par.group) {
        // unary term: needs the host copy of the whole table that msm_unary_table leaves behind (the device-resident and
        // sharded variants do not); the group model has no unary term (src/DiscreteCostFunction.h:34)
        if (ctx->Vs > 0 && ctx->have_target && ctx->Ct > 0) {
            if (!ctx->have_unary)
                FAIL("msm_total_cost: no unary table — call msm_unary_table first (msm_unary_table_dev / _peers keep no host copy, and "
                     "set_cpgrid / set_labels / set_shard / build_patches invalidate it)");
            if (ctx->h_unary_k0 != 0 || ctx->h_unary_ns != ctx->N)
                FAIL("msm_total_cost: the unary table covers only this context's shard; sum the shards' energies on the caller's side");
            for (int r = 0; r < ctx->h_unary_ns; r++)
                sums[0] += (double)ctx->h_unary[(size_t)labeling[ctx->h_unary_k0 + r] * ctx->h_unary_ns + r];
        }
    }
    if (ctx->P > 0 && !ctx->par.group) {
        if (check_pairs(ctx)) return 1;
        std::vector<int> q((size_t)ctx->P * 3);
        std::vector<int> hp((size_t)ctx->P * 2);
        CK(cudaMemcpy(hp.data(), ctx->d_pairs.p, hp.size() * 4, cudaMemcpyDeviceToHost));
        for (int p = 0; p < ctx->P; p++) { q[3 * p] = p; q[3 * p + 1] = labeling[hp[2 * p]]; q[3 * p + 2] = labeling[hp[2 * p + 1]]; }
        CK(ctx->d_scratch.ensure(q.size() * 4));
        CK(ctx->d_scratch2.ensure((size_t)ctx->P * 8));
        CK(cudaMemcpyAsync(ctx->d_scratch.p, q.data(), q.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
        k_pair_queries<<<(ctx->P + 127) / 128, 128, 0, ctx->stream>>>(pair_args(ctx), ctx->P, ctx->d_scratch.as<int>(), ctx->d_scratch2.as<double>());
        KLAUNCH_CHECK();
        std::vector<double> r(ctx->P);
        CK(cudaMemcpyAsync(r.data(), ctx->d_scratch2.p, (size_t)ctx->P * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (double v : r) sums[1] += v;
    }
    if (ctx->P > 0 && ctx->par.group) {
        // group data term (src/DiscreteGroupCostFunction.cpp:32-54) over the labelling
        if (ctx->n_maps != (long long)ctx->N * ctx->L) FAIL("msm_total_cost: group model with pairs but no patch data (msm_set_group_patches)");
        std::vector<int> q((size_t)ctx->P * 3), hp((size_t)ctx->P * 2);
        CK(cudaMemcpy(hp.data(), ctx->d_pairs.p, hp.size() * 4, cudaMemcpyDeviceToHost));
        for (int p = 0; p < ctx->P; p++) { q[3 * p] = p; q[3 * p + 1] = labeling[hp[2 * p]]; q[3 * p + 2] = labeling[hp[2 * p + 1]]; }
        std::vector<double> r(ctx->P);
        if (msm_group_pair_costs(ctx, ctx->P, q.data(), r.data())) return 1;
        for (double v : r) sums[1] += v;
    }
    if (ctx->Tn > 0) {
        std::vector<int> ht((size_t)ctx->Tn * 3), q((size_t)ctx->Tn * 4);
        CK(cudaMemcpy(ht.data(), ctx->d_trip.p, ht.size() * 4, cudaMemcpyDeviceToHost));
        for (int t = 0; t < ctx->Tn; t++) {
            q[4 * t] = t; q[4 * t + 1] = labeling[ht[3 * t]]; q[4 * t + 2] = labeling[ht[3 * t + 1]]; q[4 * t + 3] = labeling[ht[3 * t + 2]];
        }
        std::vector<double> r(ctx->Tn);
        if (msm_triplet_costs(ctx, ctx->Tn, q.data(), r.data())) return 1;
        for (double v : r) sums[2] += v;
    }
    return 0;
}

long long msm_launch_count(const msm_ctx* ctx) { return ctx->launches; }
int msm_set_timing(msm_ctx* ctx, int mode) { ctx->timing_mode = mode; return 0; }
int msm_timing_collect(msm_ctx* ctx, const char* family, double* total_ms, long long* count) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    *total_ms = 0.0; *count = 0;
    std::vector<msm_ctx::PendingEv> keep;
    for (auto& p : ctx->ev_pending) {
        if (std::strcmp(p.name, family) == 0) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) { *total_ms += ms; (*count)++; }
            ctx->ev_free.push_back(p.e0); ctx->ev_free.push_back(p.e1);
        } else keep.push_back(p);
    }
    ctx->ev_pending.swap(keep);
    return 0;
}
double msm_last_kernel_ms(const msm_ctx* ctx, const char* family) {
    auto it = ctx->last_ms.find(family);
    return it == ctx->last_ms.end() ? -1.0 : (double)it->second;
}

}  // extern "C"
