// ORACLE — TEST INFRASTRUCTURE ONLY.  Flat C entry points over oracle.h for ctypes (oracle/pyorc.py).
#include <cstring>

#include "oracle.h"
#include "feat.h"

using namespace orc;

static Mesh mesh_from(const double* xyz, int nv, const int* tris, int nf) {
    Mesh M;
    M.V.resize(nv);
    for (int i = 0; i < nv; i++) M.V[i] = Pt(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    M.F.resize(nf);
    for (int i = 0; i < nf; i++) M.F[i] = {tris[3 * i], tris[3 * i + 1], tris[3 * i + 2]};
    if (nf > 0) M.build_adjacency();
    return M;
}
static void mat_to(const Mat3& R, double* o) { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o[3 * i + j] = R.m[i][j]; }
static Mat3 mat_from(const double* o) { Mat3 R; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R.m[i][j] = o[3 * i + j]; return R; }

extern "C" {

// ---------------- geometry / model glue ----------------
void orc_icosphere_counts(int level, int* nv, int* nf) {
    long f = 20; for (int i = 0; i < level; i++) f *= 4;
    *nf = (int)f; *nv = (int)(f / 2 + 2);
}
void orc_icosphere(int level, double radius, double* xyz, int* tris) {
    Mesh M = make_mesh_from_icosa(level, radius);
    for (int i = 0; i < M.nvertices(); i++) { xyz[3 * i] = M.V[i].X; xyz[3 * i + 1] = M.V[i].Y; xyz[3 * i + 2] = M.V[i].Z; }
    for (int i = 0; i < M.ntriangles(); i++) { tris[3 * i] = M.F[i][0]; tris[3 * i + 1] = M.F[i][1]; tris[3 * i + 2] = M.F[i][2]; }
}
void orc_rotation_matrix(const double* a, const double* b, double* out9) {
    mat_to(estimate_rotation_matrix(Pt(a[0], a[1], a[2]), Pt(b[0], b[1], b[2])), out9);
}
// label grid: outputs up to cap points each; returns counts via ns/nb.
void orc_label_grid(int SGres, double maxs_dist, int cap, double* samples, int* ns, double* bary, int* nb, double* centre) {
    LabelGrid LG = label_sampling_grid(SGres, maxs_dist);
    *ns = (int)LG.samples.size(); *nb = (int)LG.barycentres.size();
    for (int i = 0; i < std::min(cap, *ns); i++) { samples[3 * i] = LG.samples[i].X; samples[3 * i + 1] = LG.samples[i].Y; samples[3 * i + 2] = LG.samples[i].Z; }
    for (int i = 0; i < std::min(cap, *nb); i++) { bary[3 * i] = LG.barycentres[i].X; bary[3 * i + 1] = LG.barycentres[i].Y; bary[3 * i + 2] = LG.barycentres[i].Z; }
    centre[0] = LG.centre.X; centre[1] = LG.centre.Y; centre[2] = LG.centre.Z;
}
void orc_spacings(const double* xyz, int nv, const int* tris, int nf, double* MAXSEP, double* MVDmax, double* MVD) {
    Mesh M = mesh_from(xyz, nv, tris, nf);
    std::vector<double> ms;
    spacings(M, ms, *MVDmax, *MVD);
    std::memcpy(MAXSEP, ms.data(), sizeof(double) * nv);
}
void orc_triplets(const double* xyz, int nv, const int* tris, int nf, int* out) {
    Mesh M = mesh_from(xyz, nv, tris, nf);
    auto t = estimate_triplets(M);
    std::memcpy(out, t.data(), sizeof(int) * t.size());
}
int orc_pairs(const double* xyz, int nv, const int* tris, int nf, int* out, int cap_pairs) {
    Mesh M = mesh_from(xyz, nv, tris, nf);
    auto p = estimate_pairs(M);
    int np = (int)p.size() / 2;
    if (out) std::memcpy(out, p.data(), sizeof(int) * 2 * std::min(np, cap_pairs));
    return np;
}
void orc_rotations(const double* centre, const double* cps, int n, double* out) {
    Pt c(centre[0], centre[1], centre[2]);
    for (int k = 0; k < n; k++) mat_to(estimate_rotation_matrix(c, Pt(cps[3 * k], cps[3 * k + 1], cps[3 * k + 2])), out + 9 * k);
}
double orc_mesh_meanangle(const double* xyz, int nv, const int* tris, int nf) {  // DiscreteCostFunction.cpp:95-102
    Mesh M = mesh_from(xyz, nv, tris, nf);
    auto angles = M.get_face_angles();
    double meanang = 0.0;
    for (const auto& a : angles) for (double j : a) meanang += j;
    return meanang / (3.0 * angles.size());
}
double orc_corr(const double* a, const double* b, const double* w, int n, int has_w) {
    std::vector<double> A(a, a + n), B(b, b + n), W;
    if (has_w) W.assign(w, w + n);
    return corr(A, B, W);
}
// which: 0 literal (reg_tools.cpp:703-748), 1 gram form
double orc_strain(const double* t0, const double* t1, double mu, double kappa, double k_exp, int which) {
    Tri T0(Pt(t0[0], t0[1], t0[2]), Pt(t0[3], t0[4], t0[5]), Pt(t0[6], t0[7], t0[8]));
    Tri T1(Pt(t1[0], t1[1], t1[2]), Pt(t1[3], t1[4], t1[5]), Pt(t1[6], t1[7], t1[8]));
    return which == 0 ? calculate_triangular_strain(T0, T1, mu, kappa, k_exp) : strain_gram(T0, T1, mu, kappa, k_exp);
}
// point location + barycentric weights. brute=1 uses the definition (all triangles).
void orc_locate(const double* xyz, int nv, const int* tris, int nf, const double* pts, int m, int* tri_out, double* w_out,
                int brute, int orthogonal) {
    Mesh M = mesh_from(xyz, nv, tris, nf);
    SphereLocator loc;
    if (!brute) loc.build(M);
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < m; i++) {
        Pt p(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
        int t = brute ? brute_closest_triangle(M, p) : loc.closest_triangle(p);
        tri_out[i] = t;
        if (w_out && t >= 0) barycentric_weights(M.V[M.F[t][0]], M.V[M.F[t][1]], M.V[M.F[t][2]], p, w_out + 3 * i, orthogonal != 0);
    }
}

// A8: metric_resample (adaptive barycentric), values[d*nv_in + j] -> out[d*nv_ref + k]
void orc_metric_resample(const double* in_xyz, int nv_in, const int* in_tris, int nf_in, const double* values, int dims, const double* ref_xyz,
                         int nv_ref, const int* ref_tris, int nf_ref, double* out) {
    Mesh A = mesh_from(in_xyz, nv_in, in_tris, nf_in), B = mesh_from(ref_xyz, nv_ref, ref_tris, nf_ref);
    std::vector<double> v(values, values + (size_t)dims * nv_in);
    const std::vector<double> r = metric_resample(A, v, dims, B);
    std::copy(r.begin(), r.end(), out);
}

// ---------------- feature pre-processing (oracle/feat.h) ----------------
void orc_smooth_data(const double* in_xyz, int nv_in, const double* values, int dims, const double* ref_xyz, int nv_ref, double sigma,
                     const double* mask, double* out) {
    std::vector<Pt> a(nv_in), b(nv_ref);
    for (int i = 0; i < nv_in; i++) a[i] = Pt(in_xyz[3 * i], in_xyz[3 * i + 1], in_xyz[3 * i + 2]);
    for (int i = 0; i < nv_ref; i++) b[i] = Pt(ref_xyz[3 * i], ref_xyz[3 * i + 1], ref_xyz[3 * i + 2]);
    std::vector<double> v(values, values + (size_t)dims * nv_in);
    const std::vector<double> r = smooth_data(a, v, dims, b, sigma, mask);
    std::copy(r.begin(), r.end(), out);
}
void orc_histogram_normalization(double* in, int nv_in, const double* ref, int nv_ref, int dims, const double* excl_in, int rows_in,
                                 const double* excl_ref, int rows_ref) {
    std::vector<double> a(in, in + (size_t)dims * nv_in), b(ref, ref + (size_t)dims * nv_ref);
    histogram_normalization(a, nv_in, b, nv_ref, dims, excl_in, rows_in, excl_ref, rows_ref);
    std::copy(a.begin(), a.end(), in);
}
void orc_variance_normalise(double* data, int dims, int nv, const double* mask) {
    std::vector<double> a(data, data + (size_t)dims * nv);
    variance_normalise(a, dims, nv, mask);
    std::copy(a.begin(), a.end(), data);
}
void orc_create_exclusion(const double* data, int dims, int nv, float lower, float upper, double* out) {
    std::vector<double> a(data, data + (size_t)dims * nv);
    const std::vector<double> m = create_exclusion(a, dims, nv, lower, upper);
    std::copy(m.begin(), m.end(), out);
}
void orc_metric_resample_excl(const double* in_xyz, int nv_in, const int* in_tris, int nf_in, const double* values, int dims, const double* mask,
                              const double* ref_xyz, int nv_ref, const int* ref_tris, int nf_ref, double* out) {
    Mesh A = mesh_from(in_xyz, nv_in, in_tris, nf_in), B = mesh_from(ref_xyz, nv_ref, ref_tris, nf_ref);
    std::vector<double> v(values, values + (size_t)dims * nv_in);
    const std::vector<double> r = metric_resample_excl(A, v, dims, B, mask);
    std::copy(r.begin(), r.end(), out);
}
// featurespace::initialize for the pairwise driver's two meshes (input, reference).  ico > 0: out_* sized [dims][V(ico)];
// ico == 0: every mesh keeps its own vertices.  Returns the template's vertex count, -1 on error.
int orc_featurespace_initialize(int ico, const double* xyz0, int nv0, const int* tris0, int nf0, const double* data0, const double* xyz1, int nv1,
                                const int* tris1, int nf1, const double* data1, int dims, double sigma0, double sigma1, float thr_lower,
                                float thr_upper, int intensitynorm, int varnorm, int cut, int exclude, double* out0, double* out1, double* excl0,
                                double* excl1) {
    try {
        std::vector<FeatureMesh> IN(2);
        IN[0].mesh = mesh_from(xyz0, nv0, tris0, nf0); IN[0].data.assign(data0, data0 + (size_t)dims * nv0); IN[0].D = dims;
        IN[1].mesh = mesh_from(xyz1, nv1, tris1, nf1); IN[1].data.assign(data1, data1 + (size_t)dims * nv1); IN[1].D = dims;
        FeatureOptions o;
        o.sigma = {sigma0, sigma1}; o.thr_lower = thr_lower; o.thr_upper = thr_upper;
        o.intensitynorm = intensitynorm != 0; o.varnorm = varnorm != 0; o.cut = cut != 0; o.exclude = exclude != 0;
        std::vector<std::vector<double>> D, E;
        const Mesh t = featurespace_initialize(ico, IN, o, D, E);
        std::copy(D[0].begin(), D[0].end(), out0); std::copy(D[1].begin(), D[1].end(), out1);
        if (excl0 && !E[0].empty()) std::copy(E[0].begin(), E[0].end(), excl0);
        if (excl1 && !E[1].empty()) std::copy(E[1].begin(), E[1].end(), excl1);
        return t.nvertices();
    } catch (const std::exception&) { return -1; }
}

// ---------------- mesh warp ----------------
void orc_mesh_interpolate(const double* pts, int m, const double* from_xyz, int nv, const int* tris, int nf, const double* to_xyz, double* out) {
    Mesh M = mesh_from(from_xyz, nv, tris, nf);
    std::vector<Pt> up(m), fin(nv);
    for (int i = 0; i < m; i++) up[i] = Pt(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
    for (int i = 0; i < nv; i++) fin[i] = Pt(to_xyz[3 * i], to_xyz[3 * i + 1], to_xyz[3 * i + 2]);
    barycentric_mesh_interpolation(up, M, fin);
    for (int i = 0; i < m; i++) { out[3 * i] = up[i].X; out[3 * i + 1] = up[i].Y; out[3 * i + 2] = up[i].Z; }
}
int orc_unfold(double* xyz, int nv, const int* tris, int nf) {
    Mesh M = mesh_from(xyz, nv, tris, nf);
    int it = unfold(M);
    for (int i = 0; i < nv; i++) { xyz[3 * i] = M.V[i].X; xyz[3 * i + 1] = M.V[i].Y; xyz[3 * i + 2] = M.V[i].Z; }
    return it;
}
int orc_count_folded(const double* xyz, int nv, const int* tris, int nf) {
    Mesh M = mesh_from(xyz, nv, tris, nf);
    int n = 0;
    for (int i = 0; i < nv; i++) n += check_for_intersections(i, M) ? 1 : 0;
    return n;
}

// ---------------- cost function object ----------------
void* orc_cf_create() { return new CostFunction(); }
void orc_cf_destroy(void* h) { delete (CostFunction*)h; }
void orc_cf_set_params(void* h, double lambda, double rexp, double mu, double kappa, double k_exp, double range, int simmeasure,
                       int rmode, int dweight, int anorm, int threads, int bary_orth) {
    Params& P = ((CostFunction*)h)->P;
    P.lambda = lambda; P.rexp = rexp; P.mu = mu; P.kappa = kappa; P.k_exp = k_exp; P.range = range;
    P.simmeasure = simmeasure; P.rmode = rmode; P.dweight = dweight; P.anorm = anorm; P.threads = threads;
    P.bary_orthogonal = bary_orth;
}
void orc_cf_set_meshes(void* h, const double* txyz, int Vt, const int* ttris, int Ft, const double* sxyz, int Vs,
                       const double* cxyz, int N, const int* ctris, int T) {
    CostFunction* cf = (CostFunction*)h;
    // the source is the template icosphere deformed (SPH_orig warped, src/mesh_registration.cpp:66,336): same topology
    Mesh tgt = mesh_from(txyz, Vt, ttris, Ft), src = mesh_from(sxyz, Vs, Vs == Vt ? ttris : nullptr, Vs == Vt ? Ft : 0), cp = mesh_from(cxyz, N, ctris, T);
    cf->set_meshes(tgt, src, cp);
    cf->set_initial_angles();
}
void orc_cf_set_orig(void* h, const double* xyz, int n) {
    CostFunction* cf = (CostFunction*)h;
    cf->ORIG.resize(n);
    for (int i = 0; i < n; i++) cf->ORIG[i] = Pt(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}
void orc_cf_set_features(void* h, int C, const double* ref, int Vt, const double* input, int Vs, int multivariate) {
    CostFunction* cf = (CostFunction*)h;
    cf->C = C; cf->multivariate = multivariate != 0;
    cf->ref.assign(C, {}); cf->input.assign(C, {});
    for (int d = 0; d < C; d++) { cf->ref[d].assign(ref + (size_t)d * Vt, ref + (size_t)(d + 1) * Vt); cf->input[d].assign(input + (size_t)d * Vs, input + (size_t)(d + 1) * Vs); }
}
void orc_cf_set_cfweight(void* h, int rows, const double* w, int Vs) {
    CostFunction* cf = (CostFunction*)h;
    cf->cfweight.assign(rows, {});
    for (int d = 0; d < rows; d++) cf->cfweight[d].assign(w + (size_t)d * Vs, w + (size_t)(d + 1) * Vs);
}
void orc_cf_set_absweights(void* h, const double* aw, int n) { ((CostFunction*)h)->AbsoluteWeights.assign(aw, aw + n); ((CostFunction*)h)->absweights_given = true; }
void orc_cf_get_absweights(void* h, double* out) { CostFunction* cf = (CostFunction*)h; std::copy(cf->AbsoluteWeights.begin(), cf->AbsoluteWeights.end(), out); }
void orc_cf_set_spacings(void* h, const double* maxsep, int n, double MVDmax) {
    CostFunction* cf = (CostFunction*)h;
    cf->MAXSEP.assign(maxsep, maxsep + n); cf->MVDmax = MVDmax;
}
void orc_cf_set_labels(void* h, const double* labels, int L, const double* ROT, int n) {
    CostFunction* cf = (CostFunction*)h;
    cf->labels.resize(L);
    for (int i = 0; i < L; i++) cf->labels[i] = Pt(labels[3 * i], labels[3 * i + 1], labels[3 * i + 2]);
    cf->ROT.resize(n);
    for (int k = 0; k < n; k++) cf->ROT[k] = mat_from(ROT + 9 * k);
}
void orc_cf_reset_source(void* h, const double* xyz) {
    CostFunction* cf = (CostFunction*)h;
    for (int i = 0; i < cf->SOURCE.nvertices(); i++) cf->SOURCE.V[i] = Pt(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}
void orc_cf_reset_cpgrid(void* h, const double* xyz) {
    CostFunction* cf = (CostFunction*)h;
    for (int i = 0; i < cf->CPgrid.nvertices(); i++) cf->CPgrid.V[i] = Pt(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
}
void orc_cf_set_pairs(void* h, const int* p, int np) { ((CostFunction*)h)->pairs.assign(p, p + 2 * np); }
void orc_cf_set_triplets(void* h, const int* t, int nt) { ((CostFunction*)h)->triplets.assign(t, t + 3 * nt); }
void orc_cf_get_source_data(void* h) { ((CostFunction*)h)->get_source_data(); }
void orc_cf_patch_sizes(void* h, int* sizes) {
    CostFunction* cf = (CostFunction*)h;
    for (int k = 0; k < cf->N(); k++) sizes[k] = (int)cf->sourceinrange[k].size();
}
void orc_cf_patch(void* h, int k, int* out) {
    CostFunction* cf = (CostFunction*)h;
    std::memcpy(out, cf->sourceinrange[k].data(), sizeof(int) * cf->sourceinrange[k].size());
}
// sub-range variants let the timed CPU baseline run a bounded sample of the same workload
void orc_cf_get_source_data_range(void* h, int k0, int k1) {
    CostFunction* cf = (CostFunction*)h;
    if ((int)cf->sourceinrange.size() != cf->N()) cf->sourceinrange.assign(cf->N(), {});
    if ((int)cf->AbsoluteWeights.size() != cf->N()) cf->AbsoluteWeights.assign(cf->N(), 1.0);
#pragma omp parallel for num_threads(cf->P.threads) schedule(dynamic, 4)
    for (int k = k0; k < k1; k++) {
        cf->sourceinrange[k].clear();
        for (int i = 0; i < cf->SOURCE.nvertices(); i++)
            if (cf->within_controlpt_range(k, i)) cf->sourceinrange[k].push_back(i);
    }
}
void orc_cf_unary_costs(void* h, double* out) { ((CostFunction*)h)->computeUnaryCosts(out); }
// same loop structure (labels outer, omp over CPs) restricted to CPs [k0,k1); out is [L][k1-k0]
void orc_cf_unary_costs_range(void* h, int k0, int k1, double* out) {
    CostFunction* cf = (CostFunction*)h;
    const int n = k1 - k0;
    for (int j = 0; j < cf->L(); j++) {
#pragma omp parallel for num_threads(cf->P.threads) schedule(dynamic, 16)
        for (int k = k0; k < k1; k++) out[(size_t)j * n + (k - k0)] = cf->computeUnaryCost(k, j);
    }
}
double orc_cf_unary_cost(void* h, int node, int label) { return ((CostFunction*)h)->computeUnaryCost(node, label); }
void orc_cf_triplet_costs(void* h, int n, const int* q, double* out) {
    CostFunction* cf = (CostFunction*)h;
#pragma omp parallel for num_threads(cf->P.threads) schedule(static)
    for (int i = 0; i < n; i++) out[i] = cf->computeTripletCost(q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]);
}
// full table [t][la][lb][lc] for triplets [t0,t1)
void orc_cf_triplet_table(void* h, int t0, int t1, double* out) {
    CostFunction* cf = (CostFunction*)h;
    const int L = cf->L();
#pragma omp parallel for num_threads(cf->P.threads) schedule(dynamic, 1)
    for (int t = t0; t < t1; t++)
        for (int a = 0; a < L; a++)
            for (int b = 0; b < L; b++)
                for (int c = 0; c < L; c++) out[(((size_t)(t - t0) * L + a) * L + b) * L + c] = cf->computeTripletCost(t, a, b, c);
}
void orc_cf_pair_table(void* h, double* out) { ((CostFunction*)h)->computePairwiseCosts(out); }
double orc_cf_pair_cost(void* h, int p, int la, int lb) { return ((CostFunction*)h)->computePairwiseCost(p, la, lb); }
double orc_cf_total(void* h, const int* labeling, int use_unary) { return ((CostFunction*)h)->evaluateTotalCostSum(labeling, use_unary != 0); }
double orc_cf_meanangle(void* h) { return ((CostFunction*)h)->MEANANGLE; }

// ---------------- group cost ----------------
void* orc_gc_create() { return new GroupCost(); }
void orc_gc_destroy(void* h) { delete (GroupCost*)h; }
void orc_gc_set(void* h, double lambda, double rexp, double mu, double kappa, int S, const double* cp_xyz /*S*N*3*/, int N,
                const int* cp_tris, int T, const double* orig_xyz /*N*3*/, const double* labels, int L, const double* ROT /*S*N*9*/,
                const int* triplets /*S*T*3 global*/) {
    GroupCost* g = (GroupCost*)h;
    g->P.lambda = lambda; g->P.rexp = rexp; g->P.mu = mu; g->P.kappa = kappa;
    g->VERTICES_PER_SUBJ = N; g->TRIPLETS_PER_SUBJ = T;
    g->CONTROLMESHES.clear();
    for (int s = 0; s < S; s++) g->CONTROLMESHES.push_back(mesh_from(cp_xyz + (size_t)s * N * 3, N, cp_tris, T));
    g->ORIG.resize(N);
    for (int i = 0; i < N; i++) g->ORIG[i] = Pt(orig_xyz[3 * i], orig_xyz[3 * i + 1], orig_xyz[3 * i + 2]);
    g->labels.resize(L);
    for (int i = 0; i < L; i++) g->labels[i] = Pt(labels[3 * i], labels[3 * i + 1], labels[3 * i + 2]);
    g->ROT.resize((size_t)S * N);
    for (size_t k = 0; k < (size_t)S * N; k++) g->ROT[k] = mat_from(ROT + 9 * k);
    g->triplets.assign(triplets, triplets + (size_t)S * T * 3);
}
void orc_gc_triplet_costs(void* h, int n, const int* q, double* out) {
    GroupCost* g = (GroupCost*)h;
    for (int i = 0; i < n; i++) out[i] = g->computeTripletCost(q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]);
}
// patch maps as CSR: entry e of map m is (keys[off[m]+e], vals[off[m]+e]); m = (subject*N + node)*L + label
void orc_gc_set_patches(void* h, int nmaps, const long long* off, const int* keys, const double* vals, const int* pairs, int np) {
    GroupCost* g = (GroupCost*)h;
    g->patch_data.assign(nmaps, {});
    for (int m = 0; m < nmaps; m++)
        for (long long e = off[m]; e < off[m + 1]; e++) g->patch_data[m][keys[e]] = vals[e];
    g->pairs.assign(pairs, pairs + 2 * (size_t)np);
}
void orc_gc_pair_costs(void* h, int n, const int* q /*n*3: pair,la,lb*/, int L, double* out) {
    GroupCost* g = (GroupCost*)h;
    for (int i = 0; i < n; i++) out[i] = g->computePairwiseCost(q[3 * i], q[3 * i + 1], q[3 * i + 2], L);
}

}  // extern "C"
