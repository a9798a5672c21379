// harness.cpp — flat C entry points over the host-side mirror classes (src/DiscreteModel.h, src/DiscreteCostFunction.h)
// so that Python (tests, bench.py) can drive the SAME call sequence the reference's driver makes
// (src/mesh_registration.cpp:316-397: reset_meshspace -> setupCostFunction -> computeUnaryCosts -> optimise ->
// applyLabeling).  Compiled twice:
//   * libmsm_host.so  (product): model + GPU-backed cost functions only;
//   * oracle/_ref/libref_optim.so (test infrastructure, -DWITH_REF_OPTIM): additionally the UNMODIFIED vendored optimisers
//     from /root/reference/include (FastPD, ELC/HOCR, Fusion), compiled where they lie against these FSL-free headers;
//   * oracle/_ref/libref_cost.so (test infrastructure, -DWITH_REF_COST -DWITH_REF_OPTIM): the SAME entry points over the
//     REFERENCE'S OWN DiscreteModel / DiscreteCostFunction classes (src/*.cpp compiled where they lie, oracle/Makefile
//     `refcost`) — the CPU reference the GPU path and the oracle port are pinned against.
#include <cstring>
#include <memory>
#include <random>

#ifdef WITH_REF_COST
#include "DiscreteGroupModel.h"  // /root/reference/src (the product mirror is NOT on the include path of this build)
#include "DiscreteModel.h"
#else
#include "src/DiscreteGroupModel.h"
#include "src/DiscreteModel.h"
#endif

#ifdef WITH_REF_OPTIM
#define HAS_FPD
#define HAS_HOCR
#include "FastPD/FastPD.h"
#include "Fusion/Fusion.h"
#ifndef WITH_REF_COST
#include "src/FusionTables.h"
#endif
#endif

#ifdef WITH_REF_DRIVER_LOOP
// The reference's driver loop and MCMC optimiser, textually (oracle/Makefile cuts Mesh_registration::run_discrete_opt out of
// src/mesh_registration.cpp into oracle/_ref/gen/ at build time; src/mcmc_opt.h is included where it lies).  In the product
// build the reference header's include guard is pre-defined, so that its `#include "DiscreteModel.h"` (which would find the
// reference's own file next to it) is a no-op and MCMC::optimise compiles against the mirror classes declared above.
#ifndef WITH_REF_COST
#define NEWMESHREG_DISCRETEMODEL_H
#endif
#include REF_MCMC_H
namespace newmeshreg {
// Exactly the members Mesh_registration::run_discrete_opt reads (src/mesh_registration.h:52-57,88-91,107-108,111-112,128,
// 135-136,146,155), with their types; everything else of the driver class is option parsing and file I/O.
class Mesh_registration {
public:
    int level = 1;
    std::shared_ptr<NonLinearSRegDiscreteModel> model;
    NEWMAT::Matrix SPHin_CFWEIGHTING, SPHref_CFWEIGHTING;
    std::string _discreteOPT;
    myparam PARAMETERS;
    bool _tricliquelikeihood = false, _incfw = false, _refcfw = false;
    std::vector<int> _mciters;
    int _numthreads = 1;
    bool _verbose = false;
    NEWMAT::Matrix combine_costfunction_weighting(const NEWMAT::Matrix&, const NEWMAT::Matrix&);
    void run_discrete_opt(newresampler::Mesh&);
};
#include "_ref/gen/run_discrete_opt.inc"
}  // namespace newmeshreg
#endif

using namespace newmeshreg;

namespace {
thread_local std::string g_err;

// The few places where this file needs more than the reference's class surface offers (raw-array meshes, in-memory
// features, reading the label set back) go through the adapters below, one set per build.
#ifdef WITH_REF_COST
newresampler::Mesh mesh_from(const double* xyz, int nv, const int* tris, int nf) {
    orc::Mesh m;
    m.V.resize(nv); m.F.resize(nf);
    for (int i = 0; i < nv; i++) m.V[i] = orc::Pt(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    for (int t = 0; t < nf; t++) m.F[t] = {tris[3 * t], tris[3 * t + 1], tris[3 * t + 2]};
    m.build_adjacency();
    return newresampler::Mesh(m);
}
std::vector<double> coords_of(const newresampler::Mesh& m) {
    std::vector<double> o((size_t)m.nvertices() * 3);
    for (int i = 0; i < m.nvertices(); i++) { const newresampler::Point& p = m.get_coord(i); o[3 * i] = p.X; o[3 * i + 1] = p.Y; o[3 * i + 2] = p.Z; }
    return o;
}
std::vector<int> faces_of(const newresampler::Mesh& m) {
    std::vector<int> o((size_t)m.ntriangles() * 3);
    for (int t = 0; t < m.ntriangles(); t++) for (int j = 0; j < 3; j++) o[3 * t + j] = m.get_triangle_vertexID(t, j);
    return o;
}
void assign_coords(newresampler::Mesh& m, const double* xyz) {
    for (int i = 0; i < m.nvertices(); i++) m.set_coord(i, newresampler::Point(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
}
// The reference's featurespace only has file-name constructors (src/featurespace.h:14-15; featurespace.cc needs FSL
// I/O and is not compiled): oracle/ref_cost_ext.cpp defines them to take their matrices from this registry instead.
}  // namespace
namespace refcost { std::map<std::string, std::shared_ptr<MISCMATHS::BFMatrix>>& feature_registry(); }
namespace {
std::shared_ptr<featurespace> make_features(const double* input, int Vin, const double* ref, int Vref, int dim) {
    auto put = [&](const char* key, const double* d, int V) {
        auto M = std::make_shared<MISCMATHS::FullBFMatrix>(dim, V);
        for (int c = 0; c < dim; c++) for (int v = 0; v < V; v++) M->Set(c + 1, v + 1, d[(size_t)c * V + v]);
        refcost::feature_registry()[key] = M;
    };
    put("harness:input", input, Vin);
    put("harness:reference", ref, Vref);
    return std::make_shared<featurespace>(std::string("harness:input"), std::string("harness:reference"));
}
struct CostFunctionPeek : NonLinearSRegDiscreteCostFunction { using NonLinearSRegDiscreteCostFunction::_sourceinrange; };
struct ModelT : NonLinearSRegDiscreteModel {
    using NonLinearSRegDiscreteModel::NonLinearSRegDiscreteModel;
    const std::vector<newresampler::Point>& get_labels() const { return m_labels; }
    // the per-control-point source patches built by get_source_data (src/DiscreteCostFunction.cpp:394-410,458-481)
    const std::vector<std::vector<int>>& patches() const { return static_cast<const CostFunctionPeek*>(costfct.get())->_sourceinrange; }
};
myparam default_parameters() {  // one resolution level (src/mesh_registration.cpp:491-556,689-715,815-847)
    myparam P;
    P["lambda"] = 0.1f;  P["exponent"] = 2.0f;  P["weight"] = false;  P["anorm"] = false;
    P["shearmodulus"] = 0.4f;  P["bulkmodulus"] = 1.6f;  P["kexponent"] = 2.0f;  P["range"] = 1.0f;
    P["simmeasure"] = 2;  P["verbosity"] = false;  P["regularisermode"] = 1;  P["numthreads"] = 1;
    P["dOPT"] = std::string("FastPD");  P["SGres"] = 4;  P["CPres"] = 2;  P["multivariate"] = false;
    P["outdir"] = std::string("");  P["TriLikelihood"] = false;  P["rescalelabels"] = false;  P["labeldist"] = 0.5f;
    P["iters"] = 3;
    return P;
}
#else
using ModelT = NonLinearSRegDiscreteModel;
newresampler::Mesh mesh_from(const double* xyz, int nv, const int* tris, int nf) {
    newresampler::Mesh m;
    m.set(xyz, nv, tris, nf);
    return m;
}
std::vector<double> coords_of(const newresampler::Mesh& m) { return m.coords(); }
std::vector<int> faces_of(const newresampler::Mesh& m) { return m.faces(); }
void assign_coords(newresampler::Mesh& m, const double* xyz) { m.set_coords(xyz); }   // one bulk copy
std::shared_ptr<featurespace> make_features(const double* input, int Vin, const double* ref, int Vref, int dim) {
    return std::make_shared<featurespace>(input, Vin, ref, Vref, dim);
}
#endif

struct ModelBox {
    std::shared_ptr<ModelT> model;
    myparam params;
    newresampler::Mesh target, source;
    NEWMAT::Matrix cfweight;   // optional cost-function weighting (rows x source vertices); empty = the driver's unit weights
};
template <typename F>
int guard(F&& f) {
    try { f(); return 0; }
    catch (const MeshregException& e) { g_err = e.errmesg ? e.errmesg : "MeshregException"; return 1; }
    catch (const std::exception& e) { g_err = e.what(); return 1; }
}

// A model whose costs are explicit tables: lets the optimisers run from oracle- or GPU-produced tables alike.
class TableCostFunction : public DiscreteCostFunction {
public:
    TableCostFunction(int N, int L, int P, int T, const double* unary, const double* pairtab, const double* triptab)
        : U(unary, unary + (size_t)N * L), PT(pairtab ? pairtab : nullptr, pairtab ? pairtab + (size_t)P * L * L : nullptr),
          TT(triptab ? triptab : nullptr, triptab ? triptab + (size_t)T * L * L * L : nullptr) {
        initialize(N, L, P, T);
        std::copy(U.begin(), U.end(), unarycosts);
    }
    void computeUnaryCosts() override { std::copy(U.begin(), U.end(), unarycosts); }
    double computeUnaryCost(int node, int label) override { return U[(size_t)label * m_num_nodes + node]; }
    double computePairwiseCost(int pair, int la, int lb) override { return PT.empty() ? 0.0 : PT[((size_t)pair * m_num_labels + lb) * m_num_labels + la]; }
    double computeTripletCost(int t, int a, int b, int c) override {
        const size_t L = m_num_labels;
        return TT.empty() ? 0.0 : TT[(((size_t)t * L + a) * L + b) * L + c];
    }
private:
    std::vector<double> U, PT, TT;
};
class TableModel : public DiscreteModel {
public:
    TableModel(int N, int L, int P, const int* prs, int T, const int* trs, std::shared_ptr<TableCostFunction> cf) : costfct(cf) {
        m_num_nodes = N; m_num_labels = L; m_num_pairs = P; m_num_triplets = T;
        initLabeling();
        pairs = new int[2 * (size_t)std::max(P, 1)];
        triplets = new int[3 * (size_t)std::max(T, 1)];
        if (P) std::copy(prs, prs + 2 * (size_t)P, pairs);
        if (T) std::copy(trs, trs + 3 * (size_t)T, triplets);
    }
    std::shared_ptr<DiscreteCostFunction> getCostFunction() override { return costfct; }
    void computeUnaryCosts() override { costfct->computeUnaryCosts(); }
    double computeUnaryCost(int n, int l) override { return costfct->computeUnaryCost(n, l); }
    double computePairwiseCost(int p, int a, int b) override { return costfct->computePairwiseCost(p, a, b); }
    double computeTripletCost(int t, int a, int b, int c) override { return costfct->computeTripletCost(t, a, b, c); }
    double evaluateTotalCostSum() override { return costfct->evaluateTotalCostSum(labeling, pairs, triplets); }
private:
    std::shared_ptr<TableCostFunction> costfct;
};
}  // namespace

extern "C" {

const char* msmhost_last_error() { return g_err.c_str(); }

// newresampler::make_mesh_from_icosa + true_rescale (src/mesh_registration.cpp:67-69)
void msmhost_icosphere_counts(int level, int* nv, int* nf) {
    long f = 20; for (int i = 0; i < level; i++) f *= 4;
    *nf = (int)f; *nv = (int)(f / 2 + 2);
}
void msmhost_icosphere(int level, double radius, double* xyz, int* tris) {
    newresampler::Mesh m = newresampler::make_mesh_from_icosa(level);
    newresampler::true_rescale(m, radius);
    auto c = coords_of(m); auto f = faces_of(m);
    std::memcpy(xyz, c.data(), c.size() * sizeof(double));
    std::memcpy(tris, f.data(), f.size() * sizeof(int));
}

// ---- NonLinearSRegDiscreteModel ----------------------------------------------------------------------------------
void* msmhost_model_create(double lambda, double rexp, double mu, double kappa, double kexp, double range, int simmeasure, int rmode,
                           int dweight, int anorm, int sgres, int cpres, double labeldist, int multivariate, const char* dopt, int threads) {
    auto* b = new ModelBox();
    b->params = default_parameters();
    myparam& P = b->params;
    P["lambda"] = (float)lambda; P["exponent"] = (float)rexp; P["shearmodulus"] = (float)mu; P["bulkmodulus"] = (float)kappa;
    P["kexponent"] = (float)kexp; P["range"] = (float)range; P["simmeasure"] = simmeasure; P["regularisermode"] = rmode;
    P["weight"] = dweight != 0; P["anorm"] = anorm != 0; P["SGres"] = sgres; P["CPres"] = cpres; P["labeldist"] = (float)labeldist;
    P["multivariate"] = multivariate != 0; P["dOPT"] = std::string(dopt); P["numthreads"] = threads;
    if (guard([&] { b->model = std::make_shared<ModelT>(P); })) { delete b; return nullptr; }
    return b;
}
void msmhost_model_destroy(void* h) { delete (ModelBox*)h; }

int msmhost_model_set_data(void* h, const double* txyz, int Vt, const int* ttris, int Ft, const double* sxyz, int Vs, const int* stris, int Fs,
                           const double* input, const double* ref, int C) {
    auto* b = (ModelBox*)h;
    return guard([&] {
        b->target = mesh_from(txyz, Vt, ttris, Ft);
        b->source = mesh_from(sxyz, Vs, stris, Fs);
        b->model->set_featurespace(make_features(input, Vs, ref, Vt, C));
        b->model->set_meshspace(b->target, b->source);
    });
}
// CONTROL = ico(CPres) at RAD (src/mesh_registration.cpp:67-69) then model->Initialize(CONTROL) (:86)
int msmhost_model_initialize(void* h, int cpres) {
    auto* b = (ModelBox*)h;
    return guard([&] {
        newresampler::Mesh CONTROL = newresampler::make_mesh_from_icosa(cpres);
        newresampler::recentre(CONTROL);
        newresampler::true_rescale(CONTROL, RAD);
        b->model->Initialize(CONTROL);
    });
}
int msmhost_model_sizes(void* h, int* N, int* L, int* P, int* T) {
    auto* b = (ModelBox*)h;
    *N = b->model->getNumNodes(); *L = b->model->getNumLabels(); *P = b->model->getNumPairs(); *T = b->model->getNumTriplets();
    return 0;
}
int msmhost_model_reset_meshspace(void* h, const double* sxyz) {  // src/mesh_registration.cpp:336
    auto* b = (ModelBox*)h;
    return guard([&] {
        assign_coords(b->source, sxyz);
        b->model->reset_meshspace(b->source);
    });
}
int msmhost_model_reset_cpgrid(void* h, const double* cxyz) {  // :393
    auto* b = (ModelBox*)h;
    return guard([&] {
        newresampler::Mesh g = b->model->get_CPgrid();
        assign_coords(g, cxyz);
        b->model->reset_CPgrid(g);
    });
}
// cost-function weighting handed to setupCostFunctionWeighting (src/mesh_registration.cpp:323-334): w[rows][source vertices]
int msmhost_model_set_weights(void* h, const double* w, int rows, int V) {
    auto* b = (ModelBox*)h;
    return guard([&] {
        b->cfweight.ReSize(rows, V);
        for (int r = 0; r < rows; r++) for (int v = 0; v < V; v++) b->cfweight(r + 1, v + 1) = w[(size_t)r * V + v];
    });
}
int msmhost_model_setup(void* h) {  // :330-337
    auto* b = (ModelBox*)h;
    return guard([&] {
        if (b->cfweight.Ncols() > 0) {
            b->model->setupCostFunctionWeighting(b->cfweight);
            b->model->setupCostFunction();
            return;
        }
#ifdef WITH_REF_COST
        // the driver hands the cost function unit weights when no weighting is supplied (:330-334); the reference's
        // own initialize() would otherwise ReSize them to zeros (src/DiscreteCostFunction.cpp:115-116).  The product
        // mirror defaults to unit weights by itself (host/src/DiscreteCostFunction.h initialize()).
        NEWMAT::Matrix CombinedWeight;
        CombinedWeight.resize(1, b->source.nvertices());
        CombinedWeight = 1;
        b->model->setupCostFunctionWeighting(CombinedWeight);
#endif
        b->model->setupCostFunction();
    });
}
int msmhost_model_unary(void* h, double* out) {  // :347 computeUnaryCosts + the table
    auto* b = (ModelBox*)h;
    return guard([&] {
        b->model->computeUnaryCosts();
        const size_t n = (size_t)b->model->getNumNodes() * b->model->getNumLabels();
        std::memcpy(out, b->model->getCostFunction()->getUnaryCosts(), n * sizeof(double));
    });
}
int msmhost_model_get(void* h, double* cpgrid, double* labels, int* pairs, int* triplets) {
    auto* b = (ModelBox*)h;
    return guard([&] {
        if (cpgrid) { auto c = coords_of(b->model->get_CPgrid()); std::memcpy(cpgrid, c.data(), c.size() * sizeof(double)); }
        if (labels) {
            const auto& l = b->model->get_labels();
            for (size_t i = 0; i < l.size(); i++) { labels[3 * i] = l[i].X; labels[3 * i + 1] = l[i].Y; labels[3 * i + 2] = l[i].Z; }
        }
        if (pairs && b->model->getNumPairs()) std::memcpy(pairs, b->model->getPairs(), sizeof(int) * 2 * b->model->getNumPairs());
        if (triplets && b->model->getNumTriplets()) std::memcpy(triplets, b->model->getTriplets(), sizeof(int) * 3 * b->model->getNumTriplets());
    });
}
// per-element virtuals, batched: kind 0 unary (q[n][2]), 1 pairwise (q[n][3]), 2 triplet (q[n][4])
int msmhost_model_costs(void* h, int kind, int n, const int* q, double* out) {
    auto* b = (ModelBox*)h;
    return guard([&] {
#ifdef WITH_REF_COST
        // only the reference's triplet virtual is re-entrant: computeUnaryCost keeps patch means in the shared `sim`
        // member (src/similarities.h:38) and computePairwiseCost moves _CPgrid vertices and puts them back (:269-292)
        const bool parallel = kind == 2;
#else
        const bool parallel = true;
#endif
#pragma omp parallel for if (parallel)
        for (int i = 0; i < n; i++) {
            if (kind == 0) out[i] = b->model->computeUnaryCost(q[2 * i], q[2 * i + 1]);
            else if (kind == 1) out[i] = b->model->computePairwiseCost(q[3 * i], q[3 * i + 1], q[3 * i + 2]);
            else out[i] = b->model->computeTripletCost(q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]);
        }
    });
}
int msmhost_model_set_labeling(void* h, const int* lab) {
    auto* b = (ModelBox*)h;
    std::memcpy(b->model->getLabeling(), lab, sizeof(int) * b->model->getNumNodes());
    return 0;
}
int msmhost_model_get_labeling(void* h, int* lab) {
    auto* b = (ModelBox*)h;
    std::memcpy(lab, b->model->getLabeling(), sizeof(int) * b->model->getNumNodes());
    return 0;
}
double msmhost_model_total(void* h) { auto* b = (ModelBox*)h; double v = 0; guard([&] { v = b->model->evaluateTotalCostSum(); }); return v; }
int msmhost_model_apply_labeling(void* h, double* new_cpgrid) {  // :386
    auto* b = (ModelBox*)h;
    return guard([&] {
        b->model->applyLabeling();
        auto c = coords_of(b->model->get_CPgrid());
        std::memcpy(new_cpgrid, c.data(), c.size() * sizeof(double));
    });
}
#ifndef WITH_REF_COST
// borrowed handle of the cost function's device context (device-resident table variants of the C ABI)
void* msmhost_model_context(void* h) { return ((ModelBox*)h)->model->getNonLinearCostFunction()->context(); }
int msmhost_model_triplet_moves(void* h, const int* labeling, int alpha, double* out) {
    auto* b = (ModelBox*)h;
    return guard([&] { b->model->getNonLinearCostFunction()->computeTripletMoves(labeling, alpha, out); });
}
int msmhost_model_triplet_moves_f32(void* h, const int* labeling, int alpha, float* out) {
    auto* b = (ModelBox*)h;
    return guard([&] { b->model->getNonLinearCostFunction()->computeTripletMoves(labeling, alpha, out); });
}
#else
// the reference has no batched move evaluation: the same [L][T][8] / [T][8] layout from its per-element virtual
// (include/Fusion/Fusion.h:187-194 bit order: bit2 -> A, bit1 -> B, bit0 -> C take the proposal label)
int msmhost_model_triplet_moves(void* h, const int* labeling, int alpha, double* out) {
    auto* b = (ModelBox*)h;
    return guard([&] {
        const int T = b->model->getNumTriplets(), L = b->model->getNumLabels();
        const int* tr = b->model->getTriplets();
        const int a0 = alpha < 0 ? 0 : alpha, a1 = alpha < 0 ? L : alpha + 1;
#pragma omp parallel for collapse(2)
        for (int a = a0; a < a1; a++)
            for (int t = 0; t < T; t++) {
                const int cur[3] = {labeling[tr[3 * t]], labeling[tr[3 * t + 1]], labeling[tr[3 * t + 2]]};
                double* o = out + ((size_t)(a - a0) * T + t) * 8;
                for (int m = 0; m < 8; m++)
                    o[m] = b->model->computeTripletCost(t, (m & 4) ? a : cur[0], (m & 2) ? a : cur[1], (m & 1) ? a : cur[2]);
            }
    });
}
#endif

// ---- mesh warp -----------------------------------------------------------------------------------
// barycentric_mesh_interpolation(SPH_up, SPH_low_init, SPH_low_final) on raw arrays; `up` is updated in place
int msmhost_mesh_interpolate(double* up, int m, const double* init_xyz, int nv, const int* tris, int nf, const double* final_xyz) {
    return guard([&] {
        newresampler::Mesh U = mesh_from(up, m, nullptr, 0);
        newresampler::Mesh A = mesh_from(init_xyz, nv, tris, nf);
        newresampler::Mesh B = mesh_from(final_xyz, nv, tris, nf);
        newresampler::barycentric_mesh_interpolation(U, A, B);
        auto c = coords_of(U);
        std::memcpy(up, c.data(), c.size() * sizeof(double));
    });
}
int msmhost_unfold(double* xyz, int nv, const int* tris, int nf) {
    return guard([&] {
        newresampler::Mesh M = mesh_from(xyz, nv, tris, nf);
        unfold(M, false);
        auto c = coords_of(M);
        std::memcpy(xyz, c.data(), c.size() * sizeof(double));
    });
}

// ---- DiscreteGroupModel (src/DiscreteGroupModel.{h,cpp}) behind the same kind of flat interface: the call sequence of
// Group_Mesh_registration (src/group_mesh_registration.cpp:20-36,49-108): set_featurespace / set_meshspace(template, SPH_orig, S) /
// Initialize(control) once, then per outer iteration reset_meshspace(subject mesh) ... setupCostFunction() -> optimise ->
// applyLabeling().  Compiled for the product mirror (GPU-backed) and, with -DWITH_REF_COST, for the reference's own classes.
namespace {
#ifdef WITH_REF_COST
struct GroupModelT : DiscreteGroupModel {
    using DiscreteGroupModel::DiscreteGroupModel;
    const std::vector<newresampler::Point>& get_labels() const { return m_labels; }
};
std::shared_ptr<featurespace> make_group_features(const double* data, int S, int V) {
    std::vector<std::string> keys;
    for (int s = 0; s < S; s++) {
        auto M = std::make_shared<MISCMATHS::FullBFMatrix>(1, V);
        for (int v = 0; v < V; v++) M->Set(1, v + 1, data[(size_t)s * V + v]);
        keys.push_back("harness:subject-" + std::to_string(s));
        refcost::feature_registry()[keys.back()] = M;
    }
    return std::make_shared<featurespace>(keys);
}
#else
using GroupModelT = DiscreteGroupModel;
std::shared_ptr<featurespace> make_group_features(const double* data, int S, int V) {
    auto f = std::make_shared<featurespace>();
    for (int s = 0; s < S; s++) f->add_subject(data + (size_t)s * V, V);
    return f;
}
#endif
struct GroupModelBox {
    std::shared_ptr<GroupModelT> model;
    myparam params;
    newresampler::Mesh templ;
    int S = 0;
};
}  // namespace

void* msmhost_group_create(double lambda, double rexp, double mu, double kappa, double range, int sgres, int cpres, double labeldist, int threads) {
    auto* b = new GroupModelBox();
    b->params = default_parameters();
    myparam& P = b->params;
    P["lambda"] = (float)lambda; P["exponent"] = (float)rexp; P["shearmodulus"] = (float)mu; P["bulkmodulus"] = (float)kappa;
    P["range"] = (float)range; P["regularisermode"] = 3; P["SGres"] = sgres; P["CPres"] = cpres; P["labeldist"] = (float)labeldist;
    P["multivariate"] = false; P["dOPT"] = std::string("HOCR"); P["numthreads"] = threads;
    if (guard([&] { b->model = std::make_shared<GroupModelT>(P); })) { delete b; return nullptr; }
    return b;
}
void msmhost_group_destroy(void* h) { delete (GroupModelBox*)h; }
// template mesh (a regular icosphere), S subjects, per-subject data on that icosphere's vertices: data[s][v]
int msmhost_group_set_data(void* h, const double* txyz, int Vt, const int* ttris, int Ft, int S, const double* data) {
    auto* b = (GroupModelBox*)h;
    return guard([&] {
        b->templ = mesh_from(txyz, Vt, ttris, Ft);
        b->S = S;
        b->model->set_featurespace(make_group_features(data, S, Vt));
        b->model->set_meshspace(b->templ, b->templ, S);
    });
}
int msmhost_group_initialize(void* h, int cpres) {
    auto* b = (GroupModelBox*)h;
    return guard([&] {
        newresampler::Mesh CONTROL = newresampler::make_mesh_from_icosa(cpres);
        newresampler::recentre(CONTROL);
        newresampler::true_rescale(CONTROL, RAD);
        b->model->Initialize(CONTROL);
    });
}
int msmhost_group_sizes(void* h, int* N, int* L, int* P, int* T) {
    auto* b = (GroupModelBox*)h;
    *N = b->model->getNumNodes(); *L = b->model->getNumLabels(); *P = b->model->getNumPairs(); *T = b->model->getNumTriplets();
    return 0;
}
int msmhost_group_reset_meshspace(void* h, int subject, const double* xyz) {
    auto* b = (GroupModelBox*)h;
    return guard([&] {
        newresampler::Mesh m = b->templ;
        assign_coords(m, xyz);
        b->model->reset_meshspace(m, subject);
    });
}
int msmhost_group_reset_cpgrid(void* h, int subject, const double* xyz) {
    auto* b = (GroupModelBox*)h;
    return guard([&] {
        newresampler::Mesh g = b->model->get_CPgrid(subject);
        assign_coords(g, xyz);
        b->model->reset_CPgrid(g, subject);
    });
}
int msmhost_group_setup(void* h) { auto* b = (GroupModelBox*)h; return guard([&] { b->model->setupCostFunction(); }); }
int msmhost_group_get(void* h, double* cpgrids /*[S][N][3]*/, double* labels, int* pairs, int* triplets) {
    auto* b = (GroupModelBox*)h;
    return guard([&] {
        if (cpgrids) {
            size_t o = 0;
            for (int s = 0; s < b->S; s++) { auto c = coords_of(b->model->get_CPgrid(s)); std::memcpy(cpgrids + o, c.data(), c.size() * sizeof(double)); o += c.size(); }
        }
        if (labels) {
            const auto& l = b->model->get_labels();
            for (size_t i = 0; i < l.size(); i++) { labels[3 * i] = l[i].X; labels[3 * i + 1] = l[i].Y; labels[3 * i + 2] = l[i].Z; }
        }
        if (pairs && b->model->getNumPairs()) std::memcpy(pairs, b->model->getPairs(), sizeof(int) * 2 * b->model->getNumPairs());
        if (triplets && b->model->getNumTriplets()) std::memcpy(triplets, b->model->getTriplets(), sizeof(int) * 3 * b->model->getNumTriplets());
    });
}
// per-element virtuals, batched: kind 1 pairwise (q[n][3]), 2 triplet (q[n][4])
int msmhost_group_costs(void* h, int kind, int n, const int* q, double* out) {
    auto* b = (GroupModelBox*)h;
    return guard([&] {
        for (int i = 0; i < n; i++) {
            if (kind == 1) out[i] = b->model->computePairwiseCost(q[3 * i], q[3 * i + 1], q[3 * i + 2]);
            else out[i] = b->model->computeTripletCost(q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]);
        }
    });
}
int msmhost_group_set_labeling(void* h, const int* lab) {
    auto* b = (GroupModelBox*)h;
    std::memcpy(b->model->getLabeling(), lab, sizeof(int) * b->model->getNumNodes());
    return 0;
}
int msmhost_group_get_labeling(void* h, int* lab) {
    auto* b = (GroupModelBox*)h;
    std::memcpy(lab, b->model->getLabeling(), sizeof(int) * b->model->getNumNodes());
    return 0;
}
double msmhost_group_total(void* h) { auto* b = (GroupModelBox*)h; double v = 0; guard([&] { v = b->model->evaluateTotalCostSum(); }); return v; }
int msmhost_group_apply_labeling(void* h, double* cpgrids) {
    auto* b = (GroupModelBox*)h;
    return guard([&] {
        b->model->applyLabeling();
        size_t o = 0;
        for (int s = 0; s < b->S; s++) { auto c = coords_of(b->model->get_CPgrid(s)); std::memcpy(cpgrids + o, c.data(), c.size() * sizeof(double)); o += c.size(); }
    });
}
#ifndef WITH_REF_COST
// featurespace(datain, dataref) + setters + initialize(ico, IN, exclude) (src/featurespace.cc:18-65) on raw arrays; same
// signature as oracle/ref_feat_driver.cpp drives the reference's own class with.  Returns the template's vertex count, -1 on error.
int msmhost_featurespace_initialize(int ico, const double* xyz0, int nv0, const int* tris0, int nf0, const double* data0, const double* xyz1, int nv1,
                                    const int* tris1, int nf1, const double* data1, int dims, double sigma0, double sigma1, float thr_lower,
                                    float thr_upper, int intensitynorm, int varnorm, int cut, int exclude, int threads, double* out0,
                                    double* out1, double* excl0, double* excl1) {
    int nv = -1;
    const int rc = guard([&] {
        std::vector<newresampler::Mesh> IN = {mesh_from(xyz0, nv0, tris0, nf0), mesh_from(xyz1, nv1, tris1, nf1)};
        featurespace fs(data0, nv0, data1, nv1, dims);
        fs.set_smoothing_parameters({sigma0, sigma1});
        fs.set_cutthreshold({thr_lower, thr_upper});
        fs.varnorm(varnorm != 0);
        fs.intensitynormalize(intensitynorm != 0, cut != 0);
        fs.set_nthreads(threads);
        const newresampler::Mesh t = fs.initialize(ico, IN, exclude != 0);
        std::memcpy(out0, fs.input_data(), (size_t)dims * fs.input_vertices() * sizeof(double));
        std::memcpy(out1, fs.ref_data(), (size_t)dims * fs.ref_vertices() * sizeof(double));
        if (excl0 && fs.get_input_excl()) std::memcpy(excl0, fs.get_input_excl()->data(), fs.get_input_excl()->size() * sizeof(double));
        if (excl1 && fs.get_reference_excl()) std::memcpy(excl1, fs.get_reference_excl()->data(), fs.get_reference_excl()->size() * sizeof(double));
        nv = t.nvertices();
    });
    return rc == 0 ? nv : -1;
}
#endif
#ifndef WITH_REF_COST
// rotated_meshes[subject * L + label] (src/DiscreteGroupModel.cpp:92), first data row, on the template's vertices
int msmhost_group_rotated(void* h, int subject, int label, double* out) {
    auto* b = (GroupModelBox*)h;
    return guard([&] { const auto& v = b->model->rotated_mesh_values(subject, label); std::memcpy(out, v.data(), v.size() * sizeof(double)); });
}
// newresampler::metric_resample on raw arrays ([EXT] A8): values[d][v_in] -> out[d][v_ref]
int msmhost_metric_resample(const double* in_xyz, int nv_in, const int* in_tris, int nf_in, const double* values, int dims, const double* ref_xyz,
                            int nv_ref, const int* ref_tris, int nf_ref, double* out) {
    return guard([&] {
        const newresampler::Mesh A = mesh_from(in_xyz, nv_in, in_tris, nf_in), B = mesh_from(ref_xyz, nv_ref, ref_tris, nf_ref);
        const std::vector<double> r = newresampler::metric_resample(A, values, dims, B);
        std::memcpy(out, r.data(), r.size() * sizeof(double));
    });
}
#endif
#ifndef WITH_REF_COST
// the 4 fusion-move costs of every inter-subject pair for proposal label alpha (include/Fusion/Fusion.h:169-172): out[P][4]
int msmhost_group_pair_moves(void* h, const int* labeling, int alpha, double* out) {
    auto* b = (GroupModelBox*)h;
    return guard([&] { b->model->getGroupCostFunction()->computePairwiseMoves(labeling, alpha, out); });
}
// multi-GPU: this rank produces the rotated / resampled data of subjects [s0, s1) only (the S L resamplings are the expensive
// part of setupCostFunction); the caller exchanges them (set / get below) before the patch maps are assembled
int msmhost_group_set_subject_shard(void* h, int s0, int s1) {
    auto* b = (GroupModelBox*)h;
    return guard([&] { b->model->set_subject_shard(s0, s1); });
}
int msmhost_group_rotated_all(void* h, int subject, double* out /*[L][Vt]*/) {
    auto* b = (GroupModelBox*)h;
    return guard([&] {
        const int L = b->model->getNumLabels();
        for (int l = 0; l < L; l++) { const auto& v = b->model->rotated_mesh_values(subject, l); std::memcpy(out + (size_t)l * v.size(), v.data(), v.size() * sizeof(double)); }
    });
}
int msmhost_group_set_rotated(void* h, int subject, const double* in /*[L][Vt]*/) {
    auto* b = (GroupModelBox*)h;
    return guard([&] { b->model->set_rotated_values(subject, in); });
}
int msmhost_group_finish_setup(void* h) {   // get_patch_data + set_patch_data once every subject's rotated data is in place
    auto* b = (GroupModelBox*)h;
    return guard([&] { b->model->finish_setup(); });
}
#endif
#ifdef WITH_REF_OPTIM
int ref_fusion_group(void* h, int threads, double* energy) {   // src/group_mesh_registration.cpp:74
    auto* b = (GroupModelBox*)h;
    return guard([&] { *energy = Fusion::optimize(b->model, false, threads); });
}
#ifndef WITH_REF_COST
int ref_fusion_tables_group(void* h, int threads, double* energy) {   // host/src/FusionTables.h optimize_group: batched GPU tables
    auto* b = (GroupModelBox*)h;
    return guard([&] { *energy = FusionTables::optimize_group(b->model, false, threads); });
}
#endif
#endif

#ifdef WITH_REF_COST
// ---- the reference's DiscreteGroupCostFunction (src/DiscreteGroupCostFunction.{h,cpp}), same flat interface as the
// oracle port's orc_gc_* (oracle/oracle_capi.cpp) so the two can be compared query by query -----------------------------
namespace {
struct GroupBox {
    DiscreteGroupCostFunction cf;
    std::vector<int> triplets, pairs;
    int L = 0;
};
}  // namespace
// patch lists of the last setupCostFunction: sizes[N]; idx (may be NULL) receives the concatenated source vertex ids
int ref_model_patches(void* h, int* sizes, int* idx) {
    auto* b = (ModelBox*)h;
    return guard([&] {
        size_t o = 0;
        const auto& P = b->model->patches();
        for (size_t k = 0; k < P.size(); k++) {
            sizes[k] = (int)P[k].size();
            if (idx) std::copy(P[k].begin(), P[k].end(), idx + o);
            o += P[k].size();
        }
    });
}
// Bounded sample of ONE outer iteration through the reference's own cost-function classes (bench.py --impl reference /
// cpu_baseline): the model cannot evaluate a subset of control points and a full HCP-scale set-up is ~1 minute of
// single-threaded brute-force patch search, so the cost functions are driven directly, in the order setupCostFunction
// uses (src/DiscreteModel.cpp:188-236), on
//   * the first `rows` control points for get_source_data + computeUnaryCosts (ONE thread: the reference's threaded
//     unary loop is racy), and
//   * `ntrip` triplets for the 8 fusion-move costs of every proposal label (computeTripletCost is re-entrant: `threads`
//     OpenMP threads, as include/Fusion/Fusion.h:180 does).
// seconds[3] = get_source_data, computeUnaryCosts, triplet moves.
int ref_cost_sample(const double* txyz, int Vt, const int* ttris, int Ft, const double* sxyz, const double* input, const double* ref, int C,
                    int multivariate, const double* cp0_xyz, const double* cp_xyz, int N, const int* ctris, int Tc, const double* maxsep,
                    double mvdmax, const double* labels, int L, double lambda, double rexp, double mu, double kappa, double kexp, double range,
                    int simmeasure, int rmode, int rows, int ntrip, const int* triplets, const int* labeling, int threads, double* unary_out,
                    double* moves_out, double* seconds) {
    return guard([&] {
        myparam P = default_parameters();
        P["lambda"] = (float)lambda; P["exponent"] = (float)rexp; P["shearmodulus"] = (float)mu; P["bulkmodulus"] = (float)kappa;
        P["kexponent"] = (float)kexp; P["range"] = (float)range; P["simmeasure"] = simmeasure; P["regularisermode"] = rmode;
        P["numthreads"] = 1;
        const newresampler::Mesh target = mesh_from(txyz, Vt, ttris, Ft);
        const newresampler::Mesh source = mesh_from(sxyz, Vt, ttris, Ft);
        auto feat = make_features(input, Vt, ref, Vt, C);
        std::vector<newresampler::Point> lab(L);
        for (int i = 0; i < L; i++) lab[i] = newresampler::Point(labels[3 * i], labels[3 * i + 1], labels[3 * i + 2]);
        auto rotations = [&](const double* cp, int n) {   // src/DiscreteModel.cpp:284-294 (labels[0] = label-grid centre)
            std::vector<NEWMAT::Matrix> R(n);
            for (int k = 0; k < n; k++) R[k] = estimate_rotation_matrix(lab[0], newresampler::Point(cp[3 * k], cp[3 * k + 1], cp[3 * k + 2]));
            return R;
        };
        NEWMAT::Matrix ones; ones.ReSize(1, Vt); ones = 1;
        seconds[0] = seconds[1] = seconds[2] = 0.0;
        if (rows > 0) {
            std::shared_ptr<NonLinearSRegDiscreteCostFunction> cf;
            if (multivariate) cf = std::make_shared<MultivariateNonLinearSRegDiscreteCostFunction>();
            else cf = std::make_shared<UnivariateNonLinearSRegDiscreteCostFunction>();
            cf->set_parameters(P);
            const newresampler::Mesh grid = mesh_from(cp_xyz, rows, nullptr, 0);   // the sampled control points, vertices only
            cf->set_meshes(target, source, grid);
            cf->set_featurespace(feat);
            NEWMAT::ColumnVector ms(rows);
            for (int k = 0; k < rows; k++) ms(k + 1) = maxsep[k];
            cf->set_spacings(ms, mvdmax);
            std::shared_ptr<newresampler::Octree> tree = std::make_shared<newresampler::Octree>(target);   // :88
            cf->set_octrees(tree);
            cf->set_dataaffintyweighting(ones);                                   // src/mesh_registration.cpp:330-335
            cf->set_labels(lab, rotations(cp_xyz, rows));
            cf->initialize(rows, L, 0, 0);
            double t0 = omp_get_wtime();
            cf->get_source_data();
            double t1 = omp_get_wtime();
            cf->computeUnaryCosts();
            double t2 = omp_get_wtime();
            seconds[0] = t1 - t0; seconds[1] = t2 - t1;
            if (unary_out) std::memcpy(unary_out, cf->getUnaryCosts(), sizeof(double) * (size_t)rows * L);
        }
        if (ntrip > 0) {
            UnivariateNonLinearSRegDiscreteCostFunction cf;
            cf.set_parameters(P);
            const newresampler::Mesh grid0 = mesh_from(cp0_xyz, N, ctris, Tc);
            cf.set_meshes(target, target, grid0);                                 // _ORIG = the undeformed template (:175-177)
            cf.set_featurespace(feat);
            cf.reset_CPgrid(mesh_from(cp_xyz, N, ctris, Tc));
            cf.set_initial_angles(grid0.get_face_angles());
            NEWMAT::ColumnVector ms(N);
            for (int k = 0; k < N; k++) ms(k + 1) = maxsep[k];
            cf.set_spacings(ms, mvdmax);
            cf.set_labels(lab, rotations(cp_xyz, N));
            std::vector<int> trip(triplets, triplets + 3 * (size_t)ntrip);
            cf.initialize(N, L, 0, ntrip);
            cf.setTriplets(trip.data());
            double t0 = omp_get_wtime();
#pragma omp parallel for collapse(2) num_threads(threads)
            for (int a = 0; a < L; a++)
                for (int t = 0; t < ntrip; t++) {
                    const int cur[3] = {labeling[trip[3 * t]], labeling[trip[3 * t + 1]], labeling[trip[3 * t + 2]]};
                    for (int m = 0; m < 8; m++) {
                        const double v = cf.computeTripletCost(t, (m & 4) ? a : cur[0], (m & 2) ? a : cur[1], (m & 1) ? a : cur[2]);
                        if (moves_out) moves_out[((size_t)a * ntrip + t) * 8 + m] = v;
                    }
                }
            seconds[2] = omp_get_wtime() - t0;
        }
    });
}

void* ref_gc_create() { return new GroupBox(); }
void ref_gc_destroy(void* h) { delete (GroupBox*)h; }
int ref_gc_set(void* h, double lambda, double rexp, double mu, double kappa, int S, const double* cp_xyz /*S*N*3*/, int N, const int* cp_tris,
               int T, const double* orig_xyz /*N*3*/, const double* labels, int L, const double* ROT /*S*N*9*/, const int* triplets /*S*T*3*/) {
    auto* g = (GroupBox*)h;
    return guard([&] {
        myparam P = default_parameters();
        P["lambda"] = (float)lambda; P["exponent"] = (float)rexp; P["shearmodulus"] = (float)mu; P["bulkmodulus"] = (float)kappa;
        g->cf.set_parameters(P);
        const newresampler::Mesh orig = mesh_from(orig_xyz, N, cp_tris, T);
        g->cf.set_meshes(orig, orig, orig, S);                                     // src/DiscreteGroupModel.cpp:139
        for (int s = 0; s < S; s++) g->cf.reset_CPgrid(mesh_from(cp_xyz + (size_t)s * N * 3, N, cp_tris, T), s);
        std::vector<newresampler::Point> lab(L);
        for (int i = 0; i < L; i++) lab[i] = newresampler::Point(labels[3 * i], labels[3 * i + 1], labels[3 * i + 2]);
        std::vector<NEWMAT::Matrix> rot((size_t)S * N, NEWMAT::Matrix(3, 3));
        for (size_t k = 0; k < rot.size(); k++)
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) rot[k](i + 1, j + 1) = ROT[9 * k + 3 * i + j];
        g->cf.set_labels(lab, rot);
        g->L = L;
        g->triplets.assign(triplets, triplets + (size_t)S * T * 3);
        g->cf.initialize(S * N, L, (int)g->pairs.size() / 2, S * T);
        g->cf.setTriplets(g->triplets.data());
    });
}
int ref_gc_triplet_costs(void* h, int n, const int* q, double* out) {
    auto* g = (GroupBox*)h;
    return guard([&] { for (int i = 0; i < n; i++) out[i] = g->cf.computeTripletCost(q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]); });
}
int ref_gc_set_patches(void* h, int nmaps, const long long* off, const int* keys, const double* vals, const int* pairs, int np) {
    auto* g = (GroupBox*)h;
    return guard([&] {
        std::vector<std::map<int, double>> pd(nmaps);
        for (int m = 0; m < nmaps; m++)
            for (long long e = off[m]; e < off[m + 1]; e++) pd[m][keys[e]] = vals[e];
        g->cf.set_patch_data(pd);
        g->pairs.assign(pairs, pairs + 2 * (size_t)np);
        g->cf.setPairs(g->pairs.data());
    });
}
int ref_gc_pair_costs(void* h, int n, const int* q, int /*L*/, double* out) {
    auto* g = (GroupBox*)h;
    return guard([&] { for (int i = 0; i < n; i++) out[i] = g->cf.computePairwiseCost(q[3 * i], q[3 * i + 1], q[3 * i + 2]); });
}
#endif

#ifdef WITH_REF_OPTIM
// ---- unmodified reference optimisers ------------------------------------------------------------------------------
int ref_available() { return 1; }
// Mesh_registration::run_discrete_opt (src/mesh_registration.cpp:316-397) restated on the mirror classes with the
// reference's own optimisers: per outer iteration reset_meshspace -> setupCostFunction -> FastPD | HOCR -> convergence
// test -> applyLabeling -> warp the source by the control-grid update -> unfold both.  `source` (the data mesh, same
// topology as the target template) is updated in place; labelings/energies/seconds are per iteration (size iters).
// Returns the number of iterations run, or -1 on error.
int ref_run_discrete_opt(void* h, int iters, int threads, double* source_xyz, int* labelings, double* energies, double* seconds_cost,
                         double* seconds_opt, double* seconds_warp) {
    auto* b = (ModelBox*)h;
    int done = 0;
    const int st = guard([&] {
        const std::string dopt = boost::get<std::string>(b->params.find("dOPT")->second);
        newresampler::Mesh source = b->source;
        assign_coords(source, source_xyz);
        double energy = 0.0, newenergy = 0.0;
        for (int iter = 1; iter <= iters; iter++) {
            double t0 = omp_get_wtime();
            NEWMAT::Matrix CombinedWeight;                       // :330-334 (no cost-function weighting supplied)
            CombinedWeight.resize(1, source.nvertices());
            CombinedWeight = 1;
            b->model->setupCostFunctionWeighting(CombinedWeight);
            b->model->reset_meshspace(source);
            b->model->setupCostFunction();
            double t1 = omp_get_wtime();
            if (dopt == "FastPD") {
                b->model->computeUnaryCosts();
                b->model->computePairwiseCosts();
                FPD::FastPD opt(b->model, 100);
                newenergy = opt.run();
                opt.getLabeling(b->model->getLabeling());
            } else if (dopt == "HOCR") {
#ifndef WITH_REF_COST
                // MSM_HOCR_DRIVER=tables: the table-aware fusion driver (src/FusionTables.h: batched GPU move tables, flat
                // binary MRF) in place of the reference's Fusion::optimize; same labellings (tests/test_gpu_model.py)
                const char* drv = std::getenv("MSM_HOCR_DRIVER");
                if (drv && std::string(drv) == "tables") newenergy = FusionTables::optimize(b->model, false, threads);
                else
#endif
                newenergy = Fusion::optimize(b->model, false, threads);
            } else throw MeshregException("Unrecognized optimiser");
            double t2 = omp_get_wtime();
            std::memcpy(labelings + (size_t)(iter - 1) * b->model->getNumNodes(), b->model->getLabeling(), sizeof(int) * b->model->getNumNodes());
            energies[iter - 1] = newenergy;
            seconds_cost[iter - 1] = t1 - t0; seconds_opt[iter - 1] = t2 - t1; seconds_warp[iter - 1] = 0;
            done = iter;
            if (iter > 1 && ((iter - 1) % 2 == 0) && (energy - newenergy < 0.001)) break;   // :370-377
            newresampler::Mesh previous_controlgrid = b->model->get_CPgrid();
            b->model->applyLabeling();
            newresampler::Mesh transformed_controlgrid = b->model->get_CPgrid();
            newresampler::barycentric_mesh_interpolation(source, previous_controlgrid, transformed_controlgrid, threads);
            unfold(transformed_controlgrid, false);
            b->model->reset_CPgrid(transformed_controlgrid);
            unfold(source, false);
            energy = newenergy;
            seconds_warp[iter - 1] = omp_get_wtime() - t2;
        }
        auto c = coords_of(source);
        std::memcpy(source_xyz, c.data(), c.size() * sizeof(double));
    });
    return st == 0 ? done : -1;
}
#ifdef WITH_REF_DRIVER_LOOP
// Mesh_registration::run_discrete_opt ITSELF (the text of src/mesh_registration.cpp:316-397, see the top of this file) on this
// box's model: `iters` outer iterations, optimiser = the model's dOPT ("HOCR" / "FastPD" / "MCMC" with mciters sweeps), unit
// cost-function weighting.  source_xyz is updated in place; cpgrid_out / labeling_out (optional) = the model's final state.
int ref_driver_run_discrete_opt(void* h, int iters, int threads, int mciters, double* source_xyz, double* cpgrid_out, int* labeling_out) {
    auto* b = (ModelBox*)h;
    return guard([&] {
        Mesh_registration reg;
        reg.model = b->model;
        reg.PARAMETERS = b->params;
        reg.PARAMETERS["iters"] = iters;
        reg._discreteOPT = boost::get<std::string>(b->params.find("dOPT")->second);
        reg._numthreads = threads;
        reg._mciters = {mciters};
        newresampler::Mesh source = b->source;
        assign_coords(source, source_xyz);
        reg.run_discrete_opt(source);
        auto c = coords_of(source);
        std::memcpy(source_xyz, c.data(), c.size() * sizeof(double));
        if (cpgrid_out) { auto g = coords_of(b->model->get_CPgrid()); std::memcpy(cpgrid_out, g.data(), g.size() * sizeof(double)); }
        if (labeling_out) std::memcpy(labeling_out, b->model->getLabeling(), sizeof(int) * b->model->getNumNodes());
    });
}
#endif
// src/mesh_registration.cpp:346-352
int ref_fastpd_model(void* h, double* energy) {
    auto* b = (ModelBox*)h;
    return guard([&] {
        b->model->computeUnaryCosts();
        b->model->computePairwiseCosts();
        FPD::FastPD opt(b->model, 100);
        *energy = opt.run();
        opt.getLabeling(b->model->getLabeling());
    });
}
// :358
int ref_fusion_model(void* h, int threads, double* energy) {
    auto* b = (ModelBox*)h;
    return guard([&] { *energy = Fusion::optimize(b->model, false, threads); });
}
#ifndef WITH_REF_COST
// the table-aware fusion driver (host/src/FusionTables.h): batched GPU tables in, the reference's ELC / FastPD cores unchanged
int ref_fusion_tables_model(void* h, int threads, double* energy) {
    auto* b = (ModelBox*)h;
    return guard([&] { *energy = FusionTables::optimize(b->model, false, threads); });
}
#endif
int ref_fastpd_tables(int N, int L, int P, const int* pairs, const double* unary, const double* pairtab, int* labeling, double* energy) {
    return guard([&] {
        auto cf = std::make_shared<TableCostFunction>(N, L, P, 0, unary, pairtab, nullptr);
        auto model = std::make_shared<TableModel>(N, L, P, pairs, 0, nullptr, cf);
        cf->setPairs(const_cast<int*>(model->getPairs()));
        FPD::FastPD opt(model, 100);
        *energy = opt.run();
        opt.getLabeling(labeling);
    });
}
int ref_fusion_tables(int N, int L, int P, const int* pairs, int T, const int* triplets, const double* unary, const double* pairtab,
                      const double* triptab, int threads, int* labeling, double* energy) {
    return guard([&] {
        auto cf = std::make_shared<TableCostFunction>(N, L, P, T, unary, pairtab, triptab);
        auto model = std::make_shared<TableModel>(N, L, P, pairs, T, triplets, cf);
        *energy = Fusion::optimize(model, false, threads);
        std::memcpy(labeling, model->getLabeling(), sizeof(int) * N);
    });
}
#endif

}  // extern "C"
