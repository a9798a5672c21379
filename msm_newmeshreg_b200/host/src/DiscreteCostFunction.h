// DiscreteCostFunction.h — host-side mirror of the reference's cost-function classes
// (src/DiscreteCostFunction.h:14-63 base, :65-193 NonLinearSRegDiscreteCostFunction, :195-209 Uni/Multivariate),
// same names, same virtuals, same table layouts — but every cost is evaluated by the sm_100a kernels behind the C ABI
// of include/msm_b200.h.  The per-element virtuals the optimisers call concurrently
// (include/Fusion/Fusion.h:147-195, include/FastPD/FastPD.h:40) are re-entrant table look-ups.
//
// Deliberate differences from the reference (all invisible to the optimisers):
//  * tripletcosts (T*L^3 doubles, src/DiscreteCostFunction.cpp:17-21) is never allocated: nothing reads it in the
//    reference (DiscreteModel::computeTripletCosts is an empty stub); getTripletCosts() returns nullptr.
//  * the unary table is computed eagerly in get_source_data() and kept pristine; computeUnaryCosts() copies it into
//    unarycosts, which FastPD then overwrites in place (include/FastPD/FastPD.h:126).
//  * set_labels ignores the ROT argument: ROT[k] = rotation(labels[0] -> CP_k) is recomputed on the device
//    (labels[0] is the label-grid centre in both label sets, src/DiscreteModel.cpp:115-116).
//  * HO tri-clique variants and regulariser modes 4/5 (anatomical meshes) are out of scope.
#pragma once
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>

#include "../../../include/msm_b200.h"
#include "featurespace.h"
#include "mesh_warp.h"

#define RAD 100.0
#define EPSILON 1.0E-8

namespace newmeshreg {

class DiscreteCostFunction {
public:
    virtual ~DiscreteCostFunction() {
        delete[] unarycosts;
        delete[] paircosts;
    }
    //---SET--//
    virtual void reset() {  // src/DiscreteCostFunction.cpp:35-39
        if (unarycosts) std::fill(unarycosts, unarycosts + (size_t)m_num_labels * m_num_nodes, 0.0);
        if (paircosts) std::fill(paircosts, paircosts + (size_t)m_num_labels * m_num_labels * m_num_pairs, 0.0);
    }
    void setPairs(int* p) { _pairs = p; pairs_dirty = true; }
    void setTriplets(int* p) { _triplets = p; triplets_dirty = true; }
    //---GET--//
    double* getUnaryCosts() { return unarycosts; }
    double* getPairwiseCosts() { return paircosts; }
    double* getTripletCosts() { return nullptr; }
    //---COMPUTE--//
    virtual void computeUnaryCosts() {}
    virtual double computeUnaryCost(int node, int label) { return 0; }
    virtual void computePairwiseCosts(const int* pairs) {}
    virtual double computePairwiseCost(int pair, int labelA, int labelB) { return 0; }
    virtual double computeTripletCost(int triplet, int labelA, int labelB, int labelC) { return 0; }
    virtual double evaluateTotalCostSum(const int* labeling, const int* pairs, const int* triplets) {  // :41-70
        double su = 0.0, sp = 0.0, st = 0.0;
        for (int i = 0; i < m_num_nodes; ++i) su += computeUnaryCost(i, labeling[i]);
        for (int p = 0; p < m_num_pairs; ++p) sp += computePairwiseCost(p, labeling[pairs[p * 2]], labeling[pairs[p * 2 + 1]]);
        for (int t = 0; t < m_num_triplets; ++t)
            st += computeTripletCost(t, labeling[triplets[t * 3]], labeling[triplets[t * 3 + 1]], labeling[triplets[t * 3 + 2]]);
        if (_verbosity)
            std::cout << "cost_sum_unary " << su << " cost_sum_pairwise " << sp << " cost_sum_triplet " << st << " total " << su + sp + st
                      << " m_num_triplets 2 " << m_num_triplets << std::endl;
        return su + sp + st;
    }

protected:
    virtual void initialize(int numNodes, int numLabels, int numPairs, int numTriplets) {  // :5-33 (minus tripletcosts)
        if (m_num_nodes != numNodes || m_num_labels != numLabels) {
            delete[] unarycosts;
            unarycosts = new double[(size_t)numNodes * numLabels];
        }
        if (m_num_pairs != numPairs || m_num_labels != numLabels) {
            delete[] paircosts;
            paircosts = new double[(size_t)numPairs * numLabels * numLabels];
        }
        m_num_nodes = numNodes; m_num_labels = numLabels; m_num_pairs = numPairs; m_num_triplets = numTriplets;
        reset();
    }

    int m_num_nodes = 0, m_num_labels = 0, m_num_pairs = 0, m_num_triplets = 0;
    double* unarycosts = nullptr;
    double* paircosts = nullptr;
    int* _pairs = nullptr;
    int* _triplets = nullptr;
    bool pairs_dirty = true, triplets_dirty = true;
    float _reglambda = 1.0;
    int _threads = 1;
    bool _debug = false;
    bool _verbosity = false;
};

class NonLinearSRegDiscreteCostFunction : public DiscreteCostFunction {
public:
    NonLinearSRegDiscreteCostFunction() {
        int dev = 0;
        if (const char* e = std::getenv("MSM_DEVICE")) dev = std::atoi(e);
        if (msm_create(dev, &ctx) != 0) throw MeshregException(dup_msg(msm_last_error(nullptr)));
    }
    ~NonLinearSRegDiscreteCostFunction() override { msm_destroy(ctx); }
    NonLinearSRegDiscreteCostFunction(const NonLinearSRegDiscreteCostFunction&) = delete;
    NonLinearSRegDiscreteCostFunction& operator=(const NonLinearSRegDiscreteCostFunction&) = delete;

    //---INIT---//
    void initialize(int numNodes, int numLabels, int numPairs, int numTriplets) override {  // :111-119
        if (_TARGET.nvertices() == 0 || _SOURCE.nvertices() == 0)
            throw MeshregException("CostFunction::You must supply source and target meshes.");
        if (_HIGHREScfweight.Ncols() != _SOURCE.nvertices()) {
            _HIGHREScfweight.ReSize(1, _SOURCE.nvertices());
            _HIGHREScfweight = 1.0;  // documented intent of the reference: unit weights (src/mesh_registration.cpp:330-334)
        }
        if (_HIGHREScfweight.Nrows() != 1 && FEAT && _HIGHREScfweight.Nrows() != FEAT->get_dim())
            throw MeshregException("DiscreteModel ERROR:: costfunction weighting has dimensions incompatible with data");
        DiscreteCostFunction::initialize(numNodes, numLabels, numPairs, numTriplets);
        tables_epoch++;  // every look-up table of the previous outer iteration is stale now
    }

    virtual void set_parameters(myparam& ALLPARAMS) {  // :121-136
        myparam::iterator it;
        it = ALLPARAMS.find("exponent"); _rexp = boost::get<float>(it->second);
        it = ALLPARAMS.find("weight"); _dweight = boost::get<bool>(it->second);
        it = ALLPARAMS.find("anorm"); _anorm = boost::get<bool>(it->second);
        it = ALLPARAMS.find("shearmodulus"); _mu = boost::get<float>(it->second);
        it = ALLPARAMS.find("bulkmodulus"); _kappa = boost::get<float>(it->second);
        it = ALLPARAMS.find("kexponent"); _k_exp = boost::get<float>(it->second);
        it = ALLPARAMS.find("lambda"); _reglambda = boost::get<float>(it->second);
        it = ALLPARAMS.find("range"); _controlptrange = boost::get<float>(it->second);
        it = ALLPARAMS.find("simmeasure"); _simmeasure = boost::get<int>(it->second);
        it = ALLPARAMS.find("verbosity"); _verbosity = boost::get<bool>(it->second);
        it = ALLPARAMS.find("regularisermode"); _rmode = boost::get<int>(it->second);
        it = ALLPARAMS.find("numthreads"); _threads = boost::get<int>(it->second);
        it = ALLPARAMS.find("dOPT"); dopt = boost::get<std::string>(it->second);
        params_dirty = true;
        tables_epoch++;   // every cached table was produced with the old parameters
    }

    //---COST CALCULATION---//
    void computeUnaryCosts() override {  // :306-313 — the table was produced by the GPU in get_source_data()
        ensure_unary();
        std::memcpy(unarycosts, unary_pristine.data(), sizeof(double) * unary_pristine.size());
    }
    double computeUnaryCost(int node, int label) override {
        ensure_unary();
        return unary_pristine[(size_t)label * m_num_nodes + node];
    }
    double computePairwiseCost(int pair, int labelA, int labelB) override {  // :256-296 via the :303 table layout
        ensure_pair_table();
        return (double)pair_table[((size_t)pair * m_num_labels + labelB) * m_num_labels + labelA];
    }
    void computePairwiseCosts(const int* pairs) override {  // :298-304
        ensure_pair_table();
        for (size_t i = 0; i < pair_table.size(); i++) paircosts[i] = (double)pair_table[i];
    }
    double computeTripletCost(int triplet, int labelA, int labelB, int labelC) override {  // :138-254
        ensure_triplet_table();
        const size_t L = m_num_labels;
        return (double)triplet_table[(((size_t)triplet * L + labelA) * L + labelB) * L + labelC];
    }
    // Lean path for a table-aware fusion driver: the 8 move costs of every triplet for label
    // alpha (out[T][8]) or all labels (alpha = -1, out[L][T][8]); include/Fusion/Fusion.h:187-194 bit order.
    void computeTripletMoves(const int* labeling, int alpha, double* out) {
        sync_device_state();
        check(msm_triplet_moves(ctx, labeling, alpha, out));
    }
    // batched accessors for a table-aware driver (host/src/FusionTables.h): the pristine unary table [L][N], and the total
    // energy of a labelling evaluated on the device (src/DiscreteCostFunction.cpp:41-70) without any T*L^3 table
    const std::vector<double>& unaryTable() { ensure_unary(); return unary_pristine; }
    double totalCostBatched(const int* labeling) {
        ensure_unary();
        sync_device_state();
        double sums[3] = {0, 0, 0};
        check(msm_total_cost(ctx, labeling, sums));
        return sums[0] + sums[1] + sums[2];
    }
    void computeTripletMoves(const int* labeling, int alpha, float* out) {  // fp32 table: half the read-back
        sync_device_state();
        check(msm_triplet_moves_f32(ctx, labeling, alpha, out));
    }

    void set_dataaffintyweighting(const NEWMAT::Matrix& HRWeight) { _HIGHREScfweight = HRWeight; source_dirty = true; }

    //---GET DATA---//
    // resample_weights (:364-383): AbsoluteWeights = metric_resample (adaptive barycentric, [EXT] A8) of the per-vertex maximum
    // over the weight rows, from the current source mesh onto the current control grid.  Unit weights (the driver's default,
    // src/mesh_registration.cpp:330-334) resample to exactly 1 whatever the kernel: skipped.
    void resample_weights() {
        if (!abs_weights_given) {
            bool unit = true;
            const int V = _SOURCE.nvertices(), R = _HIGHREScfweight.Ncols() == V ? _HIGHREScfweight.Nrows() : 0;
            for (int j = 1; j <= R && unit; j++)
                for (int k = 1; k <= V; k++) if (_HIGHREScfweight(j, k) != 1.0) { unit = false; break; }
            if (unit) {
                if (!abs_weights.empty()) { abs_weights.assign(_CPgrid.nvertices(), 1.0); check(msm_set_abs_weights(ctx, abs_weights.data())); abs_weights.clear(); }
                return;
            }
            std::vector<double> mx(V);
            for (int k = 1; k <= V; k++) {
                double m = std::numeric_limits<double>::lowest();
                for (int j = 1; j <= R; j++) if (_HIGHREScfweight(j, k) > m) m = _HIGHREScfweight(j, k);
                mx[k - 1] = m;
            }
            abs_weights = newresampler::metric_resample(_SOURCE, mx.data(), 1, _CPgrid);
            check(msm_set_abs_weights(ctx, abs_weights.data()));
        }
    }
    const std::vector<double>& absolute_weights() const { return abs_weights; }

    virtual void get_source_data() {  // :394-410 / :458-481 : patches, then the whole unary table in one launch
        { HostTimer t("    sync_device_state"); sync_device_state(); }
        { HostTimer t("    resample_weights"); resample_weights(); }
        { HostTimer t("    msm_build_patches"); check(msm_build_patches_async(ctx)); }   // status read by msm_unary_table below
        unary_pristine.resize((size_t)m_num_labels * m_num_nodes);   // every entry is written by msm_unary_table
        { HostTimer t("    msm_unary_table"); check(msm_unary_table(ctx, unary_pristine.data())); }
        unary_epoch = tables_epoch.load();
    }

    //---SET FUNCTION SPACE---//
    virtual void set_meshes(const newresampler::Mesh& target, const newresampler::Mesh& source, const newresampler::Mesh& GRID, int num = 1) {
        _TARGET = target; _SOURCE = source; _ORIG = source; _CPgrid = GRID; _oCPgrid = GRID;
        target_dirty = source_dirty = grid_dirty = true;
    }
    inline void set_iter(int iter) { _iter = iter; }
    void set_featurespace(const std::shared_ptr<featurespace>& features) { FEAT = features; target_dirty = source_dirty = true; }
    void set_labels(const std::vector<newresampler::Point>& labellist,
                    const std::vector<NEWMAT::Matrix>& ROT = std::vector<NEWMAT::Matrix>()) {
        _labels = labellist;
        labels_dirty = true;
    }
    virtual void set_patch_data(const std::vector<std::map<int, double>>& patches) {}
    virtual void set_spacings(const NEWMAT::ColumnVector& spacings, double MAX) { MAXSEP = spacings; MVDmax = MAX; grid_dirty = params_dirty = true; tables_epoch++; }
    void set_octrees(std::shared_ptr<newresampler::Octree>&) {}  // the device-side hierarchy is built in set_target
    virtual void reset_source(const newresampler::Mesh& source, int num = 0) { _SOURCE = source; source_xyz_dirty = true; }
    virtual void reset_CPgrid(const newresampler::Mesh& grid, int num = 0) { _CPgrid = grid; cp_xyz_dirty = true; }
    void set_initial_angles(const std::vector<std::vector<double>>& angles) {  // :95-102
        double meanang = 0.0;
        for (const auto& a : angles) for (double j : a) meanang += j;
        _MEANANGLE = meanang / (3.0 * angles.size());
        params_dirty = true;
    }
    void initialize_regulariser() {}  // anatomical modes only
    void reset_anatomical() {}
    // src/DiscreteCostFunction.h:146-163 — anatomical regulariser modes 4/5: outside the accelerated path
    void set_anatomical(const newresampler::Mesh&, const newresampler::Mesh&, const newresampler::Mesh&, const newresampler::Mesh&) {
        throw MeshregException("set_anatomical: the anatomical regulariser modes (regoption 4/5) are not part of the GPU cost-evaluation path");
    }
    void set_anatomical_neighbourhood(const std::vector<std::map<int, double>>&, const std::vector<std::vector<int>>&) {
        throw MeshregException("set_anatomical_neighbourhood: the anatomical regulariser modes (regoption 4/5) are not part of the GPU cost-evaluation path");
    }
    // AbsoluteWeights (resample_weights, :364-383): identically 1 for unit weights; otherwise supplied by the caller
    // because the reference obtains them from FSL's metric_resample ([EXT] A8).
    void set_absolute_weights(const std::vector<double>& aw) { abs_weights = aw; abs_weights_given = true; grid_dirty = true; }

    //---REPORT AND DEBUG---//
    void report() {}
    void debug() { _debug = true; }
    void set_mcmc_threads(int threads) { mcmc_threads = threads; }
    msm_ctx* context() { return ctx; }

protected:
    static const char* dup_msg(const char* m) {  // MeshregException keeps a raw pointer
        static thread_local std::string keep;
        keep = m ? m : "";
        return keep.c_str();
    }
    void check(int status) {
        if (status != 0) throw MeshregException(dup_msg(msm_last_error(ctx)));
    }
    virtual bool is_multivariate() const { return false; }
    virtual bool is_group() const { return false; }

    // push whatever changed on the host side of the boundary to the device context
    void sync_device_state() {
        std::lock_guard<std::mutex> lock(mtx);
        if (target_dirty) {
            if (!FEAT) throw MeshregException("CostFunction::featurespace not set");
            auto xyz = _TARGET.coords(); auto tr = _TARGET.faces();
            check(msm_set_target(ctx, xyz.data(), _TARGET.nvertices(), tr.data(), _TARGET.ntriangles(), FEAT->ref_data(), FEAT->get_dim()));
            target_dirty = false;
        }
        if (source_dirty) {
            auto xyz = _SOURCE.coords();
            const int rows = _HIGHREScfweight.Ncols() == _SOURCE.nvertices() ? _HIGHREScfweight.Nrows() : 0;
            check(msm_set_source(ctx, xyz.data(), _SOURCE.nvertices(), FEAT->input_data(), FEAT->get_dim(), rows ? _HIGHREScfweight.data() : nullptr,
                                 rows, is_multivariate() ? 1 : 0));
            source_dirty = source_xyz_dirty = false;
        }
        if (source_xyz_dirty) {
            HostTimer t("      reset_source upload");
            check(msm_reset_source(ctx, _SOURCE.coord_data()));   // Point is three packed doubles: no flattening copy
            source_xyz_dirty = false;
        }
        if (grid_dirty) {
            auto oxyz = _oCPgrid.coords(); auto tr = _oCPgrid.faces();
            std::vector<double> ms(_oCPgrid.nvertices(), 0.0);
            for (int k = 0; k < std::min(MAXSEP.Nrows(), _oCPgrid.nvertices()); k++) ms[k] = MAXSEP(k + 1);
            std::vector<double> orig((size_t)_oCPgrid.nvertices() * 3);
            for (int k = 0; k < _oCPgrid.nvertices(); k++) {  // _ORIG is indexed with CP ids (nested numbering, Q3)
                const newresampler::Point& p = orig_coord(k);
                orig[3 * k] = p.X; orig[3 * k + 1] = p.Y; orig[3 * k + 2] = p.Z;
            }
            check(msm_set_cpgrid(ctx, oxyz.data(), _oCPgrid.nvertices(), tr.data(), _oCPgrid.ntriangles(), ms.data(), orig.data(),
                                 (int)abs_weights.size() == _oCPgrid.nvertices() ? abs_weights.data() : nullptr));
            grid_dirty = false;
            cp_xyz_dirty = true;
            labels_dirty = true;
            configure_group();
        }
        if (cp_xyz_dirty) {
            auto xyz = current_cp_coords();
            check(msm_reset_cpgrid(ctx, xyz.data()));
            cp_xyz_dirty = false;
        }
        if (params_dirty) {
            msm_params p;
            msm_default_params(&p);
            p.lambda = _reglambda; p.rexp = _rexp; p.mu = _mu; p.kappa = _kappa; p.k_exp = _k_exp; p.range = _controlptrange;
            p.simmeasure = _simmeasure; p.rmode = _rmode; p.dweight = _dweight; p.anorm = _anorm; p.meanangle = _MEANANGLE;
            p.mvdmax = MVDmax; p.group = is_group() ? 1 : 0;
            check(msm_set_params(ctx, &p));
            params_dirty = false;
        }
        if (labels_dirty && !_labels.empty()) {
            std::vector<double> lab(_labels.size() * 3);
            for (size_t i = 0; i < _labels.size(); i++) { lab[3 * i] = _labels[i].X; lab[3 * i + 1] = _labels[i].Y; lab[3 * i + 2] = _labels[i].Z; }
            check(msm_set_labels(ctx, lab.data(), (int)_labels.size(), lab.data()));  // labels[0] = label-grid centre
            labels_dirty = false;
        }
        if (pairs_dirty && _pairs && m_num_pairs > 0) { check(msm_set_pairs(ctx, _pairs, m_num_pairs)); pairs_dirty = false; }
        if (triplets_dirty && _triplets && m_num_triplets > 0) { check(msm_set_triplets(ctx, _triplets, m_num_triplets)); triplets_dirty = false; }
    }
    virtual const newresampler::Point& orig_coord(int k) const { return _ORIG.get_coord(k); }
    virtual std::vector<double> current_cp_coords() const { return _CPgrid.coords(); }
    virtual void configure_group() {}

    void ensure_unary() {
        if (unary_epoch == tables_epoch.load()) return;
        std::lock_guard<std::mutex> lock(lazy);
        if (unary_epoch != tables_epoch.load()) get_source_data();
    }
    void ensure_pair_table() {
        if (pair_epoch == tables_epoch.load()) return;
        std::lock_guard<std::mutex> lock(lazy);
        if (pair_epoch == tables_epoch.load()) return;
        sync_device_state();
        pair_table.assign((size_t)m_num_pairs * m_num_labels * m_num_labels, 0.f);
        if (m_num_pairs > 0) check(msm_pair_table(ctx, pair_table.data()));
        pair_epoch = tables_epoch.load();
    }
    void ensure_triplet_table() {
        if (triplet_epoch == tables_epoch.load()) return;
        std::lock_guard<std::mutex> lock(lazy);
        if (triplet_epoch == tables_epoch.load()) return;
        sync_device_state();
        const size_t L = m_num_labels;
        triplet_table.assign((size_t)m_num_triplets * L * L * L, 0.f);
        if (m_num_triplets > 0) check(msm_triplet_table(ctx, 0, m_num_triplets, triplet_table.data()));
        triplet_epoch = tables_epoch.load();
    }

    msm_ctx* ctx = nullptr;
    std::mutex mtx, lazy;
    std::atomic<long> tables_epoch{0};
    std::atomic<long> unary_epoch{-1}, pair_epoch{-1}, triplet_epoch{-1};
    std::vector<double> unary_pristine;
    std::vector<float> pair_table, triplet_table;
    bool target_dirty = true, source_dirty = true, source_xyz_dirty = false, grid_dirty = true, cp_xyz_dirty = false,
         params_dirty = true, labels_dirty = true;

    //---MESHES---//
    newresampler::Mesh _TARGET, _SOURCE, _ORIG, _oCPgrid, _CPgrid;
    //---CF WEIGHTINGS---//
    NEWMAT::Matrix _HIGHREScfweight;
    NEWMAT::ColumnVector MAXSEP;
    std::vector<double> abs_weights;
    bool abs_weights_given = false;   // set_absolute_weights: explicit values instead of resample_weights()
    std::shared_ptr<featurespace> FEAT;
    std::vector<newresampler::Point> _labels;
    std::string dopt;
    int _iter = 0;
    double MVDmax = 0.0, _MEANANGLE = 0.0;
    float _controlptrange = 1.0;
    float _mu = 0.4f, _kappa = 1.6f, _rexp = 2.0f;
    int _simmeasure = 2, _rmode = 1;
    float _k_exp = 2.0f;
    bool _dweight = false, _anorm = false;
    int mcmc_threads = 1;
};

class UnivariateNonLinearSRegDiscreteCostFunction : public NonLinearSRegDiscreteCostFunction {
protected:
    bool is_multivariate() const override { return false; }
};

class MultivariateNonLinearSRegDiscreteCostFunction : public NonLinearSRegDiscreteCostFunction {
protected:
    bool is_multivariate() const override { return true; }
};

}  // namespace newmeshreg
