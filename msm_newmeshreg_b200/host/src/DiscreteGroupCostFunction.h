// DiscreteGroupCostFunction.h — host-side mirror of src/DiscreteGroupCostFunction.{h,cpp}: S subjects share one MRF.
// Triplet term: strain of each subject's own control triangle (fold -> 1e7, k_exp fixed at 2, .cpp:9-30); pairwise
// term: correlation of two (template vertex -> value) patch maps (.cpp:32-54).  Both are evaluated on the GPU; the
// per-element virtuals are look-ups / batched queries behind the same names.
#pragma once
#include "DiscreteCostFunction.h"

namespace newmeshreg {

class DiscreteGroupCostFunction : public NonLinearSRegDiscreteCostFunction {
public:
    void set_meshes(const newresampler::Mesh& target, const newresampler::Mesh& source, const newresampler::Mesh& GRID, int num = 1) override {
        _TEMPLATE = target;
        _ORIG = source;
        VERTICES_PER_SUBJ = GRID.nvertices();
        TRIPLETS_PER_SUBJ = GRID.ntriangles();
        _CONTROLMESHES.assign(num, GRID);
        _oCPgrid = GRID;
        grid_dirty = true;
    }
    void initialize(int numNodes, int numLabels, int numPairs, int numTriplets) override {  // .cpp:5-7
        DiscreteCostFunction::initialize(numNodes, numLabels, numPairs, numTriplets);
        tables_epoch++;
    }
    void reset_CPgrid(const newresampler::Mesh& grid, int num = 0) override { _CONTROLMESHES[num] = grid; cp_xyz_dirty = true; }

    void set_patch_data(const std::vector<std::map<int, double>>& patches) override {  // .h:39-41
        moff.assign(patches.size() + 1, 0);
        mkey.clear(); mval.clear();
        for (size_t m = 0; m < patches.size(); m++) {
            for (const auto& e : patches[m]) { mkey.push_back(e.first); mval.push_back(e.second); }
            moff[m + 1] = (long long)mkey.size();
        }
        patches_dirty = true;
        pair_epoch = -1;
    }
    // CSR variant for producers that never materialise std::map (keys ascending within a map)
    void set_patch_data_csr(const long long* off, long long nmaps, const int* keys, const double* vals) {
        moff.assign(off, off + nmaps + 1);
        mkey.assign(keys, keys + off[nmaps]);
        mval.assign(vals, vals + off[nmaps]);
        patches_dirty = true;
        pair_epoch = -1;
    }

    void get_source_data() override {}                       // no unary term (src/DiscreteCostFunction.h:34)
    void computeUnaryCosts() override {}
    double computeUnaryCost(int, int) override { return 0; }

    double computeTripletCost(int triplet, int labelA, int labelB, int labelC) override {
        ensure_triplet_table();
        const size_t L = m_num_labels;
        return (double)triplet_table[(((size_t)triplet * L + labelA) * L + labelB) * L + labelC];
    }
    // The per-element virtual of the unmodified Fusion::optimize (4 P calls per proposal label from an OpenMP loop,
    // include/Fusion/Fusion.h:163-172) is a look-up in the [P][L][L] table of the current outer iteration, built on the GPU at
    // first use; table-aware drivers use the batched entry points below instead and never build it.
    double computePairwiseCost(int pair, int labelA, int labelB) override {
        ensure_group_pair_table();
        return (double)pair_table[((size_t)pair * m_num_labels + labelB) * m_num_labels + labelA];
    }
    void computePairwiseCosts(const int*) override {}
    // batched paths for a table-aware fusion driver
    void computePairwiseQueries(int n, const int* q, double* out) { std::lock_guard<std::mutex> lock(lazy); push_patches(); check(msm_group_pair_costs(ctx, n, q, out)); }
    void computePairwiseMoves(const int* labeling, int alpha, double* out) { std::lock_guard<std::mutex> lock(lazy); push_patches(); check(msm_group_pair_moves(ctx, labeling, alpha, out)); }
    void computeTripletMovesGroup(const int* labeling, int alpha, float* out) { std::lock_guard<std::mutex> lock(lazy); sync_group(); check(msm_triplet_moves_f32(ctx, labeling, alpha, out)); }
    // evaluateTotalCostSum (src/DiscreteCostFunction.cpp:41-70) on the device: pairwise data term + triplet term at a labelling
    double totalCostBatchedGroup(const int* labeling) {
        std::lock_guard<std::mutex> lock(lazy);
        push_patches();
        double sums[3] = {0, 0, 0};
        check(msm_total_cost(ctx, labeling, sums));
        return sums[0] + sums[1] + sums[2];
    }

protected:
    bool is_group() const override { return true; }
    const newresampler::Point& orig_coord(int k) const override { return _ORIG.get_coord(k % VERTICES_PER_SUBJ); }
    std::vector<double> current_cp_coords() const override {
        std::vector<double> xyz;
        xyz.reserve((size_t)_CONTROLMESHES.size() * VERTICES_PER_SUBJ * 3);
        for (const auto& m : _CONTROLMESHES) { auto c = m.coords(); xyz.insert(xyz.end(), c.begin(), c.end()); }
        return xyz;
    }
    void configure_group() override { check(msm_set_group(ctx, (int)_CONTROLMESHES.size(), VERTICES_PER_SUBJ, TRIPLETS_PER_SUBJ)); }

    // group variant of the device sync: no target / source / features, the control grid is the concatenation of the
    // subjects' grids
    void sync_group() {
        std::lock_guard<std::mutex> lock(mtx);
        if (grid_dirty) {
            auto xyz = current_cp_coords();
            const int Ntot = (int)_CONTROLMESHES.size() * VERTICES_PER_SUBJ;
            std::vector<double> orig((size_t)Ntot * 3), ms(Ntot, 1.0);
            for (int k = 0; k < Ntot; k++) { const auto& p = orig_coord(k); orig[3 * k] = p.X; orig[3 * k + 1] = p.Y; orig[3 * k + 2] = p.Z; }
            check(msm_set_cpgrid(ctx, xyz.data(), Ntot, nullptr, 0, ms.data(), orig.data(), nullptr));
            configure_group();
            grid_dirty = cp_xyz_dirty = false;
            labels_dirty = true;
        }
        if (cp_xyz_dirty) { auto xyz = current_cp_coords(); check(msm_reset_cpgrid(ctx, xyz.data())); cp_xyz_dirty = false; }
        if (params_dirty) {
            msm_params p; msm_default_params(&p);
            p.lambda = _reglambda; p.rexp = _rexp; p.mu = _mu; p.kappa = _kappa; p.k_exp = _k_exp; p.rmode = 3; p.group = 1;
            check(msm_set_params(ctx, &p));
            params_dirty = false;
        }
        if (labels_dirty && !_labels.empty()) {
            std::vector<double> lab(_labels.size() * 3);
            for (size_t i = 0; i < _labels.size(); i++) { lab[3 * i] = _labels[i].X; lab[3 * i + 1] = _labels[i].Y; lab[3 * i + 2] = _labels[i].Z; }
            check(msm_set_labels(ctx, lab.data(), (int)_labels.size(), lab.data()));
            labels_dirty = false;
            patches_dirty = patches_dirty || !moff.empty();
        }
        if (pairs_dirty && _pairs && m_num_pairs > 0) { check(msm_set_pairs(ctx, _pairs, m_num_pairs)); pairs_dirty = false; }
        if (triplets_dirty && _triplets && m_num_triplets > 0) { check(msm_set_triplets(ctx, _triplets, m_num_triplets)); triplets_dirty = false; }
    }
    void push_patches() {
        sync_group();
        std::lock_guard<std::mutex> lock(mtx);
        if (patches_dirty && !moff.empty()) {
            check(msm_set_group_patches(ctx, (long long)moff.size() - 1, moff.data(), mkey.data(), mval.data()));
            patches_dirty = false;
        }
    }
    void ensure_group_pair_table() {
        if (pair_epoch == tables_epoch.load()) return;
        std::lock_guard<std::mutex> lock(lazy);
        if (pair_epoch == tables_epoch.load()) return;
        push_patches();
        pair_table.assign((size_t)m_num_pairs * m_num_labels * m_num_labels, 0.f);
        if (m_num_pairs > 0 && !moff.empty()) check(msm_group_pair_table(ctx, pair_table.data()));
        pair_epoch = tables_epoch.load();
    }
    void ensure_triplet_table() {
        if (triplet_epoch == tables_epoch.load()) return;
        std::lock_guard<std::mutex> lock(lazy);
        if (triplet_epoch == tables_epoch.load()) return;
        sync_group();
        const size_t L = m_num_labels;
        triplet_table.assign((size_t)m_num_triplets * L * L * L, 0.f);
        if (m_num_triplets > 0) check(msm_triplet_table(ctx, 0, m_num_triplets, triplet_table.data()));
        triplet_epoch = tables_epoch.load();
    }

private:
    std::vector<newresampler::Mesh> _CONTROLMESHES;
    newresampler::Mesh _TEMPLATE;
    std::vector<long long> moff;
    std::vector<int> mkey;
    std::vector<double> mval;
    bool patches_dirty = false;
    int TRIPLETS_PER_SUBJ = 0, VERTICES_PER_SUBJ = 0;
};

}  // namespace newmeshreg
