// DiscreteModel.h — host-side mirror of the reference's MRF model classes (src/DiscreteModel.h:10-68 base,
// :70-179 NonLinearSRegDiscreteModel; bodies follow src/DiscreteModel.cpp).  The model is the CALLER of the hot path and
// stays on the host (north_star): node/label/pair/triplet bookkeeping, the label (sampling) grid, per-CP rotations and
// the per-iteration setupCostFunction() sequence.  It is FSL-free so the unmodified vendored optimisers
// (include/FastPD/FastPD.h, include/Fusion/Fusion.h include "../src/DiscreteModel.h") compile against it.
#pragma once
#include <omp.h>

#include "DiscreteCostFunction.h"
#include "mesh_warp.h"

namespace newmeshreg {

class DiscreteModel {
public:
    virtual ~DiscreteModel() {
        delete[] labeling;
        delete[] pairs;
        delete[] triplets;
    }
    //---GET---//
    int getNumNodes() const { return m_num_nodes; }
    int getNumLabels() const { return m_num_labels; }
    int getNumPairs() const { return m_num_pairs; }
    int getNumTriplets() const { return m_num_triplets; }
    int* getLabeling() { return labeling; }
    const int* getLabeling() const { return labeling; }
    const int* getPairs() const { return pairs; }
    const int* getTriplets() const { return triplets; }
    virtual std::shared_ptr<DiscreteCostFunction> getCostFunction() = 0;
    //---COMPUTE---//
    virtual void computeUnaryCosts() {}
    virtual double computeUnaryCost(int node, int label) { return 0; }
    virtual void computePairwiseCosts() {}
    virtual double computePairwiseCost(int pair, int labelA, int labelB) { return 0; }
    virtual void computeTripletCosts() {}
    virtual double computeTripletCost(int triplet, int labelA, int labelB, int labelC) { return 0; }
    virtual double computeTripletCostTri(int trID, int labelA, int labelB, int labelC) { return 0; }
    virtual double evaluateTotalCostSum() { return 0; }
    //---MODIFY---//
    virtual void applyLabeling(int* discreteLabeling) {}
    virtual void report() {}

protected:
    void resetLabeling() { if (labeling) std::fill(labeling, labeling + m_num_nodes, 0); }
    void initLabeling() {
        if (m_num_nodes != 0) {
            delete[] labeling;
            labeling = new int[m_num_nodes];
            resetLabeling();
        }
    }
    int m_num_nodes = 0, m_num_labels = 0, m_num_pairs = 0, m_num_triplets = 0;
    int* labeling = nullptr;
    int* pairs = nullptr;
    int* triplets = nullptr;
    std::string m_outdir;
    bool m_verbosity = false;
    int _nthreads = 1;
};

class NonLinearSRegDiscreteModel : public DiscreteModel {
public:
    NonLinearSRegDiscreteModel() = default;
    explicit NonLinearSRegDiscreteModel(myparam& PAR) {
        set_parameters(PAR);
        initialize_cost_function(m_multivariate, PAR);
    }

    void set_parameters(myparam& PAR) {  // src/DiscreteModel.cpp:5-20
        auto geti = [&](const char* k) { return boost::get<int>(PAR.find(k)->second); };
        auto getb = [&](const char* k) { return boost::get<bool>(PAR.find(k)->second); };
        auto getf = [&](const char* k) { return boost::get<float>(PAR.find(k)->second); };
        auto gets = [&](const char* k) { return boost::get<std::string>(PAR.find(k)->second); };
        optimiser = gets("dOPT"); m_SGres = geti("SGres"); m_CPres = geti("CPres"); m_regoption = geti("regularisermode");
        m_multivariate = getb("multivariate"); m_verbosity = getb("verbosity"); m_outdir = gets("outdir");
        m_triclique = getb("TriLikelihood"); m_rescalelabels = getb("rescalelabels"); _nthreads = geti("numthreads");
        _labeldist = getf("labeldist"); range = getf("range");
        if (m_regoption == 1) _pairwise = true;
    }

    void initialize_cost_function(bool MV, myparam& P) {  // :22-36
        if (m_triclique) throw MeshregException("HO tri-clique cost functions are outside the accelerated path");
        if (MV) costfct = std::shared_ptr<NonLinearSRegDiscreteCostFunction>(new MultivariateNonLinearSRegDiscreteCostFunction());
        else costfct = std::shared_ptr<NonLinearSRegDiscreteCostFunction>(new UnivariateNonLinearSRegDiscreteCostFunction());
        costfct->set_parameters(P);
    }

    virtual void Initialize(const newresampler::Mesh& CONTROLGRID) {  // :38-89
        m_CPgrid = CONTROLGRID;
        m_num_nodes = m_CPgrid.nvertices();
        m_num_pairs = 0;
        initLabeling();

        NEWMAT::ColumnVector vMAXmvd(m_CPgrid.nvertices());   // largest geodesic separation to a 1-ring neighbour
        vMAXmvd = 0;
#pragma omp parallel for num_threads(_nthreads)
        for (int k = 0; k < m_CPgrid.nvertices(); k++) {
            const newresampler::Point CP = m_CPgrid.get_coord(k);
            for (auto it = m_CPgrid.nbegin(k); it != m_CPgrid.nend(k); it++) {
                const double dist = 2 * RAD * asin((CP - m_CPgrid.get_coord(*it)).norm() / (2 * RAD));
                if (dist > vMAXmvd(k + 1)) vMAXmvd(k + 1) = dist;
            }
        }
        MVD = m_CPgrid.calculate_MeanVD();
        const double MVDmax = m_CPgrid.calculate_MaxVD();
        m_maxs_dist = _labeldist * MVDmax;

        costfct->set_meshes(m_TARGET, m_SOURCE, m_CPgrid);
        costfct->set_initial_angles(m_CPgrid.get_face_angles());
        costfct->set_spacings(vMAXmvd, MVDmax);

        m_iter = 1;
        m_scale = 1.0;
        if (_pairwise) estimate_pairs();
        else estimate_triplets();
        Initialize_sampling_grid();
        get_rotations(m_ROT);
        m_inputtree = std::make_shared<newresampler::Octree>(m_TARGET);
    }

    void Initialize_sampling_grid() {  // :91-103
        m_samplinggrid = newresampler::make_mesh_from_icosa(m_SGres);
        true_rescale(m_samplinggrid, RAD);
        for (int i = 0; i < m_samplinggrid.nvertices(); i++)
            if (m_samplinggrid.get_total_neighbours(i) == 6) { m_centroid = i; break; }
        label_sampling_grid(m_centroid, m_maxs_dist, m_samplinggrid);
    }

    // Breadth-first sweep of the sampling grid around the centroid (:105-162): vertices within `dist` become the
    // "samples" label set; projected barycentres of the faces met on the way, with distinct directions, the other one.
    void label_sampling_grid(int centroid, double dist, newresampler::Mesh& Grid) {
        m_samples.clear(); m_barycentres.clear();
        std::vector<char> seen_v(Grid.nvertices(), 0), seen_f(Grid.ntriangles(), 0);
        centre = Grid.get_coord(centroid);
        m_samples.push_back(centre);
        m_barycentres.push_back(centre);
        std::vector<int> frontier{centroid}, next;
        while (!frontier.empty()) {
            for (int v : frontier) {
                for (auto j = Grid.nbegin(v); j != Grid.nend(v); j++) {
                    if (*j == centroid || seen_v[*j]) continue;
                    if ((Grid.get_coord(*j) - centre).norm() <= dist) {
                        m_samples.push_back(Grid.get_coord(*j));
                        next.push_back(*j);
                        seen_v[*j] = 1;
                    }
                }
                for (auto f = Grid.tIDbegin(v); f != Grid.tIDend(v); f++) {
                    if (seen_f[*f]) continue;
                    const newresampler::Point a = Grid.get_triangle_vertex(*f, 0), b = Grid.get_triangle_vertex(*f, 1), c = Grid.get_triangle_vertex(*f, 2);
                    newresampler::Point bary((a.X + b.X + c.X) / 3, (a.Y + b.Y + c.Y) / 3, (a.Z + b.Z + c.Z) / 3);
                    bary.normalize();
                    bary = bary * RAD;
                    const newresampler::Point d = bary - centre;
                    if (!(d.norm() <= dist && d.norm() > 0)) continue;
                    bool duplicate_direction = false;
                    for (const auto& m : m_barycentres) {  // entry 0 (the centre) yields 0/0: comparison false, as in the reference
                        const newresampler::Point e = m - centre;
                        if (std::abs(1 - ((d | e) / (d.norm() * e.norm()))) < 1e-2) duplicate_direction = true;
                    }
                    if (!duplicate_direction) m_barycentres.push_back(bary);
                    seen_f[*f] = 1;
                }
            }
            frontier.swap(next);
            next.clear();
        }
    }

    std::vector<newresampler::Point> rescale_sampling_grid() {  // :164-186
        std::vector<newresampler::Point> newlabels(m_samples.size());
        if (m_scale >= 0.25) {
            for (size_t i = 0; i < m_samples.size(); i++) {
                newresampler::Point s = centre + (centre - m_samples[i]) * m_scale;
                s.normalize();
                newlabels[i] = s * 100;
            }
        } else {
            m_scale = 1;
            newlabels = m_samples;
        }
        m_scale *= 0.8;
        return newlabels;
    }

    virtual void setupCostFunction() {  // :188-236
        HostTimer t_all("setupCostFunction");
        resetLabeling();
        { HostTimer t("  reset_CPgrid"); costfct->reset_CPgrid(m_CPgrid); }
        if (m_iter == 1) {
            costfct->initialize_regulariser();
            costfct->set_octrees(m_inputtree);
        }
        costfct->reset_anatomical();
        // ROT[k] = rotation(label-grid centre -> CP_k) is only consumed on the host by applyLabeling (:238-243; the cost
        // function recomputes it on the device): snapshot the grid now, evaluate lazily there — it is ~1/3 of this call
        m_rot_grid = m_CPgrid; m_rot_stale = true;
        if (m_rescalelabels) m_labels = rescale_sampling_grid();
        else if (m_iter % 2 == 0) m_labels = m_samples;
        else m_labels = m_barycentres;
        m_num_labels = (int)m_labels.size();
        costfct->set_labels(m_labels, m_ROT);
        { HostTimer t("  initialize"); costfct->initialize(m_num_nodes, m_num_labels, m_num_pairs, m_num_triplets); }
        if (_pairwise) costfct->setPairs(pairs);
        else costfct->setTriplets(triplets);
        { HostTimer t("  get_source_data"); costfct->get_source_data(); }  // patches + the whole unary table on the GPU
        if (optimiser == "MCMC") costfct->set_mcmc_threads(_nthreads);
        m_iter++;
    }

    //---COMPUTE COSTS---//
    void computeUnaryCosts() override { costfct->computeUnaryCosts(); }
    double computeUnaryCost(int node, int label) override { return costfct->computeUnaryCost(node, label); }
    void computePairwiseCosts() override { costfct->computePairwiseCosts(pairs); }
    double computePairwiseCost(int pair, int labelA, int labelB) override { return costfct->computePairwiseCost(pair, labelA, labelB); }
    double computeTripletCost(int triplet, int labelA, int labelB, int labelC) override { return costfct->computeTripletCost(triplet, labelA, labelB, labelC); }
    double evaluateTotalCostSum() override { return costfct->evaluateTotalCostSum(labeling, pairs, triplets); }

    //---SETUP MODEL---//
    virtual void set_meshspace(const newresampler::Mesh& target, const newresampler::Mesh& source, int num = 1) { m_TARGET = target; m_SOURCE = source; }
    // src/DiscreteModel.h:99-105 — called by the driver when anatomical meshes are supplied (src/mesh_registration.cpp:77-78,
    // regulariser modes 4/5).  Those modes are outside the accelerated path: the calls exist so that
    // the driver compiles unchanged, and say so when they are reached.
    void set_anatomical_meshspace(const newresampler::Mesh& ref_sphere, const newresampler::Mesh& ref_anat, const newresampler::Mesh& source_sphere,
                                  const newresampler::Mesh& source_anat) {
        costfct->set_anatomical(ref_sphere, ref_anat, source_sphere, source_anat);
    }
    void set_anatomical_neighbourhood(const std::vector<std::map<int, double>>& weights, const std::vector<std::vector<int>>& neighbourhood) {
        costfct->set_anatomical_neighbourhood(weights, neighbourhood);
    }
    void set_featurespace(const std::shared_ptr<featurespace>& FEATURES) { FEAT = FEATURES; costfct->set_featurespace(FEATURES); }
    void setupCostFunctionWeighting(const NEWMAT::Matrix& Weight) { costfct->set_dataaffintyweighting(Weight); }
    virtual void reset_meshspace(const newresampler::Mesh& source, int num = 0) { m_SOURCE = source; costfct->reset_source(source); }
    virtual void reset_CPgrid(const newresampler::Mesh& grid, int num = 0) { m_CPgrid = grid; }
    virtual void warp_CPgrid(newresampler::Mesh& START, newresampler::Mesh& END, int num = 0) {  // src/DiscreteModel.h:122-125
        newresampler::barycentric_mesh_interpolation(m_CPgrid, START, END, _nthreads);
        unfold(m_CPgrid, m_verbosity);
    }
    void set_debug() { m_debug = true; costfct->debug(); }
    void report() override { if (costfct) costfct->report(); }
    newresampler::Mesh get_TARGET() { return m_TARGET; }
    virtual newresampler::Mesh get_CPgrid(int num = 0) { return m_CPgrid; }
    std::shared_ptr<DiscreteCostFunction> getCostFunction() override { return costfct; }
    std::shared_ptr<NonLinearSRegDiscreteCostFunction> getNonLinearCostFunction() { return costfct; }
    bool is_triclique() const { return m_triclique; }
    const std::vector<newresampler::Point>& get_labels() const { return m_labels; }
    const newresampler::Point& get_centre() const { return centre; }

    using DiscreteModel::applyLabeling;
    virtual void applyLabeling() {  // :238-243
        ensure_rotations();
        for (int i = 0; i < m_CPgrid.nvertices(); i++) m_CPgrid.set_coord(i, m_ROT[i] * m_labels[labeling[i]]);
    }

protected:
    newresampler::Mesh m_TARGET, m_SOURCE, m_CPgrid, m_samplinggrid;
    std::shared_ptr<newresampler::Octree> m_inputtree;
    std::shared_ptr<featurespace> FEAT;
    int m_SGres = 4, m_CPres = 2, m_iter = 0, m_centroid = 0, m_regoption = 2;
    double m_maxs_dist = 0.0, MVD = 0.0, _labeldist = 0.5, range = 1;
    float m_scale = 0.0;
    bool m_multivariate = false, m_debug = false, m_triclique = false, _pairwise = false, m_rescalelabels = false;
    std::string optimiser;
    newresampler::Point centre;
    std::vector<newresampler::Point> m_samples, m_barycentres, m_labels;
    std::vector<NEWMAT::Matrix> m_ROT;
    newresampler::Mesh m_rot_grid;   // control grid as of the last setupCostFunction (what m_ROT refers to)
    bool m_rot_stale = false;
    void ensure_rotations() {
        if (!m_rot_stale) return;
        newresampler::Mesh current = m_CPgrid;
        m_CPgrid = m_rot_grid;
        get_rotations(m_ROT);
        m_CPgrid = current;
        m_rot_stale = false;
    }
    std::shared_ptr<NonLinearSRegDiscreteCostFunction> costfct;

    virtual void estimate_pairs() {  // :245-265 edges i<j in vertex / neighbour order
        std::vector<int> p;
        for (int i = 0; i < m_CPgrid.nvertices(); i++)
            for (auto j = m_CPgrid.nbegin(i); j != m_CPgrid.nend(i); j++)
                if (*j > i) { p.push_back(i); p.push_back(*j); }
        m_num_pairs = (int)p.size() / 2;
        delete[] pairs;
        pairs = new int[p.size()];
        std::copy(p.begin(), p.end(), pairs);
    }
    virtual void estimate_triplets() {  // :267-282 faces with ids sorted ascending
        m_num_triplets = m_CPgrid.ntriangles();
        delete[] triplets;
        triplets = new int[(size_t)m_num_triplets * 3];
        for (int i = 0; i < m_CPgrid.ntriangles(); i++) {
            int ids[3] = {m_CPgrid.get_triangle_vertexID(i, 0), m_CPgrid.get_triangle_vertexID(i, 1), m_CPgrid.get_triangle_vertexID(i, 2)};
            std::sort(ids, ids + 3);
            std::copy(ids, ids + 3, triplets + 3 * i);
        }
    }
    virtual void get_rotations(std::vector<NEWMAT::Matrix>& ROT) {  // :284-294
        ROT.assign(m_CPgrid.nvertices(), NEWMAT::Matrix());
        const newresampler::Point ci = m_samplinggrid.get_coord(m_centroid);
#pragma omp parallel for num_threads(_nthreads)
        for (int k = 0; k < m_CPgrid.nvertices(); k++) ROT[k] = newresampler::estimate_rotation_matrix(ci, m_CPgrid.get_coord(k));
    }
};

}  // namespace newmeshreg
