// DiscreteGroupModel.h — host-side mirror of src/DiscreteGroupModel.{h,cpp}: one MRF over S subjects x N control
// points.  Bookkeeping (inter-subject pairs, per-subject triplets, rotations, applyLabeling) follows the reference; the
// producers of the pairwise data term, get_rotated_meshes / get_patch_data (src/DiscreteGroupModel.cpp:72-117), depend on
// FSL's metric_resample ([EXT] A8) and are a "next" row of the scope table: patch maps are handed
// in through set_patch_data() by the caller.
#pragma once
#include "DiscreteGroupCostFunction.h"
#include "DiscreteModel.h"

namespace newmeshreg {

class DiscreteGroupModel : public NonLinearSRegDiscreteModel {
public:
    DiscreteGroupModel() = default;
    explicit DiscreteGroupModel(myparam& p) {
        set_parameters(p);
        groupcf = std::make_shared<DiscreteGroupCostFunction>();
        costfct = groupcf;
        costfct->set_parameters(p);
    }

    void set_meshspace(const newresampler::Mesh& target, const newresampler::Mesh& source, int num = 1) override {
        m_template = target;
        m_datameshes.assign(num, source);
        m_num_subjects = num;
    }
    void reset_meshspace(const newresampler::Mesh& source, int num = 0) override { m_datameshes[num] = source; }
    void reset_CPgrid(const newresampler::Mesh& grid, int num = 0) override { m_controlmeshes[num] = grid; }
    newresampler::Mesh get_CPgrid(int num = 0) override { return m_controlmeshes[num]; }

    void applyLabeling() override { applyLabeling(labeling); }
    void applyLabeling(int* dlabels) override {  // src/DiscreteGroupModel.h:49-55
#pragma omp parallel for num_threads(_nthreads)
        for (int s = 0; s < m_num_subjects; s++)
            for (int v = 0; v < control_grid_size; v++)
                m_controlmeshes[s].set_coord(v, m_ROT[s * control_grid_size + v] * m_labels[dlabels[s * control_grid_size + v]]);
    }

    void Initialize(const newresampler::Mesh& controlgrid) override {  // src/DiscreteGroupModel.cpp:119-162
        m_controlmeshes.assign(m_num_subjects, controlgrid);
        control_grid_size = controlgrid.nvertices();
        m_num_nodes = control_grid_size * m_num_subjects;
        initLabeling();
        spacings.assign(m_num_subjects, NEWMAT::ColumnVector());
        for (int s = 0; s < m_num_subjects; s++) {
            NEWMAT::ColumnVector v(control_grid_size);
            v = 0;
            for (int k = 0; k < control_grid_size; k++) {
                const newresampler::Point CP = m_controlmeshes[s].get_coord(k);
                for (auto it = m_controlmeshes[s].nbegin(k); it != m_controlmeshes[s].nend(k); it++) {
                    const double d = 2 * RAD * asin((CP - m_controlmeshes[s].get_coord(*it)).norm() / (2 * RAD));
                    if (d > v(k + 1)) v(k + 1) = d;
                }
            }
            spacings[s] = v;
        }
        MVD = controlgrid.calculate_MeanVD();
        m_maxs_dist = _labeldist * controlgrid.calculate_MaxVD();
        costfct->set_meshes(m_template, m_datameshes[0], controlgrid, m_num_subjects);
        m_iter = 1;
        initialize_pairs();
        estimate_pairs();
        estimate_triplets();
        Initialize_sampling_grid();
        get_rotations(m_ROT);
    }

    void setupCostFunction() override {  // src/DiscreteGroupModel.cpp:164-201 (data-term producers: see header comment)
        resetLabeling();
        for (int s = 0; s < m_num_subjects; s++) costfct->reset_CPgrid(m_controlmeshes[s], s);
        get_rotations(m_ROT);
        costfct->set_iter(m_iter);
        m_labels = (m_iter % 2 == 0) ? m_samples : m_barycentres;
        m_num_labels = (int)m_labels.size();
        costfct->set_labels(m_labels, m_ROT);
        costfct->initialize(m_num_nodes, m_num_labels, m_num_pairs, m_num_triplets);
        estimate_pairs();
        costfct->setPairs(pairs);
        costfct->setTriplets(triplets);
        m_iter++;
    }
    std::shared_ptr<DiscreteGroupCostFunction> getGroupCostFunction() { return groupcf; }
    int subjects() const { return m_num_subjects; }
    int nodes_per_subject() const { return control_grid_size; }
    const NEWMAT::ColumnVector& subject_spacings(int s) const { return spacings[s]; }

protected:
    void initialize_pairs() {  // :5-14
        m_num_pairs = control_grid_size * m_num_subjects * (m_num_subjects - 1) / 2;
        delete[] pairs;
        pairs = new int[2 * (size_t)std::max(m_num_pairs, 1)];
    }
    void estimate_pairs() override {  // :16-34 : node of subject A <-> closest control point of every later subject B
        int pair = 0;
        for (int a = 0; a < m_num_subjects; a++)
            for (int v = 0; v < control_grid_size; v++) {
                const newresampler::Point CP = m_controlmeshes[a].get_coord(v);
                for (int b = a + 1; b < m_num_subjects; b++) {
                    int best = 0; double bd = 1e300;   // Octree::get_closest_vertex_ID ([EXT] A3): nearest vertex
                    for (int u = 0; u < control_grid_size; u++) {
                        const double d = (m_controlmeshes[b].get_coord(u) - CP).norm();
                        if (d < bd) { bd = d; best = u; }
                    }
                    pairs[2 * pair] = a * control_grid_size + v;
                    pairs[2 * pair + 1] = b * control_grid_size + best;
                    pair++;
                }
            }
    }
    void estimate_triplets() override {  // :36-58
        const int nt = m_controlmeshes[0].ntriangles();
        m_num_triplets = m_num_subjects * nt;
        delete[] triplets;
        triplets = new int[3 * (size_t)m_num_triplets];
        for (int s = 0; s < m_num_subjects; s++)
            for (int t = 0; t < nt; t++) {
                int ids[3];
                for (int j = 0; j < 3; j++) ids[j] = m_controlmeshes[s].get_triangle_vertexID(t, j) + s * control_grid_size;
                std::sort(ids, ids + 3);
                std::copy(ids, ids + 3, triplets + ((size_t)s * nt + t) * 3);
            }
    }
    void get_rotations(std::vector<NEWMAT::Matrix>& ROT) override {  // :60-70
        ROT.assign((size_t)m_num_subjects * control_grid_size, NEWMAT::Matrix());
        const newresampler::Point ci = m_samplinggrid.get_coord(m_centroid);
#pragma omp parallel for num_threads(_nthreads)
        for (int s = 0; s < m_num_subjects; s++)
            for (int v = 0; v < control_grid_size; v++)
                ROT[s * control_grid_size + v] = newresampler::estimate_rotation_matrix(ci, m_controlmeshes[s].get_coord(v));
    }

    std::vector<newresampler::Mesh> m_datameshes, m_controlmeshes;
    newresampler::Mesh m_template;
    std::vector<NEWMAT::ColumnVector> spacings;
    std::shared_ptr<DiscreteGroupCostFunction> groupcf;
    int m_num_subjects = 0, control_grid_size = 0;
};

}  // namespace newmeshreg
