// DiscreteGroupModel.h — host-side mirror of src/DiscreteGroupModel.{h,cpp}: one MRF over S subjects x N control
// points.  Bookkeeping (inter-subject pairs, per-subject triplets, rotations, applyLabeling) follows the reference.  The
// producers of the pairwise data term, get_rotated_meshes / get_patch_data (src/DiscreteGroupModel.cpp:72-117; SURVEY.md §8f
// rank 3), keep the reference's semantics but not its O(S L N V) scans: every "within range" test runs through the patch
// construction kernels (points_in_range), every resampling through GPU point location (metric_resample, [EXT] A8), and
// the maps reach the cost function as CSR without ever becoming std::map.
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
        if (FEAT && FEAT->subjects() == m_num_subjects) {   // data term (a model driven without features keeps the regulariser only)
            get_rotated_meshes();
            if (shard_s0 == 0 && shard_s1 >= m_num_subjects) get_patch_data();   // sharded: after the exchange (finish_setup)
        }
        costfct->setPairs(pairs);
        costfct->setTriplets(triplets);
        m_iter++;
    }
    // src/DiscreteGroupModel.cpp:72-95.  Per subject: every data vertex takes the rotation of the LAST control point (highest
    // index) whose range it is in — what the reference's nested loops leave behind (`rotated_mesh` is declared outside the
    // label loop, so a vertex in no range keeps its position from the previous label); `SP * R` is row vector times matrix.
    // Then metric_resample onto the template.  Only the first data row is kept: get_patch_data reads row 1 (:109).
    // Multi-GPU (SURVEY.md §8e groupwise row): one process per GPU, subjects [s0, s1) produced here, the S L V_t resampled
    // values exchanged by the caller (one all-gather per outer iteration), then finish_setup() on every rank.
    void set_subject_shard(int s0, int s1) { shard_s0 = std::max(0, s0); shard_s1 = s1; }
    void set_rotated_values(int subject, const double* in) {
        const int Vt = m_template.nvertices();
        for (int l = 0; l < m_num_labels; l++) rotated_values[(size_t)subject * m_num_labels + l].assign(in + (size_t)l * Vt, in + (size_t)(l + 1) * Vt);
    }
    void finish_setup() { get_patch_data(); }
    void get_rotated_meshes() {
        const int L = m_num_labels, N = control_grid_size;
        rotated_values.assign((size_t)m_num_subjects * L, std::vector<double>());
        for (int s = std::max(0, shard_s0); s < std::min(m_num_subjects, shard_s1); s++) {
            const newresampler::Mesh& data = m_datameshes[s];
            const int V = data.nvertices();
            std::vector<double> sp(N);
            for (int v = 0; v < N; v++) sp[v] = spacings[s](v + 1);
            std::vector<int> off, idx;
            newresampler::points_in_range(data.coord_data(), V, true, m_controlmeshes[s].coord_data(), N, sp.data(), range, off, idx);
            std::vector<int> owner(V, -1);
            for (int v = 0; v < N; v++)
                for (int e = off[v]; e < off[v + 1]; e++) owner[idx[e]] = v;      // ascending v: the last one in range wins
            newresampler::Mesh rotated = data;
            const std::vector<double>& values = FEAT->get_data_matrix(s);
            for (int l = 0; l < L; l++) {
                std::vector<NEWMAT::Matrix> R(N);
#pragma omp parallel for num_threads(_nthreads)
                for (int v = 0; v < N; v++)
                    R[v] = newresampler::estimate_rotation_matrix(m_controlmeshes[s].get_coord(v), m_ROT[s * N + v] * m_labels[l]);
                for (int i = 0; i < V; i++) {
                    if (owner[i] < 0) continue;
                    const newresampler::Point& p = data.get_coord(i);
                    const NEWMAT::Matrix& M = R[owner[i]];
                    rotated.set_coord(i, newresampler::Point(p.X * M(1, 1) + p.Y * M(2, 1) + p.Z * M(3, 1), p.X * M(1, 2) + p.Y * M(2, 2) + p.Z * M(3, 2),
                                                             p.X * M(1, 3) + p.Y * M(2, 3) + p.Z * M(3, 3)));
                }
                rotated_values[(size_t)s * L + l] = newresampler::metric_resample(rotated, values.data(), 1, m_template);
            }
        }
    }
    // :97-117.  Map (subject, node, label) = template vertices within range of the DISPLACED control point -> resampled value;
    // assembled directly as the CSR the cost function uploads (keys ascending = std::map order).
    void get_patch_data() {
        const int L = m_num_labels, N = control_grid_size, Vt = m_template.nvertices();
        const size_t nmaps = (size_t)m_num_subjects * N * L;
        std::vector<std::vector<int>> offs((size_t)m_num_subjects * L), idxs((size_t)m_num_subjects * L);
        std::vector<long long> moff(nmaps + 1, 0);
        for (int s = 0; s < m_num_subjects; s++) {
            std::vector<double> sp(N);
            for (int v = 0; v < N; v++) sp[v] = spacings[s](v + 1);
            for (int l = 0; l < L; l++) {
                std::vector<double> cps((size_t)N * 3);
                for (int v = 0; v < N; v++) {
                    const newresampler::Point c = m_ROT[s * N + v] * m_labels[l];
                    cps[3 * v] = c.X; cps[3 * v + 1] = c.Y; cps[3 * v + 2] = c.Z;
                }
                auto& off = offs[(size_t)s * L + l]; auto& idx = idxs[(size_t)s * L + l];
                newresampler::points_in_range(m_template.coord_data(), Vt, s == 0 && l == 0, cps.data(), N, sp.data(), range, off, idx);
                for (int v = 0; v < N; v++) moff[((size_t)s * N + v) * L + l + 1] = off[v + 1] - off[v];
            }
        }
        for (size_t m = 0; m < nmaps; m++) moff[m + 1] += moff[m];
        std::vector<int> keys((size_t)moff[nmaps]);
        std::vector<double> vals((size_t)moff[nmaps]);
        for (int s = 0; s < m_num_subjects; s++)
            for (int l = 0; l < L; l++) {
                const auto& off = offs[(size_t)s * L + l]; const auto& idx = idxs[(size_t)s * L + l];
                const std::vector<double>& rv = rotated_values[(size_t)s * L + l];
                for (int v = 0; v < N; v++) {
                    long long o = moff[((size_t)s * N + v) * L + l];
                    for (int e = off[v]; e < off[v + 1]; e++, o++) { keys[o] = idx[e]; vals[o] = rv[idx[e]]; }
                }
            }
        groupcf->set_patch_data_csr(moff.data(), (long long)nmaps, keys.data(), vals.data());
    }
    const std::vector<double>& rotated_mesh_values(int subject, int label) const { return rotated_values[(size_t)subject * m_num_labels + label]; }
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
        // Octree::get_closest_vertex_ID ([EXT] A3): the triangle of B's control mesh under the point, then its nearest corner
        const int N = control_grid_size;
        std::vector<std::vector<int>> near((size_t)m_num_subjects * m_num_subjects);
        for (int b = 1; b < m_num_subjects; b++) {
            std::vector<double> q; std::vector<int> who;
            for (int a = 0; a < b; a++) { auto c = m_controlmeshes[a].coords(); q.insert(q.end(), c.begin(), c.end()); who.push_back(a); }
            std::vector<int> tri; std::vector<double> w;
            newresampler::locate_in_mesh(q.data(), (int)q.size() / 3, m_controlmeshes[b], tri, w);
            for (size_t ai = 0; ai < who.size(); ai++) {
                auto& out = near[(size_t)who[ai] * m_num_subjects + b];
                out.resize(N);
                for (int v = 0; v < N; v++) {
                    const size_t i = ai * N + v;
                    const newresampler::Point p(q[3 * i], q[3 * i + 1], q[3 * i + 2]);
                    int best = 0; double bd = 1e300;
                    for (int j = 0; j < 3; j++) {
                        const double d = (m_controlmeshes[b].get_triangle_vertex(tri[i], j) - p).norm();
                        if (d < bd) { bd = d; best = j; }
                    }
                    out[v] = m_controlmeshes[b].get_triangle_vertexID(tri[i], best);
                }
            }
        }
        int pair = 0;
        for (int a = 0; a < m_num_subjects; a++)
            for (int v = 0; v < N; v++)
                for (int b = a + 1; b < m_num_subjects; b++) {
                    pairs[2 * pair] = a * N + v;
                    pairs[2 * pair + 1] = b * N + near[(size_t)a * m_num_subjects + b][v];
                    pair++;
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
    std::vector<std::vector<double>> rotated_values;   // [subject * L + label][template vertex]: rotated_meshes, first data row
    int m_num_subjects = 0, control_grid_size = 0;
    int shard_s0 = 0, shard_s1 = 1 << 30;
};

}  // namespace newmeshreg
