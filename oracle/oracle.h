// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/geom.h header).  PARITY PINNED AGAINST THE REFERENCE ITSELF: the
// reference has no tests or golden vectors for this path, but its own cost-evaluation sources
// (src/DiscreteCostFunction.cpp, similarities.cc, reg_tools.cpp, DiscreteModel.cpp, DiscreteGroup*.cpp) compile
// unmodified where they lie once the absent FSL libraries are stood in for (oracle/ext/fsl_ext.h ->
// oracle/_ref/libref_cost.so, oracle/Makefile `refcost`).  tests/test_reference_pin.py holds this port to that build
// (unary 1e-12 abs, strain / pairwise 1e-9 rel, patch lists / cliques / label sets bit exact) and
// tests/golden/small_tables.npz is generated from it.  Not pinnable: the [EXT] geometry of oracle/geom.h.
//
// oracle/oracle.h : double-precision, single-thread-semantics restatement of the reference hot path:
//   * model glue          src/DiscreteModel.cpp:38-162,188-294   (MAXSEP, label grid, rotations, pairs, triplets)
//   * patch construction  src/DiscreteCostFunction.cpp:104-109,394-410,458-481
//   * unary data term     src/DiscreteCostFunction.cpp:306-313,412-448,483-526 + src/similarities.cc:122-173
//   * triplet regulariser src/DiscreteCostFunction.cpp:138-201,246-253 + src/reg_tools.cpp:272-318,556-601,703-748
//   * pairwise regulariser src/DiscreteCostFunction.cpp:256-304
//   * group costs         src/DiscreteGroupCostFunction.cpp:9-54
//   * total energy        src/DiscreteCostFunction.cpp:41-70
#pragma once
#include <memory>
#include <string>

#include "geom.h"

namespace orc {

// ---- similarity: src/similarities.h:27-29, src/similarities.cc:122-173 (means kept LOCAL: race-free, Q11) ----
double corr(const std::vector<double>& A, const std::vector<double>& B, const std::vector<double>& W);
inline double get_sim_for_min(const std::vector<double>& in, const std::vector<double>& ref, const std::vector<double>& w) {
    return 1 - (1 + corr(in, ref, w)) / 2;
}

// ---- strain: src/reg_tools.cpp:272-318 (calculate_tri), :556-601 (triangle_strain), :703-748 ----
struct Tangs { Pt e1, e2; };
Tangs calculate_tri(const Pt& a);
double triangle_strain(const double AA[3][3], const double BB[3][3], double MU, double KAPPA, double k_exp);
double calculate_triangular_strain(const Tri& ORIG_tr, const Tri& FINAL_tr, double mu, double kappa, double k_exp = 2.0);
// Basis-free form; asserted equal to the literal route in tests.
double strain_gram(const Tri& ORIG_tr, const Tri& FINAL_tr, double mu, double kappa, double k_exp = 2.0);

struct Params {               // src/DiscreteCostFunction.cpp:121-136 (set_parameters)
    double lambda = 1.0;      // "lambda"  -> _reglambda
    double rexp = 2.0;        // "exponent"
    double mu = 0.4;          // "shearmodulus"
    double kappa = 1.6;       // "bulkmodulus"
    double k_exp = 2.0;       // "kexponent"
    double range = 1.0;       // "range"   -> _controlptrange
    int simmeasure = 2;       // "simmeasure"
    int rmode = 1;            // "regularisermode"
    bool dweight = false;     // "weight"
    bool anorm = false;       // "anorm"
    int threads = 1;          // "numthreads"
    bool bary_orthogonal = false;  // A4 switch (not a reference parameter)
};

// ---- model glue: src/DiscreteModel.cpp ----
struct LabelGrid {
    std::vector<Pt> samples, barycentres;
    Pt centre;
    int centroid = 0;
};
LabelGrid label_sampling_grid(int SGres, double maxs_dist);                              // :91-162
void spacings(const Mesh& CP, std::vector<double>& MAXSEP, double& MVDmax, double& MVD); // :48-67
std::vector<int> estimate_triplets(const Mesh& CP);                                      // :267-282
std::vector<int> estimate_pairs(const Mesh& CP);                                         // :245-265
std::vector<Mat3> get_rotations(const Pt& centre, const std::vector<Pt>& cps);           // :284-294

// ---- mesh warp: [EXT] A8 barycentric_mesh_interpolation + src/reg_tools.cpp:62-181 unfold ----
// up[i] <- normalise( sum_j w_j * low_final[n_j] ) * RAD with (n_j, w_j) = closest triangle of low_init + barycentric weights
void barycentric_mesh_interpolation(std::vector<Pt>& up, const Mesh& low_init, const std::vector<Pt>& low_final);
bool check_for_intersections(int ind, const Mesh& M);   // :121-132
Pt spatialgradient(int index, const Mesh& M);           // :97-119
int unfold(Mesh& SOURCE);                               // :134-181, returns the number of sweeps performed

// ---- cost function: NonLinearSRegDiscreteCostFunction + Univariate/Multivariate variants ----
class CostFunction {
public:
    Params P;
    Mesh TARGET, SOURCE, CPgrid, oCPgrid;
    std::vector<Pt> ORIG;                       // _ORIG coords, indexed by CP id (nested numbering, Q3)
    int C = 1;                                  // FEAT->get_dim()
    std::vector<std::vector<double>> ref;       // FEAT->get_ref_val(d+1, v+1)   [C][Vt]
    std::vector<std::vector<double>> input;     // FEAT->get_input_val(d+1, v+1) [C][Vs]
    std::vector<std::vector<double>> cfweight;  // _HIGHREScfweight rows (0, 1 or C rows)
    std::vector<double> MAXSEP, AbsoluteWeights;
    bool absweights_given = false;   // set_absweights: explicit values instead of resample_weights()
    double MVDmax = 0, MEANANGLE = 0;
    std::vector<Pt> labels;
    std::vector<Mat3> ROT;
    std::vector<int> pairs, triplets;
    std::vector<std::vector<int>> sourceinrange;
    bool multivariate = false;
    SphereLocator targettree;

    void set_meshes(const Mesh& target, const Mesh& source, const Mesh& grid);  // DiscreteCostFunction.h:166-168
    void set_initial_angles();                                                  // DiscreteCostFunction.cpp:95-102
    bool within_controlpt_range(int CPindex, int sourceindex) const;            // :104-109
    void get_source_data();                                                     // :394-410 / :458-481
    double computeUnaryCost(int node, int label) const;                         // :437-448 / :509-526
    void computeUnaryCosts(double* unary) const;                                // :306-313
    double computeTripletCost(int triplet, int la, int lb, int lc) const;       // :138-254
    double computePairwiseCost(int pair, int la, int lb);                       // :256-296 (mutates/restores CPgrid)
    void computePairwiseCosts(double* out);                                     // :298-304
    double evaluateTotalCostSum(const int* labeling, bool use_unary);           // :41-70
    int L() const { return (int)labels.size(); }
    int N() const { return CPgrid.nvertices(); }
};

// ---- group cost function: src/DiscreteGroupCostFunction.cpp ----
struct GroupCost {
    Params P;
    int VERTICES_PER_SUBJ = 0, TRIPLETS_PER_SUBJ = 0;
    std::vector<Mesh> CONTROLMESHES;
    std::vector<Pt> ORIG;
    std::vector<Pt> labels;
    std::vector<Mat3> ROT;      // S*N global
    std::vector<int> triplets;  // global ids
    std::vector<int> pairs;
    std::vector<std::map<int, double>> patch_data;
    double computeTripletCost(int triplet, int la, int lb, int lc) const;  // :9-30
    double computePairwiseCost(int pair, int la, int lb, int L) const;     // :32-54
};

}  // namespace orc
