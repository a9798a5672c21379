// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).  Pinned against the reference's own sources compiled in place
// (oracle/_ref/libref_cost.so, tests/test_reference_pin.py).
#include "oracle.h"

#include <limits>

#include <omp.h>

#include <cstring>

namespace orc {

// =========================================================================================================
// similarity — src/similarities.cc:122-173.  `initialize` (weighted means) then `corr`.
// =========================================================================================================
double corr(const std::vector<double>& A, const std::vector<double>& B, const std::vector<double>& weights) {
    double meanA = 0.0, meanB = 0.0, sum = 0.0;
    for (size_t i = 0; i < A.size(); i++) {  // :128-137
        double varwght = 1.0;
        if (weights.size() == A.size()) varwght = weights[i];
        meanA += varwght * A[i];
        meanB += varwght * B[i];
        sum += varwght;
    }
    if (sum > 0.0) { meanA /= sum; meanB /= sum; }  // :139-142

    double prod = 0.0, varA = 0.0, varB = 0.0;
    sum = 0.0;
    for (size_t s = 0; s < A.size(); s++) {  // :152-162
        double varwght = 1.0;
        if (weights.size() == A.size()) varwght = weights[s];
        prod += varwght * (A[s] - meanA) * (B[s] - meanB);
        varA += varwght * (A[s] - meanA) * (A[s] - meanA);
        varB += varwght * (B[s] - meanB) * (B[s] - meanB);
        sum += varwght;
    }
    if (sum > 0) { prod /= sum; varA /= sum; varB /= sum; }  // :164-169
    if (varA == 0.0 || varB == 0.0) return 0.0;               // :171
    return prod / (std::sqrt(varA) * std::sqrt(varB));        // :172
}

// =========================================================================================================
// strain — src/reg_tools.cpp
// =========================================================================================================
Tangs calculate_tri(const Pt& a) {  // :272-318
    Tangs T;
    Pt b(1.0, 0.0, 0.0);
    Pt c = a * b;
    double len = c.X * c.X + c.Y * c.Y + c.Z * c.Z;
    if (len == 0.0) {
        b = Pt(0.0, 1.0, 0.0);
        c = a * b;
        len = c.X * c.X + c.Y * c.Y + c.Z * c.Z;
    }
    len = std::sqrt(len);
    if (len == 0.0) len = 1;
    T.e1 = Pt(c.X / len, c.Y / len, c.Z / len);
    b = a * c;
    len = std::sqrt(b.X * b.X + b.Y * b.Y + b.Z * b.Z);
    if (len == 0) len = 1;
    T.e2 = Pt(b.X / len, b.Y / len, b.Z / len);
    return T;
}

double triangle_strain(const double AA[3][3], const double BB[3][3], double MU, double KAPPA, double k_exp) {  // :556-601
    // all t0 distances (AA(i,j) is 1-based in the reference; rows = vertices, cols = 2-D coords)
    double c0 = AA[1][0] - AA[0][0], c1 = AA[1][1] - AA[0][1];
    double c4 = AA[2][0] - AA[0][0], c5 = AA[2][1] - AA[0][1];
    double c0c = BB[1][0] - BB[0][0], c1c = BB[1][1] - BB[0][1];
    double c4c = BB[2][0] - BB[0][0], c5c = BB[2][1] - BB[0][1];
    // Edges = [c0 c4; c1 c5], edges likewise (:580-581); F = edges * Edges.i() (:583)
    double det = c0 * c5 - c4 * c1;
    double Ei[2][2] = {{c5 / det, -c4 / det}, {-c1 / det, c0 / det}};
    double e[2][2] = {{c0c, c4c}, {c1c, c5c}};
    double F[2][2];
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) F[i][j] = e[i][0] * Ei[0][j] + e[i][1] * Ei[1][j];
    Mat3 F3D;  // :585-586 (third row/col = e3)
    F3D.m[0][0] = F[0][0]; F3D.m[0][1] = F[0][1]; F3D.m[0][2] = 0;
    F3D.m[1][0] = F[1][0]; F3D.m[1][1] = F[1][1]; F3D.m[1][2] = 0;
    F3D.m[2][0] = 0; F3D.m[2][1] = 0; F3D.m[2][2] = 1;
    Mat3 F3D_2 = F3D.t() * F3D;  // :588
    double I1 = F3D_2.trace();   // :589
    double I3 = F3D_2.det();     // :590
    double J = std::sqrt(I3);
    double I1st_new = (I1 - 1.0) / J;
    double R;
    if (I1st_new <= 2) R = 1.0;
    else R = 0.5 * (I1st_new + std::sqrt(I1st_new * I1st_new - 4));
    double Rshared = std::pow(R, k_exp), Jshared = std::pow(J, k_exp);
    return 0.5 * (MU * (Rshared + 1.0 / Rshared - 2) + KAPPA * (Jshared + 1.0 / Jshared - 2));  // :601
}

double calculate_triangular_strain(const Tri& ORIG_tr, const Tri& FINAL_tr, double mu, double kappa, double k_exp) {  // :703-748
    Pt Normal_O = ORIG_tr.normal(), Normal_F = FINAL_tr.normal();
    Tangs T = calculate_tri(Normal_O);
    Tangs T_trans = calculate_tri(Normal_F);
    Mat3 TRANS = form_matrix_from_points(T.e1, T.e2, Normal_O);
    if (TRANS.det() < 0) {  // :718-724 swap first two columns
        for (int r = 0; r < 3; r++) std::swap(TRANS.m[r][0], TRANS.m[r][1]);
    }
    Mat3 TRANS2 = form_matrix_from_points(T_trans.e1, T_trans.e2, Normal_F);
    if (TRANS.det() < 0) {  // :726 (Q9: tests TRANS, not TRANS2 — reproduced literally; never true here)
        for (int r = 0; r < 3; r++) std::swap(TRANS2.m[r][0], TRANS2.m[r][1]);
    }
    double ORIG2D[3][3], FINAL2D[3][3];
    for (int i = 0; i < 3; i++) {  // :734-744 rows = vertices; X2D = X3D * TRANS
        const Pt& vo = ORIG_tr.v[i];
        const Pt& vf = FINAL_tr.v[i];
        for (int j = 0; j < 3; j++) {
            ORIG2D[i][j] = vo.X * TRANS.m[0][j] + vo.Y * TRANS.m[1][j] + vo.Z * TRANS.m[2][j];
            FINAL2D[i][j] = vf.X * TRANS2.m[0][j] + vf.Y * TRANS2.m[1][j] + vf.Z * TRANS2.m[2][j];
        }
    }
    return triangle_strain(ORIG2D, FINAL2D, mu, kappa, k_exp);
}

double strain_gram(const Tri& T0, const Tri& T1, double mu, double kappa, double k_exp) {
    Pt E1 = T0.v[1] - T0.v[0], E2 = T0.v[2] - T0.v[0];
    Pt e1 = T1.v[1] - T1.v[0], e2 = T1.v[2] - T1.v[0];
    double G11 = E1 | E1, G12 = E1 | E2, G22 = E2 | E2;
    double g11 = e1 | e1, g12 = e1 | e2, g22 = e2 | e2;
    double det0 = G11 * G22 - G12 * G12, det1 = g11 * g22 - g12 * g12;
    // tr(G1 * G0^-1), G0^-1 = [G22 -G12; -G12 G11]/det0
    double trC = (g11 * G22 - 2 * g12 * G12 + g22 * G11) / det0;
    double I3 = det1 / det0;
    double J = std::sqrt(I3);
    double I1s = trC / J;
    double R = (I1s <= 2) ? 1.0 : 0.5 * (I1s + std::sqrt(I1s * I1s - 4));
    double Rs = std::pow(R, k_exp), Js = std::pow(J, k_exp);
    return 0.5 * (mu * (Rs + 1.0 / Rs - 2) + kappa * (Js + 1.0 / Js - 2));
}

// =========================================================================================================
// model glue — src/DiscreteModel.cpp
// =========================================================================================================
LabelGrid label_sampling_grid(int SGres, double dist) {
    LabelGrid LG;
    Mesh Grid = make_mesh_from_icosa(SGres, RAD);  // :93-94
    for (int i = 0; i < Grid.nvertices(); i++)     // :96-101 first vertex with 6 neighbours
        if (Grid.get_total_neighbours(i) == 6) { LG.centroid = i; break; }
    const int centroid = LG.centroid;
    std::vector<int> getneighbours, newneighbours;
    std::vector<char> found(Grid.nvertices(), 0), found_tr(Grid.ntriangles(), 0);
    Pt centre = Grid.get_coord(centroid);
    LG.centre = centre;
    getneighbours.push_back(centroid);
    LG.samples.push_back(centre);
    LG.barycentres.push_back(centre);
    while (!getneighbours.empty()) {  // :119-161
        for (int gn : getneighbours) {
            for (int j : Grid.nbr[gn]) {
                Pt sample = Grid.get_coord(j);
                if ((sample - centre).norm() <= dist && found[j] == 0 && j != centroid) {
                    LG.samples.push_back(sample);
                    newneighbours.push_back(j);
                    found[j] = 1;
                }
            }
            for (int t : Grid.vtri[gn]) {
                Pt v1 = Grid.V[Grid.F[t][0]], v2 = Grid.V[Grid.F[t][1]], v3 = Grid.V[Grid.F[t][2]];
                Pt bary((v1.X + v2.X + v3.X) / 3, (v1.Y + v2.Y + v3.Y) / 3, (v1.Z + v2.Z + v3.Z) / 3);
                bary.normalize();
                bary = bary * RAD;
                if ((bary - centre).norm() <= dist && (bary - centre).norm() > 0 && found_tr[t] == 0) {
                    for (auto& m : LG.barycentres)  // :147-150 (index 0 gives 0/0 = NaN -> comparison false)
                        if (std::abs(1 - (((bary - centre) | (m - centre)) / ((bary - centre).norm() * (m - centre).norm()))) < 1e-2)
                            found_tr[t] = 1;
                    if (found_tr[t] == 0) LG.barycentres.push_back(bary);
                    found_tr[t] = 1;
                }
            }
        }
        getneighbours = newneighbours;
        newneighbours.clear();
    }
    return LG;
}

void spacings(const Mesh& CP, std::vector<double>& MAXSEP, double& MVDmax, double& MVD) {  // :48-67
    MAXSEP.assign(CP.nvertices(), 0.0);
    for (int k = 0; k < CP.nvertices(); k++) {
        Pt c = CP.get_coord(k);
        for (int j : CP.nbr[k]) {
            double dist = 2 * RAD * std::asin((c - CP.get_coord(j)).norm() / (2 * RAD));
            if (dist > MAXSEP[k]) MAXSEP[k] = dist;
        }
    }
    MVD = CP.calculate_MeanVD();
    MVDmax = CP.calculate_MaxVD();
}

std::vector<int> estimate_triplets(const Mesh& CP) {  // :267-282
    std::vector<int> t(3 * CP.ntriangles());
    for (int i = 0; i < CP.ntriangles(); i++) {
        int ids[3] = {CP.F[i][0], CP.F[i][1], CP.F[i][2]};
        std::sort(std::begin(ids), std::end(ids));
        t[3 * i] = ids[0]; t[3 * i + 1] = ids[1]; t[3 * i + 2] = ids[2];
    }
    return t;
}

std::vector<int> estimate_pairs(const Mesh& CP) {  // :245-265
    std::vector<int> p;
    for (int i = 0; i < CP.nvertices(); i++)
        for (int j : CP.nbr[i])
            if (j > i) { p.push_back(i); p.push_back(j); }
    return p;
}

std::vector<Mat3> get_rotations(const Pt& ci, const std::vector<Pt>& cps) {  // :284-294
    std::vector<Mat3> R(cps.size());
    for (size_t k = 0; k < cps.size(); k++) R[k] = estimate_rotation_matrix(ci, cps[k]);
    return R;
}

// =========================================================================================================
// mesh warp — [EXT] A8 barycentric_mesh_interpolation; unfold and helpers src/reg_tools.cpp:62-181
// =========================================================================================================
void barycentric_mesh_interpolation(std::vector<Pt>& up, const Mesh& low_init, const std::vector<Pt>& low_final) {
    SphereLocator loc(low_init);
#pragma omp parallel for schedule(dynamic, 256)
    for (int i = 0; i < (int)up.size(); i++) {
        const Pt ci = up[i];
        const int t = loc.closest_triangle(ci);
        const auto& f = low_init.F[t];
        double w[3];
        barycentric_weights(low_init.V[f[0]], low_init.V[f[1]], low_init.V[f[2]], ci, w);
        Pt q = low_final[f[0]] * w[0] + low_final[f[1]] * w[1] + low_final[f[2]] * w[2];
        q.normalize();
        up[i] = q * RAD;
    }
}

static void computeNormal2EdgeOfTriangle(const Pt& v0, const Pt& v1, const Pt& v2, Pt& norm2edge) {  // :62-82
    Pt s1 = v2 - v0, s2 = v1 - v0;
    if (s1.norm() > 1e-10) s1.normalize(); else s1 = Pt(0, 0, 0);
    if (s2.norm() > 1e-10) s2.normalize(); else s2 = Pt(0, 0, 0);
    Pt norm2triangle = s1 * s2;
    if (norm2triangle.norm() > 1e-10) norm2triangle.normalize(); else norm2triangle = Pt(0, 0, 0);
    norm2edge = s2 * norm2triangle;
    if ((s1 | norm2edge) < 0) norm2edge = norm2edge * -1;
}

static Pt computeGradientOfBarycentricTriangle(const Pt& v0, const Pt& v1, const Pt& v2) {  // :84-95
    Pt norm2edge;
    computeNormal2EdgeOfTriangle(v0, v1, v2, norm2edge);
    const double base = (v1 - v0).norm();
    return norm2edge * 0.5 * base;
}

Pt spatialgradient(int index, const Mesh& M) {  // :97-119
    Pt grad, ci = M.get_coord(index);
    for (int j : M.vtri[index]) {
        const Pt v0 = M.V[M.F[j][0]], v1 = M.V[M.F[j][1]], v2 = M.V[M.F[j][2]];
        Pt dA;
        if ((ci - v0).norm() == 0) dA = computeGradientOfBarycentricTriangle(v1, v2, v0);
        else if ((ci - v1).norm() == 0) dA = computeGradientOfBarycentricTriangle(v2, v0, v1);
        else dA = computeGradientOfBarycentricTriangle(v0, v1, v2);
        grad = grad + dA;
    }
    return grad;
}

bool check_for_intersections(int ind, const Mesh& M) {  // :121-132 (get_triangle_from_vertex(ind,0) = first incident face)
    const Pt N = M.get_triangle(M.vtri[ind][0]).normal();
    for (int j : M.vtri[ind])
        if ((N | M.get_triangle(j).normal()) <= 0.5) return true;
    return false;
}

int unfold(Mesh& SOURCE) {  // :134-181
    std::vector<int> folded;
    std::vector<Pt> grads;
    int it = 0;
    while (true) {
        folded.clear(); grads.clear();
        for (int i = 0; i < SOURCE.nvertices(); i++)
            if (check_for_intersections(i, SOURCE)) folded.push_back(i);
        if (folded.empty()) return it;
        for (int v : folded) grads.push_back(spatialgradient(v, SOURCE));
        for (size_t i = 0; i < folded.size(); i++) {
            double step = 1.0;
            const Pt ci = SOURCE.get_coord(folded[i]);
            const Pt grad = grads[i];
            Pt pp;
            do {
                pp = ci - grad * step;
                pp.normalize();
                SOURCE.set_coord(folded[i], pp * RAD);
                step *= 0.5;
            } while (check_for_intersections(folded[i], SOURCE) && step > 1e-3);
            SOURCE.set_coord(folded[i], pp * RAD);
        }
        it++;
        if (it == 1000) break;
    }
    return it;
}

// =========================================================================================================
// cost function — src/DiscreteCostFunction.cpp
// =========================================================================================================
void CostFunction::set_meshes(const Mesh& target, const Mesh& source, const Mesh& grid) {  // .h:166-168
    TARGET = target; SOURCE = source; CPgrid = grid; oCPgrid = grid;
    ORIG.assign(source.V.begin(), source.V.end());
    targettree.build(TARGET);
}

void CostFunction::set_initial_angles() {  // :95-102 with angles = CPgrid.get_face_angles() (DiscreteModel.cpp:71-72)
    auto angles = CPgrid.get_face_angles();
    double meanang = 0.0;
    for (const auto& a : angles) for (double j : a) meanang += j;
    MEANANGLE = meanang / (3.0 * angles.size());
}

bool CostFunction::within_controlpt_range(int CPindex, int sourceindex) const {  // :104-109
    return ((2 * RAD * std::asin((CPgrid.get_coord(CPindex) - SOURCE.get_coord(sourceindex)).norm() / (2 * RAD))) <
            P.range * MAXSEP[CPindex]);
}

void CostFunction::get_source_data() {  // :394-410 (uni) / :458-481 (multi): same membership, ascending source index
    sourceinrange.assign(CPgrid.nvertices(), {});
    for (int k = 0; k < CPgrid.nvertices(); k++)
        for (int i = 0; i < SOURCE.nvertices(); i++)
            if (within_controlpt_range(k, i)) sourceinrange[k].push_back(i);
    // resample_weights() (:364-383): maximum over the weight rows per source vertex, metric_resample'd ([EXT] A8, geom.h)
    // from the CURRENT source mesh onto the CURRENT control grid.  Explicitly supplied AbsoluteWeights win (tests).
    if (!absweights_given) {
        if (cfweight.empty()) AbsoluteWeights.assign(CPgrid.nvertices(), 1.0);   // the reference never runs without weights; unit weights -> 1
        else {
            std::vector<double> mx(SOURCE.nvertices());
            for (int i = 0; i < SOURCE.nvertices(); i++) {
                double m = std::numeric_limits<double>::lowest();
                for (const auto& row : cfweight) if (row[i] > m) m = row[i];
                mx[i] = m;
            }
            AbsoluteWeights = metric_resample(SOURCE, mx, 1, CPgrid);
        }
    }
}

static inline double weight_of(const std::vector<std::vector<double>>& cfw, int d, int i) {
    // :404-407 / :474-477 : row d (0-based) if present else 1.0
    if ((int)cfw.size() > d) return cfw[d][i];
    return 1.0;
}

double CostFunction::computeUnaryCost(int node, int label) const {
    // :439 / :513
    Mat3 R = estimate_rotation_matrix(CPgrid.get_coord(node), ROT[node] * labels[label]);
    const auto& patch = sourceinrange[node];
    double cost;
    if (!multivariate) {
        std::vector<double> src(patch.size()), tgt(patch.size()), w(patch.size());
        for (size_t i = 0; i < patch.size(); i++) {  // :417-434
            Pt tmp = R * SOURCE.get_coord(patch[i]);
            int t = targettree.closest_triangle(tmp);
            const auto& f = TARGET.F[t];
            tgt[i] = barycentric_weight(TARGET.V[f[0]], TARGET.V[f[1]], TARGET.V[f[2]], tmp, ref[0][f[0]], ref[0][f[1]],
                                        ref[0][f[2]], P.bary_orthogonal);
            src[i] = input[0][patch[i]];
            w[i] = weight_of(cfweight, 0, patch[i]);
        }
        cost = AbsoluteWeights[node] * get_sim_for_min(src, tgt, w);  // :441-442
    } else {
        cost = 0.0;
        std::vector<double> src(C), tgt(C), w(C);
        for (int i : patch) {  // :485-506 then :515-518
            Pt tmp = R * SOURCE.get_coord(i);
            int t = targettree.closest_triangle(tmp);
            const auto& f = TARGET.F[t];
            double bw[3];
            barycentric_weights(TARGET.V[f[0]], TARGET.V[f[1]], TARGET.V[f[2]], tmp, bw, P.bary_orthogonal);
            for (int d = 0; d < C; d++) {
                tgt[d] = bw[0] * ref[d][f[0]] + bw[1] * ref[d][f[1]] + bw[2] * ref[d][f[2]];
                src[d] = input[d][i];
                w[d] = weight_of(cfweight, d, i);
            }
            cost += get_sim_for_min(src, tgt, w);
        }
        if (!patch.empty()) cost /= patch.size();  // :520
        cost = AbsoluteWeights[node] * cost;       // :522-525
    }
    if (P.simmeasure == 1 || P.simmeasure == 2) return cost;  // :444-447
    return -cost;
}

void CostFunction::computeUnaryCosts(double* unary) const {  // :306-313: labels outer (sequential), CPs omp
    const int n = N();
    for (int j = 0; j < L(); j++) {
#pragma omp parallel for num_threads(P.threads) schedule(dynamic, 16)
        for (int k = 0; k < n; k++) unary[j * n + k] = computeUnaryCost(k, j);
    }
}

double CostFunction::computeTripletCost(int triplet, int labelA, int labelB, int labelC) const {  // :138-254
    double cost = 0.0, weight = 1.0;
    const int id[3] = {triplets[3 * triplet], triplets[3 * triplet + 1], triplets[3 * triplet + 2]};
    Pt v0 = CPgrid.get_coord(id[0]), v1 = CPgrid.get_coord(id[1]), v2 = CPgrid.get_coord(id[2]);
    Pt n0 = ROT[id[0]] * labels[labelA], n1 = ROT[id[1]] * labels[labelB], n2 = ROT[id[2]] * labels[labelC];
    Tri TRI(n0, n1, n2, 0), TRI_noDEF(v0, v1, v2, 0);
    double likelihood = 0.0;  // triplet_likelihood is 0 unless --triclique (:160)
    if ((TRI.normal() | TRI_noDEF.normal()) < 0) {  // :163-167
        cost = 1.0;
        weight += 1e6;
    } else {
        switch (P.rmode) {
            case 2: {  // :172-192
                double tweight = 1.0, diff = 0.0;
                Tri TRI_ORIG(ORIG[id[0]], ORIG[id[1]], ORIG[id[2]], 0);
                auto deformed_angles = TRI.angles();
                auto orig_angles = TRI_ORIG.angles();
                double distortion = std::log2(TRI.area() / TRI_ORIG.area());
                if (P.dweight && distortion != 1) tweight = std::exp(std::abs(distortion - 1));
                else tweight = 1.0;
                for (int i = 0; i < 3; i++) diff += std::pow(deformed_angles[i] - orig_angles[i], 2);
                cost = tweight * (std::sqrt(diff));
                if (P.anorm) weight *= (1 / MEANANGLE);
                break;
            }
            case 3: {  // :193-201
                Tri TRI_ORIG(ORIG[id[0]], ORIG[id[1]], ORIG[id[2]], 0);
                cost = calculate_triangular_strain(TRI_ORIG, TRI, P.mu, P.kappa, P.k_exp);
                break;
            }
            default:  // :241-242 (modes 4,5 need anatomical meshes: out of scope)
                throw std::runtime_error("DiscreteModel computeTripletCost regoption does not exist");
        }
    }
    if (std::abs(cost) < 1e-8) cost = 0.0;                          // :246
    return likelihood + weight * P.lambda * std::pow(cost, P.rexp);  // :253
}

double CostFunction::computePairwiseCost(int pair, int labelA, int labelB) {  // :256-296
    const int a = pairs[2 * pair], b = pairs[2 * pair + 1];
    Mat3 R1 = estimate_rotation_matrix(CPgrid.get_coord(a), ROT[a] * labels[labelA]);
    Mat3 R2 = estimate_rotation_matrix(CPgrid.get_coord(b), ROT[b] * labels[labelB]);
    Mat3 R_diff = R1.t() * R2;
    const double theta_MVD = 2 * std::asin(MVDmax / (2 * RAD));
    const double theta = std::acos((R_diff.trace() - 1) / 2);
    double weight = 1.0, totalcost = 0.0, cost = 0.0;
    Pt v0 = CPgrid.get_coord(a), v1 = CPgrid.get_coord(b);
    CPgrid.set_coord(a, ROT[a] * labels[labelA]);
    CPgrid.set_coord(b, ROT[b] * labels[labelB]);
    if (std::fabs(1 - (R_diff.trace() - 1) / 2) > EPSILON) {
        for (int j : CPgrid.vtri[a]) {  // :274-282 triangles incident to the FIRST vertex only (Q7)
            if ((oCPgrid.get_triangle(j).normal() | CPgrid.get_triangle(j).normal()) < 0) {
                cost = 1.0;  // dead store (Q7)
                weight += 1e6;
                break;
            }
        }
        cost = (std::sqrt(2) * theta) / theta_MVD;
        if (P.rexp == 1) totalcost = weight * P.lambda * cost;
        else totalcost = weight * P.lambda * std::pow(cost, P.rexp);
    }
    CPgrid.set_coord(a, v0);
    CPgrid.set_coord(b, v1);
    return totalcost;
}

void CostFunction::computePairwiseCosts(double* out) {  // :298-304
    const int np = (int)pairs.size() / 2, nl = L();
    for (int i = 0; i < np; i++)
        for (int j = 0; j < nl; j++)
            for (int k = 0; k < nl; k++) out[(size_t)i * nl * nl + k * nl + j] = computePairwiseCost(i, j, k);
}

double CostFunction::evaluateTotalCostSum(const int* labeling, bool use_unary) {  // :41-70
    double su = 0, sp = 0, st = 0;
    if (use_unary)
        for (int i = 0; i < N(); i++) su += computeUnaryCost(i, labeling[i]);
    for (int p = 0; p < (int)pairs.size() / 2; p++) sp += computePairwiseCost(p, labeling[pairs[2 * p]], labeling[pairs[2 * p + 1]]);
    for (int t = 0; t < (int)triplets.size() / 3; t++)
        st += computeTripletCost(t, labeling[triplets[3 * t]], labeling[triplets[3 * t + 1]], labeling[triplets[3 * t + 2]]);
    return su + sp + st;
}

// =========================================================================================================
// group cost function — src/DiscreteGroupCostFunction.cpp
// =========================================================================================================
double GroupCost::computeTripletCost(int triplet, int labelA, int labelB, int labelC) const {  // :9-30
    int meshID = (int)std::floor(triplet / (double)TRIPLETS_PER_SUBJ);
    int m0 = triplets[3 * triplet] - meshID * VERTICES_PER_SUBJ;
    int m1 = triplets[3 * triplet + 1] - meshID * VERTICES_PER_SUBJ;
    int m2 = triplets[3 * triplet + 2] - meshID * VERTICES_PER_SUBJ;
    Tri TRI(ROT[triplets[3 * triplet]] * labels[labelA], ROT[triplets[3 * triplet + 1]] * labels[labelB],
            ROT[triplets[3 * triplet + 2]] * labels[labelC], 0);
    const Mesh& cm = CONTROLMESHES[meshID];
    Tri TRI_noDEF(cm.get_coord(m0), cm.get_coord(m1), cm.get_coord(m2), 0);
    if ((TRI.normal() | TRI_noDEF.normal()) < 0) return 1e7;  // :23 (Q8)
    Tri TRI_ORIG(ORIG[m0], ORIG[m1], ORIG[m2], 0);
    return P.lambda * std::pow(calculate_triangular_strain(TRI_ORIG, TRI, P.mu, P.kappa), P.rexp);  // :29 k_exp default 2.0
}

double GroupCost::computePairwiseCost(int pair, int labelA, int labelB, int L) const {  // :32-54
    std::vector<double> A, B;
    int subject_A = (int)std::floor((double)pairs[2 * pair] / VERTICES_PER_SUBJ);
    int subject_B = (int)std::floor((double)pairs[2 * pair + 1] / VERTICES_PER_SUBJ);
    int node_A = pairs[2 * pair] - subject_A * VERTICES_PER_SUBJ;
    int node_B = pairs[2 * pair + 1] - subject_B * VERTICES_PER_SUBJ;
    const auto& patchA = patch_data[(size_t)subject_A * VERTICES_PER_SUBJ * L + (size_t)node_A * L + labelA];
    const auto& patchB = patch_data[(size_t)subject_B * VERTICES_PER_SUBJ * L + (size_t)node_B * L + labelB];
    for (const auto& e : patchA) {
        auto it = patchB.find(e.first);
        if (it != patchB.end()) { A.push_back(e.second); B.push_back(it->second); }
    }
    return get_sim_for_min(A, B, std::vector<double>());
}

}  // namespace orc
