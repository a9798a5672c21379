// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product: only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may build, link or call anything under oracle/.
//
// oracle/geom.h : FSL-free restatement of the external geometry semantics the reference hot path relies on
// (SURVEY.md §0.3 assumptions A1-A7).  The reference links FSL `newresampler` (src/Makefile:5, no version pin),
// which is NOT present under /root/reference, so these are restatements of its published behaviour anchored on
// the reference's own call sites — the one part of the oracle that cannot be pinned against reference code (the
// library is a dependency absent from the tree); the reference build of oracle/_ref/libref_cost.so uses the SAME
// functions through oracle/ext/fsl_ext.h, so these semantics exist exactly once.
//
// Everything is double precision, single-thread semantics, no global state.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <unordered_map>
#include <utility>
#include <vector>

namespace orc {

constexpr double RAD = 100.0;       // src/DiscreteCostFunction.h:9
constexpr double EPSILON = 1.0E-8;  // src/DiscreteCostFunction.h:10

// ---------------------------------------------------------------------------------------------------------
// A1: Point algebra.  Call sites: src/reg_tools.cpp:64-80 (a*b cross, a|b dot, p*s scale, norm, normalize),
//     src/DiscreteCostFunction.cpp:151,163.
// ---------------------------------------------------------------------------------------------------------
struct Pt {
    double X = 0, Y = 0, Z = 0;
    Pt() = default;
    Pt(double x, double y, double z) : X(x), Y(y), Z(z) {}
    double norm() const { return std::sqrt(X * X + Y * Y + Z * Z); }
    void normalize() {
        double n = norm();
        if (n != 0) { X /= n; Y /= n; Z /= n; }
    }
};
inline Pt operator+(const Pt& a, const Pt& b) { return {a.X + b.X, a.Y + b.Y, a.Z + b.Z}; }
inline Pt operator-(const Pt& a, const Pt& b) { return {a.X - b.X, a.Y - b.Y, a.Z - b.Z}; }
inline Pt operator*(const Pt& a, double s) { return {a.X * s, a.Y * s, a.Z * s}; }
inline Pt operator/(const Pt& a, double s) { return {a.X / s, a.Y / s, a.Z / s}; }
inline double operator|(const Pt& a, const Pt& b) { return a.X * b.X + a.Y * b.Y + a.Z * b.Z; }  // dot
inline Pt operator*(const Pt& a, const Pt& b) {                                                      // cross
    return {a.Y * b.Z - a.Z * b.Y, a.Z * b.X - a.X * b.Z, a.X * b.Y - a.Y * b.X};
}

// A9 stand-in for a 3x3 NEWMAT::Matrix (row-major, 0-based here).
struct Mat3 {
    double m[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    static Mat3 zero() { Mat3 r; for (auto& row : r.m) for (double& v : row) v = 0; return r; }
    Mat3 t() const { Mat3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = m[j][i]; return r; }
    double trace() const { return m[0][0] + m[1][1] + m[2][2]; }
    double det() const {
        return m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
               m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
    }
};
inline Mat3 operator*(const Mat3& a, const Mat3& b) {
    Mat3 r = Mat3::zero();
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) r.m[i][j] += a.m[i][k] * b.m[k][j];
    return r;
}
inline Mat3 operator+(const Mat3& a, const Mat3& b) {
    Mat3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] + b.m[i][j]; return r;
}
inline Mat3 operator*(const Mat3& a, double s) {
    Mat3 r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] * s; return r;
}
inline Pt operator*(const Mat3& M, const Pt& p) {  // `R * Point` (src/DiscreteCostFunction.cpp:151,419)
    return {M.m[0][0] * p.X + M.m[0][1] * p.Y + M.m[0][2] * p.Z, M.m[1][0] * p.X + M.m[1][1] * p.Y + M.m[1][2] * p.Z,
            M.m[2][0] * p.X + M.m[2][1] * p.Y + M.m[2][2] * p.Z};
}
// A6: form_matrix_from_points(p1,p2,p3): points as COLUMNS (src/reg_tools.cpp:717,725).
inline Mat3 form_matrix_from_points(const Pt& p1, const Pt& p2, const Pt& p3) {
    Mat3 r;
    r.m[0][0] = p1.X; r.m[0][1] = p2.X; r.m[0][2] = p3.X;
    r.m[1][0] = p1.Y; r.m[1][1] = p2.Y; r.m[1][2] = p3.Y;
    r.m[2][0] = p1.Z; r.m[2][1] = p2.Z; r.m[2][2] = p3.Z;
    return r;
}

// ---------------------------------------------------------------------------------------------------------
// A2: estimate_rotation_matrix(a,b): minimal (Rodrigues) rotation taking direction a onto direction b;
//     identity when the angle is < 1e-8.  Call sites: src/DiscreteModel.cpp:293,
//     src/DiscreteCostFunction.cpp:258-259,439,513, src/DiscreteGroupModel.cpp:69,85.
// ---------------------------------------------------------------------------------------------------------
inline Mat3 estimate_rotation_matrix(const Pt& p1, const Pt& p2) {
    Pt ci = p1, index = p2;
    ci.normalize();
    index.normalize();
    double c = ci | index;
    if (c > 1.0) c = 1.0;
    if (c < -1.0) c = -1.0;
    double theta = std::acos(c);
    Mat3 I;
    if (std::fabs(theta) < 1e-8) return I;
    Pt k = ci * index;
    if (k.norm() < 1e-12 && c < 0) {
        // Antiparallel directions: the minimal rotation is by pi about ANY axis perpendicular to a; the published formula
        // degenerates (axis = normalised zero vector).  Every regular icosphere contains the antipode of the label-grid
        // centre, so this input does occur (src/DiscreteModel.cpp:284-294).  Convention fixed here (and in the product):
        // axis = a x e, e = the coordinate axis along which a has its smallest component.
        const double ax = std::fabs(ci.X), ay = std::fabs(ci.Y), az = std::fabs(ci.Z);
        Pt e = (ax <= ay && ax <= az) ? Pt(1, 0, 0) : ((ay <= az) ? Pt(0, 1, 0) : Pt(0, 0, 1));
        Pt u = ci * e;
        u.normalize();
        Mat3 R;  // I + 2 (u u^T - I)
        const double uu[3] = {u.X, u.Y, u.Z};
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) R.m[i][j] = 2.0 * uu[i] * uu[j] - (i == j ? 1.0 : 0.0);
        return R;
    }
    k.normalize();
    Mat3 u = Mat3::zero();
    u.m[0][1] = -k.Z; u.m[0][2] = k.Y;
    u.m[1][0] = k.Z;  u.m[1][2] = -k.X;
    u.m[2][0] = -k.Y; u.m[2][1] = k.X;
    return I + u * std::sin(theta) + (u * u) * (1.0 - std::cos(theta));
}

// ---------------------------------------------------------------------------------------------------------
// A5: Triangle(v0,v1,v2): unit normal (v1-v0)x(v2-v0), area, interior angles.
//     Call sites: src/DiscreteCostFunction.cpp:155-181.
// ---------------------------------------------------------------------------------------------------------
struct Tri {
    Pt v[3];
    int n[3] = {-1, -1, -1};
    int id = -1;
    Tri() = default;
    Tri(const Pt& a, const Pt& b, const Pt& c, int id_ = 0) : id(id_) { v[0] = a; v[1] = b; v[2] = c; }
    Pt normal() const {
        Pt nn = (v[1] - v[0]) * (v[2] - v[0]);
        nn.normalize();
        return nn;
    }
    double area() const { return 0.5 * ((v[1] - v[0]) * (v[2] - v[0])).norm(); }
    std::array<double, 3> angles() const {
        std::array<double, 3> a{};
        for (int i = 0; i < 3; i++) {
            Pt e1 = v[(i + 1) % 3] - v[i], e2 = v[(i + 2) % 3] - v[i];
            double c = (e1 | e2) / (e1.norm() * e2.norm());
            if (c > 1) c = 1;
            if (c < -1) c = -1;
            a[i] = std::acos(c);
        }
        return a;
    }
};

// ---------------------------------------------------------------------------------------------------------
// A4: barycentric_weight(v0,v1,v2,p,f0,f1,f2): project p into the triangle's plane along the ray from the
//     origin (MSM lineage `projectPoint`: PP = p * (n.v0)/(n.p)), weights = sub-triangle areas / their sum.
//     Call sites: src/DiscreteCostFunction.cpp:430,502.   `orthogonal=true` selects the other plausible
//     convention (orthogonal projection); it is kept switchable because the FSL source is not available.
// ---------------------------------------------------------------------------------------------------------
inline Pt project_point(const Pt& p, const Pt& v0, const Pt& v1, const Pt& v2, bool orthogonal = false) {
    Pt s1 = v2 - v0, s2 = v1 - v0;
    s1.normalize();
    s2.normalize();
    Pt s3 = s1 * s2;
    s3.normalize();
    if (orthogonal) return p - s3 * ((p - v0) | s3);
    double si = (s3 | v0) / (s3 | p);
    return p * si;
}
inline double tri_area(const Pt& a, const Pt& b, const Pt& c) { return 0.5 * ((b - a) * (c - a)).norm(); }
inline void barycentric_weights(const Pt& v0, const Pt& v1, const Pt& v2, const Pt& p, double w[3], bool orthogonal = false) {
    Pt PP = project_point(p, v0, v1, v2, orthogonal);
    double Aa = tri_area(PP, v1, v2), Ab = tri_area(PP, v0, v2), Ac = tri_area(PP, v0, v1);
    double A = Aa + Ab + Ac;
    w[0] = Aa / A; w[1] = Ab / A; w[2] = Ac / A;
}
inline double barycentric_weight(const Pt& v0, const Pt& v1, const Pt& v2, const Pt& p, double f0, double f1, double f2,
                                 bool orthogonal = false) {
    double w[3];
    barycentric_weights(v0, v1, v2, p, w, orthogonal);
    return w[0] * f0 + w[1] * f1 + w[2] * f2;
}

// ---------------------------------------------------------------------------------------------------------
// A7: Mesh container + icosphere generator (make_mesh_from_icosa / true_rescale / recentre stand-in).
//     k-times 4-split icosahedron, midpoints pushed to radius R, NESTED vertex numbering (coarse ids kept),
//     child faces of face f are 4f..4f+3.  Call sites: src/featurespace.cc:29-31,
//     src/mesh_registration.cpp:67-69, src/DiscreteModel.cpp:44-71,93-103.
//     Adjacency iteration order (nbegin/nend, tIDbegin/tIDend) is [EXT]-unknown; convention here:
//     triangles ascending; neighbours in order of first appearance over those triangles (v0,v1,v2 order).
// ---------------------------------------------------------------------------------------------------------
struct Mesh {
    std::vector<Pt> V;
    std::vector<std::array<int, 3>> F;
    std::vector<std::vector<int>> nbr;   // vertex -> neighbour vertices
    std::vector<std::vector<int>> vtri;  // vertex -> incident triangle ids

    int nvertices() const { return (int)V.size(); }
    int ntriangles() const { return (int)F.size(); }
    const Pt& get_coord(int i) const { return V[i]; }
    void set_coord(int i, const Pt& p) { V[i] = p; }
    Tri get_triangle(int t) const {
        Tri T(V[F[t][0]], V[F[t][1]], V[F[t][2]], t);
        T.n[0] = F[t][0]; T.n[1] = F[t][1]; T.n[2] = F[t][2];
        return T;
    }
    int get_total_neighbours(int i) const { return (int)nbr[i].size(); }

    void build_adjacency() {
        nbr.assign(V.size(), {});
        vtri.assign(V.size(), {});
        for (int t = 0; t < (int)F.size(); t++)
            for (int j = 0; j < 3; j++) vtri[F[t][j]].push_back(t);
        for (int v = 0; v < (int)V.size(); v++)
            for (int t : vtri[v])
                for (int j = 0; j < 3; j++) {
                    int u = F[t][j];
                    if (u != v && std::find(nbr[v].begin(), nbr[v].end(), u) == nbr[v].end()) nbr[v].push_back(u);
                }
    }
    double calculate_MaxVD() const {  // max Euclidean edge length (src/DiscreteModel.cpp:65)
        double mx = 0;
        for (int i = 0; i < nvertices(); i++)
            for (int j : nbr[i]) mx = std::max(mx, (V[i] - V[j]).norm());
        return mx;
    }
    double calculate_MeanVD() const {
        double s = 0; long n = 0;
        for (int i = 0; i < nvertices(); i++)
            for (int j : nbr[i]) { s += (V[i] - V[j]).norm(); n++; }
        return n ? s / n : 0;
    }
    std::vector<std::array<double, 3>> get_face_angles() const {
        std::vector<std::array<double, 3>> a(F.size());
        for (int t = 0; t < (int)F.size(); t++) a[t] = get_triangle(t).angles();
        return a;
    }
};

inline Mesh make_mesh_from_icosa(int level, double radius = RAD) {
    Mesh M;
    const double t = (1.0 + std::sqrt(5.0)) / 2.0;
    const double base[12][3] = {{-1, t, 0}, {1, t, 0}, {-1, -t, 0}, {1, -t, 0}, {0, -1, t}, {0, 1, t},
                                {0, -1, -t}, {0, 1, -t}, {t, 0, -1}, {t, 0, 1}, {-t, 0, -1}, {-t, 0, 1}};
    const int faces[20][3] = {{0, 11, 5}, {0, 5, 1},  {0, 1, 7},   {0, 7, 10}, {0, 10, 11}, {1, 5, 9}, {5, 11, 4},
                              {11, 10, 2}, {10, 7, 6}, {7, 1, 8},  {3, 9, 4},  {3, 4, 2},   {3, 2, 6}, {3, 6, 8},
                              {3, 8, 9},  {4, 9, 5},  {2, 4, 11},  {6, 2, 10}, {8, 6, 7},   {9, 8, 1}};
    // the reference builds the unit icosphere and then calls true_rescale(m, RAD) (src/mesh_registration.cpp:67-69,
    // src/DiscreteModel.cpp:93-94): subdivide at unit radius, rescale at the end
    for (auto& b : base) {
        Pt p(b[0], b[1], b[2]);
        p.normalize();
        M.V.push_back(p);
    }
    for (auto& f : faces) M.F.push_back({f[0], f[1], f[2]});
    for (int l = 0; l < level; l++) {
        std::map<std::pair<int, int>, int> mid;
        auto midpoint = [&](int a, int b) {
            std::pair<int, int> key(std::min(a, b), std::max(a, b));
            auto it = mid.find(key);
            if (it != mid.end()) return it->second;
            Pt p = M.V[a] + M.V[b];
            p.normalize();
            M.V.push_back(p);
            int id = (int)M.V.size() - 1;
            mid[key] = id;
            return id;
        };
        std::vector<std::array<int, 3>> NF;
        NF.reserve(M.F.size() * 4);
        for (auto& f : M.F) {
            int a = f[0], b = f[1], c = f[2];
            int ab = midpoint(a, b), bc = midpoint(b, c), ca = midpoint(c, a);
            NF.push_back({a, ab, ca});
            NF.push_back({ab, b, bc});
            NF.push_back({ca, bc, c});
            NF.push_back({ab, bc, ca});
        }
        M.F.swap(NF);
    }
    for (auto& p : M.V) {  // true_rescale
        p.normalize();
        p = p * radius;
    }
    M.build_adjacency();
    return M;
}

// ---------------------------------------------------------------------------------------------------------
// A3: Octree(mesh)::get_closest_triangle(p): the triangle of a closed spherical mesh whose radial cone
//     contains p.  Restated as a uniform voxel grid over unit directions (any spherical mesh, deformed or
//     regular); `brute_closest_triangle` is the definition it is validated against in tests.
//     Call sites: src/DiscreteModel.cpp:88, src/DiscreteCostFunction.cpp:421,492.
// ---------------------------------------------------------------------------------------------------------
// >0 inside test for the cone of (a,b,c): all three triple products non-negative (with tolerance tol).
inline bool in_cone(const Pt& p, const Pt& a, const Pt& b, const Pt& c, double tol, double* minval = nullptr) {
    double s0 = p | (a * b), s1 = p | (b * c), s2 = p | (c * a);
    double mn = std::min(s0, std::min(s1, s2));
    if (minval) *minval = mn;
    return mn >= -tol;
}

inline int brute_closest_triangle(const Mesh& M, const Pt& p) {
    int best = -1;
    double bestv = -1e300;
    Pt u = p; u.normalize();
    for (int t = 0; t < M.ntriangles(); t++) {
        Pt a = M.V[M.F[t][0]], b = M.V[M.F[t][1]], c = M.V[M.F[t][2]];
        a.normalize(); b.normalize(); c.normalize();
        if ((u | (a + b + c)) <= 0) continue;  // wrong hemisphere
        double mn;
        in_cone(u, a, b, c, 0.0, &mn);
        if (mn > bestv) { bestv = mn; best = t; }
    }
    return best;
}

class SphereLocator {
public:
    SphereLocator() = default;
    explicit SphereLocator(const Mesh& M) { build(M); }
    void build(const Mesh& M) {
        mesh = &M;
        const int nt = M.ntriangles();
        // cell size ~ mean angular edge length
        double mean_edge = 0;
        for (int t = 0; t < nt; t++) mean_edge += unit(M.V[M.F[t][0]], M.V[M.F[t][1]]);
        mean_edge /= std::max(nt, 1);
        G = std::max(4, std::min(256, (int)std::floor(2.0 / std::max(mean_edge, 1e-6))));
        h = 2.0 / G;
        std::vector<int> count((size_t)G * G * G + 1, 0);
        std::vector<std::array<int, 6>> box(nt);
        for (int t = 0; t < nt; t++) {
            Pt a = M.V[M.F[t][0]], b = M.V[M.F[t][1]], c = M.V[M.F[t][2]];
            a.normalize(); b.normalize(); c.normalize();
            double emax = std::max((a - b).norm(), std::max((b - c).norm(), (c - a).norm()));
            // spherical triangle bulges beyond the chord triangle by at most the sagitta of its longest edge-circle
            double sag = 1.0 - std::sqrt(std::max(0.0, 1.0 - emax * emax / 3.0)) + 1e-9;
            double lo[3] = {std::min(a.X, std::min(b.X, c.X)) - sag, std::min(a.Y, std::min(b.Y, c.Y)) - sag,
                            std::min(a.Z, std::min(b.Z, c.Z)) - sag};
            double hi[3] = {std::max(a.X, std::max(b.X, c.X)) + sag, std::max(a.Y, std::max(b.Y, c.Y)) + sag,
                            std::max(a.Z, std::max(b.Z, c.Z)) + sag};
            for (int d = 0; d < 3; d++) { box[t][d] = cell(lo[d]); box[t][3 + d] = cell(hi[d]); }
            for (int x = box[t][0]; x <= box[t][3]; x++)
                for (int y = box[t][1]; y <= box[t][4]; y++)
                    for (int z = box[t][2]; z <= box[t][5]; z++) count[idx(x, y, z) + 1]++;
        }
        start.assign(count.size(), 0);
        for (size_t i = 1; i < count.size(); i++) start[i] = start[i - 1] + count[i];
        items.assign(start.back(), 0);
        std::vector<int> fill(start.begin(), start.end() - 1);
        for (int t = 0; t < nt; t++)
            for (int x = box[t][0]; x <= box[t][3]; x++)
                for (int y = box[t][1]; y <= box[t][4]; y++)
                    for (int z = box[t][2]; z <= box[t][5]; z++) items[fill[idx(x, y, z)]++] = t;
    }
    // Returns triangle id whose radial cone contains p (best-margin candidate on ties); -1 on miss.
    int closest_triangle(const Pt& p) const {
        Pt u = p; u.normalize();
        size_t c = idx(cell(u.X), cell(u.Y), cell(u.Z));
        int best = -1; double bestv = -1e300;
        for (int k = start[c]; k < start[c + 1]; k++) {
            int t = items[k];
            Pt a = mesh->V[mesh->F[t][0]], b = mesh->V[mesh->F[t][1]], cc = mesh->V[mesh->F[t][2]];
            a.normalize(); b.normalize(); cc.normalize();
            if ((u | (a + b + cc)) <= 0) continue;
            double mn; in_cone(u, a, b, cc, 0.0, &mn);
            if (mn > bestv) { bestv = mn; best = t; }
        }
        if (best < 0 || bestv < -1e-9) {  // folded / degenerate neighbourhood: fall back to the definition
            int b2 = brute_closest_triangle(*mesh, p);
            if (b2 >= 0) best = b2;
        }
        return best;
    }
private:
    static double unit(const Pt& a, const Pt& b) { Pt x = a, y = b; x.normalize(); y.normalize(); return (x - y).norm(); }
    int cell(double x) const { int c = (int)std::floor((x + 1.0) / h); return std::max(0, std::min(G - 1, c)); }
    size_t idx(int x, int y, int z) const { return ((size_t)x * G + y) * G + z; }
    const Mesh* mesh = nullptr;
    int G = 0; double h = 0;
    std::vector<int> start, items;
};

// ---------------------------------------------------------------------------------------------------------
// A8: metric_resample(in, ref) — FSL newresampler's ADAPTIVE BARYCENTRIC resampling ("ADAP_BARY") of per-vertex data
//     from mesh `in` onto the vertices of mesh `ref`.  Call sites: src/DiscreteGroupModel.cpp:92 (every subject's
//     patch-wise rotated data mesh -> the template, once per subject and label), src/DiscreteCostFunction.cpp:364-383
//     (resample_weights: cost-function weights -> control grid), src/featurespace.cc:46, src/group_mesh_registration.cpp:67.
//     The FSL source is not in the tree; restated from the published algorithm it ports (Connectome Workbench
//     SurfaceResamplingHelper::computeWeightsAdapBaryArea, "ADAP_BARY_AREA"), with per-vertex areas taken from the two
//     spheres themselves (one third of the incident triangle areas):
//       forward[k]  = barycentric weights of ref vertex k in the `in` triangle under it          (3 sources)
//       reverse[j]  = barycentric weights of in vertex j in the `ref` triangle under it, re-ordered per ref vertex
//       adapt[k]    = reverse_reorder[k] if it has MORE sources than forward[k], else forward[k]   ("whichever mesh is finer")
//       w          *= area_ref[k];   corr[j] = sum over k of w(k,j);   w *= area_in[j] / corr[j];   normalise over j.
//     Sources are visited in ascending vertex order (std::map order), sums accumulate in that order.
//     A constant field is reproduced exactly (weights sum to 1), which is why unit cost-function weights give
//     AbsoluteWeights == 1 whatever the details.  No exclusion mask on this path.
// ---------------------------------------------------------------------------------------------------------
inline std::vector<double> vertex_areas(const Mesh& M) {
    std::vector<double> a(M.nvertices(), 0.0);
    for (int t = 0; t < M.ntriangles(); t++) {
        const double A = tri_area(M.V[M.F[t][0]], M.V[M.F[t][1]], M.V[M.F[t][2]]) / 3.0;
        for (int j = 0; j < 3; j++) a[M.F[t][j]] += A;
    }
    return a;
}
struct ResampleWeights {      // CSR over the vertices of `ref`; sources ascending within a row
    std::vector<int> off, src;
    std::vector<double> w;
};
// tri/w of every query (closest triangle of M under p, barycentric weights in that triangle's vertex order)
inline void locate_all(const Mesh& M, const std::vector<Pt>& pts, std::vector<int>& tri, std::vector<double>& w) {
    SphereLocator loc(M);
    tri.assign(pts.size(), -1); w.assign(pts.size() * 3, 0.0);
    for (size_t i = 0; i < pts.size(); i++) {
        const int t = loc.closest_triangle(pts[i]);
        tri[i] = t;
        if (t >= 0) barycentric_weights(M.V[M.F[t][0]], M.V[M.F[t][1]], M.V[M.F[t][2]], pts[i], &w[3 * i]);
    }
}
// The weight assembly, shared by every implementation of A8: inputs are the two sets of point locations.
inline ResampleWeights adaptive_barycentric_assemble(int Vin, int Vref, const int* in_faces /*[Fin][3]*/, const int* ref_faces /*[Fref][3]*/,
                                                     const int* fwd_tri, const double* fwd_w /*per ref vertex, in `in`*/, const int* rev_tri,
                                                     const double* rev_w /*per in vertex, in `ref`*/, const double* area_in, const double* area_ref) {
    std::vector<std::map<int, double>> adapt(Vref), rr(Vref);
    for (int j = 0; j < Vin; j++) {
        if (rev_tri[j] < 0) continue;
        for (int c = 0; c < 3; c++) rr[ref_faces[3 * rev_tri[j] + c]][j] = rev_w[3 * j + c];
    }
    std::vector<double> corr(Vin, 0.0);
    for (int k = 0; k < Vref; k++) {
        std::map<int, double> fw;
        if (fwd_tri[k] >= 0)
            for (int c = 0; c < 3; c++) fw[in_faces[3 * fwd_tri[k] + c]] = fwd_w[3 * k + c];
        adapt[k] = rr[k].size() > fw.size() ? rr[k] : fw;
        for (auto& e : adapt[k]) { e.second *= area_ref[k]; corr[e.first] += e.second; }
    }
    ResampleWeights R;
    R.off.assign(Vref + 1, 0);
    for (int k = 0; k < Vref; k++) {
        double sum = 0.0;
        for (auto& e : adapt[k]) { e.second *= area_in[e.first] / corr[e.first]; sum += e.second; }
        for (auto& e : adapt[k]) { R.src.push_back(e.first); R.w.push_back(sum != 0.0 ? e.second / sum : e.second); }
        R.off[k + 1] = (int)R.src.size();
    }
    return R;
}
inline ResampleWeights adaptive_barycentric_weights(const Mesh& in, const Mesh& ref) {
    std::vector<int> ft, rt; std::vector<double> fw, rw;
    locate_all(in, ref.V, ft, fw);
    if (ref.ntriangles() > 0) locate_all(ref, in.V, rt, rw);
    else { rt.assign(in.nvertices(), -1); rw.assign((size_t)in.nvertices() * 3, 0.0); }   // `ref` is a bare point set: forward weights only
    std::vector<int> fi((size_t)in.ntriangles() * 3), fr((size_t)ref.ntriangles() * 3);
    for (int t = 0; t < in.ntriangles(); t++) for (int j = 0; j < 3; j++) fi[3 * t + j] = in.F[t][j];
    for (int t = 0; t < ref.ntriangles(); t++) for (int j = 0; j < 3; j++) fr[3 * t + j] = ref.F[t][j];
    const std::vector<double> ai = vertex_areas(in);
    const std::vector<double> ar = ref.ntriangles() > 0 ? vertex_areas(ref) : std::vector<double>(ref.nvertices(), 1.0);
    return adaptive_barycentric_assemble(in.nvertices(), ref.nvertices(), fi.data(), fr.data(), ft.data(), fw.data(), rt.data(), rw.data(), ai.data(),
                                         ar.data());
}
// values[d*Vin + j] -> out[d*Vref + k]
inline std::vector<double> metric_resample(const Mesh& in, const std::vector<double>& values, int dims, const Mesh& ref) {
    const ResampleWeights R = adaptive_barycentric_weights(in, ref);
    const int Vin = in.nvertices(), Vref = ref.nvertices();
    std::vector<double> out((size_t)dims * Vref, 0.0);
    for (int d = 0; d < dims; d++)
        for (int k = 0; k < Vref; k++) {
            double v = 0.0;
            for (int e = R.off[k]; e < R.off[k + 1]; e++) v += R.w[e] * values[(size_t)d * Vin + R.src[e]];
            out[(size_t)d * Vref + k] = v;
        }
    return out;
}

}  // namespace orc
