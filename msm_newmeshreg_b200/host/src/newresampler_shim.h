// newresampler_shim.h — FSL-free stand-ins for the external types that appear in the signatures of the reference's
// DiscreteCostFunction / DiscreteModel interfaces (FSL newresampler + NEWMAT).  The reference
// links them from FSL (src/Makefile:5); here they are the minimum the boundary needs, so that host code written against
// the reference's class surface compiles unchanged and the unmodified vendored optimisers (include/FastPD, include/ELC,
// include/Fusion) can be driven by the GPU-backed cost functions.
//
// This is product host code (the model side of the boundary stays on the host, north_star); it does NOT evaluate costs.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace NEWMAT {
// 1-based dense matrix, the subset used at the boundary (NEWMAT::Matrix / ColumnVector; A9).
class Matrix {
public:
    Matrix() = default;
    Matrix(int r, int c) : r_(r), c_(c), d_((size_t)r * c, 0.0) {}
    void ReSize(int r, int c) { r_ = r; c_ = c; d_.assign((size_t)r * c, 0.0); }
    void resize(int r, int c) { ReSize(r, c); }
    int Nrows() const { return r_; }
    int Ncols() const { return c_; }
    double& operator()(int i, int j) { return d_[(size_t)(i - 1) * c_ + (j - 1)]; }
    double operator()(int i, int j) const { return d_[(size_t)(i - 1) * c_ + (j - 1)]; }
    Matrix& operator=(double v) { std::fill(d_.begin(), d_.end(), v); return *this; }
    const double* data() const { return d_.data(); }
    double* data() { return d_.data(); }
protected:
    int r_ = 0, c_ = 0;
    std::vector<double> d_;
};
class ColumnVector : public Matrix {
public:
    ColumnVector() = default;
    explicit ColumnVector(int n) : Matrix(n, 1) {}
    void ReSize(int n) { Matrix::ReSize(n, 1); }
    double& operator()(int i) { return d_[i - 1]; }
    double operator()(int i) const { return d_[i - 1]; }
    ColumnVector& operator=(double v) { std::fill(d_.begin(), d_.end(), v); return *this; }
};
}  // namespace NEWMAT

namespace newresampler {

struct Point {  // A1
    double X = 0, Y = 0, Z = 0;
    Point() = default;
    Point(double x, double y, double z) : X(x), Y(y), Z(z) {}
    double norm() const { return std::sqrt(X * X + Y * Y + Z * Z); }
    void normalize() { double n = norm(); if (n != 0) { X /= n; Y /= n; Z /= n; } }
};
inline Point operator+(const Point& a, const Point& b) { return {a.X + b.X, a.Y + b.Y, a.Z + b.Z}; }
inline Point operator-(const Point& a, const Point& b) { return {a.X - b.X, a.Y - b.Y, a.Z - b.Z}; }
inline Point operator*(const Point& a, double s) { return {a.X * s, a.Y * s, a.Z * s}; }
inline double operator|(const Point& a, const Point& b) { return a.X * b.X + a.Y * b.Y + a.Z * b.Z; }
inline Point operator*(const Point& a, const Point& b) { return {a.Y * b.Z - a.Z * b.Y, a.Z * b.X - a.X * b.Z, a.X * b.Y - a.Y * b.X}; }
inline Point operator*(const NEWMAT::Matrix& M, const Point& p) {  // `R * Point` (src/DiscreteModel.cpp:242)
    return {M(1, 1) * p.X + M(1, 2) * p.Y + M(1, 3) * p.Z, M(2, 1) * p.X + M(2, 2) * p.Y + M(2, 3) * p.Z,
            M(3, 1) * p.X + M(3, 2) * p.Y + M(3, 3) * p.Z};
}

// A2: minimal rotation taking direction a onto b (identity below 1e-8 rad)
inline NEWMAT::Matrix estimate_rotation_matrix(const Point& p1, const Point& p2) {
    Point a = p1, b = p2;
    a.normalize(); b.normalize();
    double c = std::max(-1.0, std::min(1.0, a | b));
    double theta = std::acos(c);
    NEWMAT::Matrix R(3, 3);
    R(1, 1) = R(2, 2) = R(3, 3) = 1.0;
    if (std::fabs(theta) < 1e-8) return R;
    Point k = a * b;
    if (k.norm() < 1e-12 && c < 0) {
        // antiparallel (the antipode of the label-grid centre is a control point of every regular icosphere): rotation by
        // pi about a x e, e = coordinate axis of a's smallest component — same convention as the device (locate.cuh)
        const double ax = std::fabs(a.X), ay = std::fabs(a.Y), az = std::fabs(a.Z);
        Point u = a * ((ax <= ay && ax <= az) ? Point(1, 0, 0) : ((ay <= az) ? Point(0, 1, 0) : Point(0, 0, 1)));
        u.normalize();
        const double uu[3] = {u.X, u.Y, u.Z};
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) R(i + 1, j + 1) = 2.0 * uu[i] * uu[j] - (i == j ? 1.0 : 0.0);
        return R;
    }
    k.normalize();
    const double s = std::sin(theta), oc = 1.0 - std::cos(theta);
    const double u[3][3] = {{0, -k.Z, k.Y}, {k.Z, 0, -k.X}, {-k.Y, k.X, 0}};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double uu = 0;
            for (int m = 0; m < 3; m++) uu += u[i][m] * u[m][j];
            R(i + 1, j + 1) = (R(i + 1, j + 1) + u[i][j] * s) + uu * oc;   // (I + u sin) + u^2 (1 - cos): this association, so that
                                                                            // displaced control points are bit-identical to the
                                                                            // oracle's and exact range ties resolve the same way
        }
    return R;
}

class Triangle {  // A5 (the subset the model side uses)
public:
    Triangle() = default;
    Triangle(const Point& a, const Point& b, const Point& c, int id) : id_(id) { v_[0] = a; v_[1] = b; v_[2] = c; }
    const Point& get_vertex_coord(int i) const { return v_[i]; }
    int get_vertex_no(int i) const { return n_[i]; }
    void set_vertex_no(int i, int n) { n_[i] = n; }
    int get_no() const { return id_; }
    Point normal() const { Point n = (v_[1] - v_[0]) * (v_[2] - v_[0]); n.normalize(); return n; }
    double get_area() const { return 0.5 * ((v_[1] - v_[0]) * (v_[2] - v_[0])).norm(); }
    std::vector<double> get_angles() const {
        std::vector<double> a(3);
        for (int i = 0; i < 3; i++) {
            Point e1 = v_[(i + 1) % 3] - v_[i], e2 = v_[(i + 2) % 3] - v_[i];
            double c = (e1 | e2) / (e1.norm() * e2.norm());
            a[i] = std::acos(std::max(-1.0, std::min(1.0, c)));
        }
        return a;
    }
private:
    Point v_[3];
    int n_[3] = {-1, -1, -1};
    int id_ = -1;
};

class Mesh {  // A7 (container + adjacency; no I/O: on-disk formats stay outside the hot path)
public:
    int nvertices() const { return (int)V().size(); }
    int ntriangles() const { return (int)topo_->F.size(); }
    const Point& get_coord(int i) const { return V()[i]; }
    void set_coord(int i, const Point& p) { Vw()[i] = p; }
    // all vertices at once from a flat [nv][3] array: one bulk copy into storage of its own (other copies keep theirs)
    void set_coords(const double* xyz) {
        static_assert(sizeof(Point) == 3 * sizeof(double), "Point must be three packed doubles");
        const Point* first = reinterpret_cast<const Point*>(xyz);
        V_ = std::make_shared<std::vector<Point>>(first, first + V().size());   // one pass, no zero fill
    }
    int get_triangle_vertexID(int t, int j) const { return topo_->F[t][j]; }
    Point get_triangle_vertex(int t, int j) const { return V()[topo_->F[t][j]]; }
    Triangle get_triangle(int t) const {
        Triangle T(V()[topo_->F[t][0]], V()[topo_->F[t][1]], V()[topo_->F[t][2]], t);
        for (int j = 0; j < 3; j++) T.set_vertex_no(j, topo_->F[t][j]);
        return T;
    }
    std::vector<int>::const_iterator nbegin(int i) const { return adj().nbr[i].begin(); }
    std::vector<int>::const_iterator nend(int i) const { return adj().nbr[i].end(); }
    std::vector<int>::const_iterator tIDbegin(int i) const { return adj().vtri[i].begin(); }
    std::vector<int>::const_iterator tIDend(int i) const { return adj().vtri[i].end(); }
    int get_total_neighbours(int i) const { return (int)adj().nbr[i].size(); }
    int get_resolution() const { return resolution_; }
    double calculate_MaxVD() const {
        double mx = 0;
        for (int i = 0; i < nvertices(); i++) for (int j : adj().nbr[i]) mx = std::max(mx, (V()[i] - V()[j]).norm());
        return mx;
    }
    double calculate_MeanVD() const {
        double s = 0; long n = 0;
        for (int i = 0; i < nvertices(); i++) for (int j : adj().nbr[i]) { s += (V()[i] - V()[j]).norm(); n++; }
        return n ? s / n : 0;
    }
    std::vector<std::vector<double>> get_face_angles() const {
        std::vector<std::vector<double>> a(topo_->F.size());
        for (int t = 0; t < ntriangles(); t++) a[t] = get_triangle(t).get_angles();
        return a;
    }
    // per-vertex values ("pvalues", one row per dimension): carried for the driver's cost-function weighting
    // (src/mesh_registration.cpp:323-327); shared between copies until one side sets new ones
    void set_pvalues(const NEWMAT::Matrix& M) { pv_ = std::make_shared<NEWMAT::Matrix>(M); }
    NEWMAT::Matrix get_pvalues() const { return pv_ ? *pv_ : NEWMAT::Matrix(); }
    int npvalues() const { return pv_ ? pv_->Ncols() : 0; }
    int get_dimension() const { return pv_ ? pv_->Nrows() : 0; }
    double get_pvalue(int i, int dim = 0) const { return (*pv_)(dim + 1, i + 1); }
    // construction
    void set(const double* xyz, int nv, const int* tris, int nf, int resolution = -1) {
        topo_ = std::make_shared<Topo>();  // never mutate a topology another copy may share
        V_ = std::make_shared<std::vector<Point>>(nv);
        for (int i = 0; i < nv; i++) (*V_)[i] = Point(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
        topo_->F.resize(nf);
        for (int t = 0; t < nf; t++) topo_->F[t] = {tris[3 * t], tris[3 * t + 1], tris[3 * t + 2]};
        resolution_ = resolution;   // adjacency: built on first use (see adj())
    }
    // Point is standard layout {X, Y, Z}: the vertex array IS the flat [nv][3] coordinate array the C ABI takes
    const double* coord_data() const { static_assert(sizeof(Point) == 3 * sizeof(double), "Point must be three packed doubles"); return V().empty() ? nullptr : &V()[0].X; }
    std::vector<double> coords() const {
        std::vector<double> o((size_t)nvertices() * 3);
        for (int i = 0; i < nvertices(); i++) { o[3 * i] = V()[i].X; o[3 * i + 1] = V()[i].Y; o[3 * i + 2] = V()[i].Z; }
        return o;
    }
    std::vector<int> faces() const {
        std::vector<int> o((size_t)ntriangles() * 3);
        for (int t = 0; t < ntriangles(); t++) for (int j = 0; j < 3; j++) o[3 * t + j] = topo_->F[t][j];
        return o;
    }
    // adjacency order: faces ascending; neighbours by first appearance over those faces.  Built on first use and shared by all
    // copies of the mesh: the data-sized meshes that only pass through resampling / point location never need it
    // (56 ms for an ico7 mesh).
    void build_adjacency() { (void)adj(); }
private:
    // Coordinates are copy-on-write: the drivers copy meshes freely (reset_meshspace / reset_CPgrid / get_CPgrid take and
    // return meshes by value, src/DiscreteModel.h:108-119), a copy costs nothing until one side moves a vertex.  (Not for
    // concurrent mutation of copies of one mesh — nothing in the path does that.)
    std::shared_ptr<std::vector<Point>> V_ = std::make_shared<std::vector<Point>>();
    const std::vector<Point>& V() const { return *V_; }
    std::vector<Point>& Vw() {
        if (V_.use_count() > 1) V_ = std::make_shared<std::vector<Point>>(*V_);
        return *V_;
    }
    // topology is immutable once set and shared between copies: the drivers copy meshes freely
    // (reset_meshspace / reset_CPgrid take meshes by value semantics), only coordinates ever change
    struct Topo {
        std::vector<std::array<int, 3>> F;
        mutable std::vector<std::vector<int>> nbr, vtri;
        mutable std::once_flag adjacency_once;
    };
    const Topo& adj() const {
        const Topo& T = *topo_;
        const size_t nv = V().size();
        std::call_once(T.adjacency_once, [&T, nv] {
            T.nbr.assign(nv, {}); T.vtri.assign(nv, {});
            for (int t = 0; t < (int)T.F.size(); t++) for (int j = 0; j < 3; j++) T.vtri[T.F[t][j]].push_back(t);
            for (int v = 0; v < (int)nv; v++)
                for (int t : T.vtri[v])
                    for (int j = 0; j < 3; j++) {
                        const int u = T.F[t][j];
                        if (u != v && std::find(T.nbr[v].begin(), T.nbr[v].end(), u) == T.nbr[v].end()) T.nbr[v].push_back(u);
                    }
        });
        return T;
    }
    std::shared_ptr<Topo> topo_ = std::make_shared<Topo>();
    std::shared_ptr<NEWMAT::Matrix> pv_;
    int resolution_ = -1;
};

// make_mesh_from_icosa(k): k-times 4-split icosahedron, unit radius, nested numbering, child faces 4f..4f+3.
inline Mesh make_mesh_from_icosa_uncached(int level);
// Every level of a registration asks for the same few icospheres again (control grid, sampling grid, data template): built once
// per level and handed out as copies (topology shared, coordinates copy-on-write) — ico7 costs 115 ms to build.
inline Mesh make_mesh_from_icosa(int level) {
    static std::mutex mtx;
    static std::map<int, Mesh> cache;
    std::lock_guard<std::mutex> lock(mtx);
    auto it = cache.find(level);
    if (it == cache.end()) it = cache.emplace(level, make_mesh_from_icosa_uncached(level)).first;
    return it->second;
}
inline Mesh make_mesh_from_icosa_uncached(int level) {
    const double t = (1.0 + std::sqrt(5.0)) / 2.0;
    const double base[12][3] = {{-1, t, 0}, {1, t, 0}, {-1, -t, 0}, {1, -t, 0}, {0, -1, t}, {0, 1, t},
                                {0, -1, -t}, {0, 1, -t}, {t, 0, -1}, {t, 0, 1}, {-t, 0, -1}, {-t, 0, 1}};
    const int faces[20][3] = {{0, 11, 5}, {0, 5, 1},  {0, 1, 7},   {0, 7, 10}, {0, 10, 11}, {1, 5, 9}, {5, 11, 4},
                              {11, 10, 2}, {10, 7, 6}, {7, 1, 8},  {3, 9, 4},  {3, 4, 2},   {3, 2, 6}, {3, 6, 8},
                              {3, 8, 9},  {4, 9, 5},  {2, 4, 11},  {6, 2, 10}, {8, 6, 7},   {9, 8, 1}};
    std::vector<Point> V;
    std::vector<std::array<int, 3>> F;
    for (auto& b : base) { Point p(b[0], b[1], b[2]); p.normalize(); V.push_back(p); }
    for (auto& f : faces) F.push_back({f[0], f[1], f[2]});
    for (int l = 0; l < level; l++) {
        std::map<std::pair<int, int>, int> mid;
        auto midpoint = [&](int a, int b) {
            std::pair<int, int> key(std::min(a, b), std::max(a, b));
            auto it = mid.find(key);
            if (it != mid.end()) return it->second;
            Point p = V[a] + V[b];
            p.normalize();
            V.push_back(p);
            return mid[key] = (int)V.size() - 1;
        };
        std::vector<std::array<int, 3>> NF;
        NF.reserve(F.size() * 4);
        for (auto& f : F) {
            int a = f[0], b = f[1], c = f[2];
            int ab = midpoint(a, b), bc = midpoint(b, c), ca = midpoint(c, a);
            NF.push_back({a, ab, ca}); NF.push_back({ab, b, bc}); NF.push_back({ca, bc, c}); NF.push_back({ab, bc, ca});
        }
        F.swap(NF);
    }
    Mesh M;
    std::vector<double> xyz(V.size() * 3);
    std::vector<int> tr(F.size() * 3);
    for (size_t i = 0; i < V.size(); i++) { xyz[3 * i] = V[i].X; xyz[3 * i + 1] = V[i].Y; xyz[3 * i + 2] = V[i].Z; }
    for (size_t i = 0; i < F.size(); i++) for (int j = 0; j < 3; j++) tr[3 * i + j] = F[i][j];
    M.set(xyz.data(), (int)V.size(), tr.data(), (int)F.size(), level);
    return M;
}
inline void true_rescale(Mesh& m, double rad) {
    for (int i = 0; i < m.nvertices(); i++) { Point p = m.get_coord(i); p.normalize(); m.set_coord(i, p * rad); }
}
inline void recentre(Mesh& m) {
    Point c;
    for (int i = 0; i < m.nvertices(); i++) c = c + m.get_coord(i);
    c = c * (1.0 / std::max(1, m.nvertices()));
    for (int i = 0; i < m.nvertices(); i++) m.set_coord(i, m.get_coord(i) - c);
}

// The Octree of the reference (built at src/DiscreteModel.cpp:88) is replaced by the device-side icosahedral hierarchy;
// the type only survives as an opaque token in the setter signatures.
class Octree {
public:
    Octree() = default;
    explicit Octree(const Mesh&) {}
};

}  // namespace newresampler
