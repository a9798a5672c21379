// fsl_ext.h — TEST INFRASTRUCTURE.  Stand-ins for the THIRD-PARTY libraries the reference links but does not vendor
// (src/Makefile:5: -lfsl-newresampler -lfsl-miscmaths -lfsl-newmat(armawrap) ..., plus boost::variant), so that the
// reference's OWN cost-function sources — src/DiscreteCostFunction.cpp, src/similarities.cc, src/reg_tools.cpp,
// src/DiscreteModel.cpp ... — compile FROM WHERE THEY LIE under /root/reference into oracle/_ref/libref_cost.so
// (recipe: oracle/Makefile, target `refcost`).  Nothing of the reference is copied: this file only supplies the
// external API surface those sources are written against.
//
// What is restated here is the published behaviour of FSL 6.0.7 fsl-newresampler / fsl-miscmaths / armawrap-NEWMAT
// (SURVEY.md §0.3 A1-A9); every geometric primitive forwards to oracle/geom.h so that the [EXT] semantics exist
// exactly once in this repository.  Only `tests/`, `__graft_entry__.smoke()` and bench.py's CPU legs may use it.
#pragma once
// The FSL headers the reference includes reach the C headers <math.h> / <stdlib.h> (NIfTI/GIfTI I/O, armadillo), whose
// libstdc++ wrappers put the floating-point overloads of abs() into the global namespace.  The reference relies on that:
// src/DiscreteModel.cpp:148 calls an UNQUALIFIED abs() on a double (the direction de-duplication of the barycentre label
// set).  With only <cmath>/<cstdlib> the call would bind to ::abs(int), truncate its argument and leave 4 labels instead
// of 19 — so the same two headers are included here, first.  (An assumption about [EXT]; recorded in DESIGN.md.)
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <variant>
#include <vector>

#include <functional>

#include "../geom.h"
#include "../feat.h"

// ---------------------------------------------------------------------------------------------------------------
// boost::variant / boost::get over std::variant (src/reg_tools.h:7-8,12-13)
// ---------------------------------------------------------------------------------------------------------------
namespace boost {
template <typename... Ts> using variant = std::variant<Ts...>;
template <typename T, typename... Ts> const T& get(const std::variant<Ts...>& v) { return std::get<T>(v); }
template <typename T, typename... Ts> T& get(std::variant<Ts...>& v) { return std::get<T>(v); }
}  // namespace boost

// ---------------------------------------------------------------------------------------------------------------
// NEWMAT (A9): 1-based dense matrices, the operations the path's sources use.
// ---------------------------------------------------------------------------------------------------------------
namespace NEWMAT {

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

    Matrix t() const {
        Matrix o(c_, r_);
        for (int i = 1; i <= r_; i++) for (int j = 1; j <= c_; j++) o(j, i) = (*this)(i, j);
        return o;
    }
    double Trace() const { double s = 0; for (int i = 1; i <= std::min(r_, c_); i++) s += (*this)(i, i); return s; }
    double Sum() const { double s = 0; for (double v : d_) s += v; return s; }
    // LU with partial pivoting (NEWMAT's CroutMatrix does the same; rounding may differ in the last ulp)
    double Determinant() const {
        if (r_ != c_) throw std::runtime_error("Determinant of a non-square matrix");
        Matrix a = *this; double det = 1;
        for (int k = 1; k <= r_; k++) {
            int p = k;
            for (int i = k + 1; i <= r_; i++) if (std::fabs(a(i, k)) > std::fabs(a(p, k))) p = i;
            if (a(p, k) == 0) return 0;
            if (p != k) { for (int j = 1; j <= r_; j++) std::swap(a(p, j), a(k, j)); det = -det; }
            det *= a(k, k);
            for (int i = k + 1; i <= r_; i++) { double f = a(i, k) / a(k, k); for (int j = k; j <= r_; j++) a(i, j) -= f * a(k, j); }
        }
        return det;
    }
    Matrix i() const {
        if (r_ != c_) throw std::runtime_error("inverse of a non-square matrix");
        const int n = r_;
        Matrix a = *this, inv(n, n);
        for (int k = 1; k <= n; k++) inv(k, k) = 1;
        for (int k = 1; k <= n; k++) {
            int p = k;
            for (int q = k + 1; q <= n; q++) if (std::fabs(a(q, k)) > std::fabs(a(p, k))) p = q;
            if (a(p, k) == 0) throw std::runtime_error("matrix is singular");
            if (p != k) for (int j = 1; j <= n; j++) { std::swap(a(p, j), a(k, j)); std::swap(inv(p, j), inv(k, j)); }
            const double d = a(k, k);
            for (int j = 1; j <= n; j++) { a(k, j) /= d; inv(k, j) /= d; }
            for (int q = 1; q <= n; q++) {
                if (q == k) continue;
                const double f = a(q, k);
                if (f == 0) continue;
                for (int j = 1; j <= n; j++) { a(q, j) -= f * a(k, j); inv(q, j) -= f * inv(k, j); }
            }
        }
        return inv;
    }
    Matrix SubMatrix(int r0, int r1, int c0, int c1) const {
        Matrix o(r1 - r0 + 1, c1 - c0 + 1);
        for (int i = r0; i <= r1; i++) for (int j = c0; j <= c1; j++) o(i - r0 + 1, j - c0 + 1) = (*this)(i, j);
        return o;
    }
    Matrix Row(int i) const { return SubMatrix(i, i, 1, c_); }
    Matrix Column(int j) const { return SubMatrix(1, r_, j, j); }
protected:
    int r_ = 0, c_ = 0;
    std::vector<double> d_;
};
inline Matrix operator*(const Matrix& a, const Matrix& b) {
    if (a.Ncols() != b.Nrows()) throw std::runtime_error("NEWMAT: incompatible dimensions in product");
    Matrix o(a.Nrows(), b.Ncols());
    for (int i = 1; i <= a.Nrows(); i++)
        for (int j = 1; j <= b.Ncols(); j++) {
            double s = 0;
            for (int k = 1; k <= a.Ncols(); k++) s += a(i, k) * b(k, j);
            o(i, j) = s;
        }
    return o;
}
inline Matrix operator+(const Matrix& a, const Matrix& b) {
    Matrix o(a.Nrows(), a.Ncols());
    for (int i = 1; i <= a.Nrows(); i++) for (int j = 1; j <= a.Ncols(); j++) o(i, j) = a(i, j) + b(i, j);
    return o;
}
inline Matrix operator-(const Matrix& a, const Matrix& b) {
    Matrix o(a.Nrows(), a.Ncols());
    for (int i = 1; i <= a.Nrows(); i++) for (int j = 1; j <= a.Ncols(); j++) o(i, j) = a(i, j) - b(i, j);
    return o;
}
inline Matrix operator*(const Matrix& a, double s) {
    Matrix o(a.Nrows(), a.Ncols());
    for (int i = 1; i <= a.Nrows(); i++) for (int j = 1; j <= a.Ncols(); j++) o(i, j) = a(i, j) * s;
    return o;
}
inline Matrix operator*(double s, const Matrix& a) { return a * s; }
inline Matrix operator/(const Matrix& a, double s) { return a * (1.0 / s); }

class ColumnVector : public Matrix {
public:
    ColumnVector() = default;
    explicit ColumnVector(int n) : Matrix(n, 1) {}
    ColumnVector(const Matrix& m) : Matrix(m) {
        if (m.Ncols() != 1 && m.Nrows() * m.Ncols() != 0) throw std::runtime_error("NEWMAT: not a column vector");
    }
    void ReSize(int n) { Matrix::ReSize(n, 1); }
    void resize(int n) { ReSize(n); }
    using Matrix::operator();
    double& operator()(int i) { return d_[i - 1]; }
    double operator()(int i) const { return d_[i - 1]; }
    ColumnVector& operator=(double v) { std::fill(d_.begin(), d_.end(), v); return *this; }
};
class RowVector : public Matrix {
public:
    RowVector() = default;
    explicit RowVector(int n) : Matrix(1, n) {}
    RowVector(const Matrix& m) : Matrix(m) {}
    void ReSize(int n) { Matrix::ReSize(1, n); }
    using Matrix::operator();
    double& operator()(int i) { return d_[i - 1]; }
    double operator()(int i) const { return d_[i - 1]; }
    RowVector& operator=(double v) { std::fill(d_.begin(), d_.end(), v); return *this; }
};
class DiagonalMatrix : public Matrix {
public:
    DiagonalMatrix() = default;
    explicit DiagonalMatrix(int n) : Matrix(n, n) {}
    void ReSize(int n) { Matrix::ReSize(n, n); }
    using Matrix::operator();
    double& operator()(int i) { return Matrix::operator()(i, i); }
    double operator()(int i) const { return Matrix::operator()(i, i); }
    DiagonalMatrix& operator=(double v) { for (int i = 1; i <= r_; i++) (*this)(i) = v; return *this; }
};
// Singular value decomposition A = U D V' is only reached from the post-hoc strain report (src/reg_tools.cpp:370-
// 556, calculate_strains), which is outside the cost-evaluation path: declared so the file compiles, loud if called.
inline void SVD(const Matrix&, DiagonalMatrix&, Matrix&, Matrix&) { throw std::runtime_error("fsl_ext: NEWMAT::SVD is not provided"); }
inline void SVD(const Matrix&, DiagonalMatrix&, Matrix&) { throw std::runtime_error("fsl_ext: NEWMAT::SVD is not provided"); }
inline void SVD(const Matrix&, DiagonalMatrix&) { throw std::runtime_error("fsl_ext: NEWMAT::SVD is not provided"); }

}  // namespace NEWMAT

// ---------------------------------------------------------------------------------------------------------------
// MISCMATHS: pow, BFMatrix family, SpMat, Histogram (only pow and the dense BFMatrix are live on the path)
// ---------------------------------------------------------------------------------------------------------------
namespace MISCMATHS {

inline double pow(double x, double y) { return std::pow(x, y); }
inline double pow(double x, float y) { return std::pow(x, (double)y); }
inline double pow(double x, int y) { return std::pow(x, (double)y); }
inline float pow(float x, float y) { return std::pow(x, y); }

template <typename T>
class SpMat {
public:
    SpMat() = default;
    SpMat(unsigned int m, unsigned int n) : m_(m), n_(n) {}
    explicit SpMat(const std::string&) { throw std::runtime_error("fsl_ext: sparse matrix file I/O is not provided"); }
    unsigned int Nrows() const { return m_; }
    unsigned int Ncols() const { return n_; }
    T Peek(unsigned int r, unsigned int c) const { auto it = v_.find({r, c}); return it == v_.end() ? T(0) : it->second; }
    void Set(unsigned int r, unsigned int c, const T& v) { v_[{r, c}] = v; }
    T& Set(unsigned int r, unsigned int c) { return v_[{r, c}]; }
    void AddTo(unsigned int r, unsigned int c, const T& v) { v_[{r, c}] += v; }
private:
    unsigned int m_ = 0, n_ = 0;
    std::map<std::pair<unsigned int, unsigned int>, T> v_;
};

class BFMatrix;
class BFMatrixColumnIterator {  // walks the rows of one column (1-based)
public:
    BFMatrixColumnIterator(const BFMatrix* m, unsigned int col, unsigned int row) : m_(m), col_(col), row_(row) {}
    double operator*() const;
    BFMatrixColumnIterator& operator++() { row_++; return *this; }
    BFMatrixColumnIterator operator++(int) { BFMatrixColumnIterator t = *this; row_++; return t; }
    bool operator==(const BFMatrixColumnIterator& o) const { return row_ == o.row_ && col_ == o.col_; }
    bool operator!=(const BFMatrixColumnIterator& o) const { return !(*this == o); }
    unsigned int Row() const { return row_; }
private:
    const BFMatrix* m_; unsigned int col_, row_;
};
class BFMatrix {  // dense storage; rows = feature dimensions, columns = vertices (src/featurespace.h:30-37)
public:
    BFMatrix() = default;
    BFMatrix(unsigned int m, unsigned int n) : M_(m, n) {}
    explicit BFMatrix(const NEWMAT::Matrix& M) : M_(M) {}
    virtual ~BFMatrix() = default;
    virtual unsigned int Nrows() const { return (unsigned int)M_.Nrows(); }
    virtual unsigned int Ncols() const { return (unsigned int)M_.Ncols(); }
    virtual double Peek(unsigned int r, unsigned int c) const { return M_((int)r, (int)c); }
    virtual void Set(unsigned int r, unsigned int c, double v) { M_((int)r, (int)c) = v; }
    virtual void Resize(unsigned int m, unsigned int n) { M_.ReSize((int)m, (int)n); }
    virtual NEWMAT::Matrix AsMatrix() const { return M_; }
    virtual std::shared_ptr<BFMatrix> Transpose() const { return std::make_shared<BFMatrix>(M_.t()); }
    BFMatrixColumnIterator begin(unsigned int col) const { return BFMatrixColumnIterator(this, col, 1); }
    BFMatrixColumnIterator end(unsigned int col) const { return BFMatrixColumnIterator(this, col, Nrows() + 1); }
protected:
    NEWMAT::Matrix M_;
};
inline double BFMatrixColumnIterator::operator*() const { return m_->Peek(row_, col_); }
class FullBFMatrix : public BFMatrix {
public:
    using BFMatrix::BFMatrix;
    FullBFMatrix() = default;
};
template <typename T>
class SparseBFMatrix : public BFMatrix {  // never instantiated here: dynamic_pointer_cast to it yields null
public:
    using BFMatrix::BFMatrix;
    SparseBFMatrix() = default;
    explicit SparseBFMatrix(const SpMat<T>&) { throw std::runtime_error("fsl_ext: SparseBFMatrix is not provided"); }
};

// A10: MISCMATHS::Histogram as restated in oracle/feat.h (orc::Histogram) — the calls of src/reg_tools.cpp:792-806
class Histogram {
public:
    Histogram() = default;
    Histogram(const NEWMAT::ColumnVector& data, int bins) : h_(std::make_shared<orc::Histogram>(vec(data), bins)) {}
    void generate() { h_->generate(std::vector<double>(h_->source.size(), 1.0)); }
    void generate(const NEWMAT::ColumnVector& exclude) { h_->generate(vec(exclude)); }
    void generateCDF() { h_->generate_cdf(); }
    void setexclusion(const NEWMAT::ColumnVector& exclude) { h_->excl = vec(exclude); }
    void match(Histogram& target) { h_->match(*target.h_); }
    NEWMAT::ColumnVector getsourceData() const {
        NEWMAT::ColumnVector v((int)h_->source.size());
        for (int i = 1; i <= v.Nrows(); i++) v(i) = h_->source[i - 1];
        return v;
    }
private:
    static std::vector<double> vec(const NEWMAT::ColumnVector& c) {
        std::vector<double> v(c.Nrows());
        for (int i = 1; i <= c.Nrows(); i++) v[i - 1] = c(i);
        return v;
    }
    std::shared_ptr<orc::Histogram> h_;
};

}  // namespace MISCMATHS

// ---------------------------------------------------------------------------------------------------------------
// newresampler (A1-A8): Point / Triangle / Mesh / Octree and the free functions, over oracle/geom.h
// ---------------------------------------------------------------------------------------------------------------
namespace newresampler {

// A1.  A class of its own (not an alias of orc::Pt): the reference calls estimate_rotation_matrix, project_point,
// form_matrix_from_points ... unqualified and relies on argument-dependent lookup into namespace newresampler.
class Point {
public:
    double X = 0, Y = 0, Z = 0;
    Point() = default;
    Point(double x, double y, double z) : X(x), Y(y), Z(z) {}
    Point(const orc::Pt& p) : X(p.X), Y(p.Y), Z(p.Z) {}
    operator orc::Pt() const { return orc::Pt(X, Y, Z); }
    double norm() const { return std::sqrt(X * X + Y * Y + Z * Z); }
    void normalize() { double n = norm(); if (n != 0) { X /= n; Y /= n; Z /= n; } }
    Point& operator+=(const Point& b) { X += b.X; Y += b.Y; Z += b.Z; return *this; }
    Point& operator-=(const Point& b) { X -= b.X; Y -= b.Y; Z -= b.Z; return *this; }
    Point& operator*=(double s) { X *= s; Y *= s; Z *= s; return *this; }
    Point& operator/=(double s) { X /= s; Y /= s; Z /= s; return *this; }
};
inline Point operator+(const Point& a, const Point& b) { return {a.X + b.X, a.Y + b.Y, a.Z + b.Z}; }
inline Point operator-(const Point& a, const Point& b) { return {a.X - b.X, a.Y - b.Y, a.Z - b.Z}; }
inline Point operator*(const Point& a, double s) { return {a.X * s, a.Y * s, a.Z * s}; }
inline Point operator*(double s, const Point& a) { return a * s; }
inline Point operator/(const Point& a, double s) { return {a.X / s, a.Y / s, a.Z / s}; }
inline double operator|(const Point& a, const Point& b) { return a.X * b.X + a.Y * b.Y + a.Z * b.Z; }  // dot
inline Point operator*(const Point& a, const Point& b) { return {a.Y * b.Z - a.Z * b.Y, a.Z * b.X - a.X * b.Z, a.X * b.Y - a.Y * b.X}; }  // cross
inline bool operator==(const Point& a, const Point& b) { return a.X == b.X && a.Y == b.Y && a.Z == b.Z; }
inline bool operator!=(const Point& a, const Point& b) { return !(a == b); }
inline Point operator*(const NEWMAT::Matrix& M, const Point& p) {  // `R * Point` (src/DiscreteCostFunction.cpp:151,419)
    return {M(1, 1) * p.X + M(1, 2) * p.Y + M(1, 3) * p.Z, M(2, 1) * p.X + M(2, 2) * p.Y + M(2, 3) * p.Z,
            M(3, 1) * p.X + M(3, 2) * p.Y + M(3, 3) * p.Z};
}
inline Point operator*(const Point& p, const NEWMAT::Matrix& M) {  // row vector times matrix (src/DiscreteGroupModel.cpp:89)
    return {p.X * M(1, 1) + p.Y * M(2, 1) + p.Z * M(3, 1), p.X * M(1, 2) + p.Y * M(2, 2) + p.Z * M(3, 2),
            p.X * M(1, 3) + p.Y * M(2, 3) + p.Z * M(3, 3)};
}
inline std::ostream& operator<<(std::ostream& o, const Point& p) { return o << p.X << " " << p.Y << " " << p.Z; }

class MeshException : public std::exception {
public:
    const char* errmesg;
    explicit MeshException(const char* msg) : errmesg(msg) {}
    const char* what() const noexcept override { return errmesg; }
};

inline NEWMAT::Matrix to_newmat_fwd(const orc::Mat3& R) {
    NEWMAT::Matrix M(3, 3);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) M(i + 1, j + 1) = R.m[i][j];
    return M;
}
struct Tangs { Point e1, e2; };  // tangent-plane basis (src/reg_tools.cpp:209-368)
// coordinates of p in a tangent basis (post-hoc strain report only, src/reg_tools.cpp:392)
inline void project_point(const Point& p, const Tangs& T, double& y1, double& y2) { y1 = p | T.e1; y2 = p | T.e2; }
// A6: the three points as COLUMNS of a 3x3 matrix (rows when `trans`)
inline NEWMAT::Matrix form_matrix_from_points(const Point& p1, const Point& p2, const Point& p3, bool trans = false) {
    NEWMAT::Matrix M = to_newmat_fwd(orc::form_matrix_from_points(p1, p2, p3));
    return trans ? M.t() : M;
}

inline NEWMAT::Matrix to_newmat(const orc::Mat3& R) { return to_newmat_fwd(R); }
// A2: minimal rotation taking direction p1 onto p2 (oracle/geom.h:89)
inline NEWMAT::Matrix estimate_rotation_matrix(const Point& p1, const Point& p2) { return to_newmat(orc::estimate_rotation_matrix(p1, p2)); }

class Triangle {  // A5
public:
    Triangle() = default;
    Triangle(const Point& a, const Point& b, const Point& c, int no) : t_(orc::Pt(a), orc::Pt(b), orc::Pt(c), no) {}
    void set(const Point& a, const Point& b, const Point& c, int no) { t_ = orc::Tri(orc::Pt(a), orc::Pt(b), orc::Pt(c), no); }
    Point get_vertex_coord(int i) const { return Point(t_.v[i]); }
    void set_vertex(int i, const Point& p) { t_.v[i] = p; }
    int get_vertex_no(int i) const { return t_.n[i]; }
    void set_vertex_no(int i, int n) { t_.n[i] = n; }
    int get_no() const { return t_.id; }
    Point normal() const { return Point(t_.normal()); }
    double get_area() const { return t_.area(); }
    std::vector<double> get_angles() const { auto a = t_.angles(); return {a[0], a[1], a[2]}; }
    Point centroid() const { return Point((t_.v[0] + t_.v[1] + t_.v[2]) / 3.0); }
    const orc::Tri& tri() const { return t_; }
private:
    orc::Tri t_;
};

class Mesh {  // A7 (container + adjacency + per-vertex values; no file I/O)
public:
    Mesh() = default;
    explicit Mesh(const orc::Mesh& m) : m_(std::make_shared<orc::Mesh>(m)), V_(m.V.begin(), m.V.end()) {}
    int nvertices() const { return (int)V_.size(); }
    int ntriangles() const { return m_ ? (int)m_->F.size() : 0; }
    const Point& get_coord(int i) const { return V_[i]; }
    void set_coord(int i, const Point& p) { V_[i] = p; }
    int get_triangle_vertexID(int t, int j) const { return m_->F[t][j]; }
    Point get_triangle_vertex(int t, int j) const { return V_[m_->F[t][j]]; }
    Triangle get_triangle(int t) const {
        Triangle T(V_[m_->F[t][0]], V_[m_->F[t][1]], V_[m_->F[t][2]], t);
        for (int j = 0; j < 3; j++) T.set_vertex_no(j, m_->F[t][j]);
        return T;
    }
    Triangle get_triangle_from_vertex(int v, int k) const { return get_triangle(m_->vtri[v][k]); }
    Point get_triangle_normal(int t) const { return get_triangle(t).normal(); }
    // vertex normal: normalised sum of the unit normals of the incident faces
    Point get_normal(int v) const {
        Point n;
        for (int t : m_->vtri[v]) n = n + get_triangle(t).normal();
        n.normalize();
        return n;
    }
    Point local_normal(int v) const { return get_normal(v); }
    std::vector<int>::const_iterator nbegin(int i) const { return m_->nbr[i].begin(); }
    std::vector<int>::const_iterator nend(int i) const { return m_->nbr[i].end(); }
    std::vector<int>::const_iterator tIDbegin(int i) const { return m_->vtri[i].begin(); }
    std::vector<int>::const_iterator tIDend(int i) const { return m_->vtri[i].end(); }
    int get_total_neighbours(int i) const { return (int)m_->nbr[i].size(); }
    int get_total_triangles(int i) const { return (int)m_->vtri[i].size(); }
    double calculate_MaxVD() const { return synced().calculate_MaxVD(); }
    double calculate_MeanVD() const { return synced().calculate_MeanVD(); }
    std::vector<std::vector<double>> get_face_angles() const {
        std::vector<std::vector<double>> a(ntriangles());
        for (int t = 0; t < ntriangles(); t++) a[t] = get_triangle(t).get_angles();
        return a;
    }
    // per-vertex values ("pvalues": one row per dimension)
    int npvalues() const { return nvertices(); }
    int get_dimension() const { return pv_.Nrows(); }
    double get_pvalue(int i, int dim = 0) const { return pv_(dim + 1, i + 1); }
    void set_pvalue(int i, double v, int dim = 0) { ensure_pv(dim + 1); pv_(dim + 1, i + 1) = v; }
    void set_pvalues(const NEWMAT::Matrix& M) { pv_ = M; }
    NEWMAT::Matrix get_pvalues() const { return pv_; }
    void save(const std::string&) const {}
    // no file I/O: "loading" asks the in-memory registry a test harness may have installed (oracle/ref_feat_driver.cpp)
    static std::function<void(Mesh&, const std::string&)>& loader() { static std::function<void(Mesh&, const std::string&)> f; return f; }
    void load(const std::string& name, bool = true, bool = true) {
        if (!loader()) throw MeshException("fsl_ext: mesh file I/O is not provided");
        loader()(*this, name);
    }
    // geometry snapshot with the current coordinates (topology shared)
    orc::Mesh synced() const { orc::Mesh o = *m_; o.V.assign(V_.begin(), V_.end()); return o; }
    const orc::Mesh& topo() const { return *m_; }
private:
    void ensure_pv(int rows) {
        if (pv_.Nrows() >= rows && pv_.Ncols() == nvertices()) return;
        NEWMAT::Matrix n(std::max(rows, pv_.Nrows()), nvertices());
        for (int i = 1; i <= pv_.Nrows(); i++) for (int j = 1; j <= std::min(pv_.Ncols(), nvertices()); j++) n(i, j) = pv_(i, j);
        pv_ = n;
    }
    std::shared_ptr<const orc::Mesh> m_;  // topology + adjacency (immutable, shared between copies)
    std::vector<Point> V_;
    NEWMAT::Matrix pv_;
};

inline Mesh make_mesh_from_icosa(int level) { return Mesh(orc::make_mesh_from_icosa(level, 1.0)); }
inline void true_rescale(Mesh& m, double rad) {
    for (int i = 0; i < m.nvertices(); i++) { Point p = m.get_coord(i); p.normalize(); m.set_coord(i, p * rad); }
}
inline void recentre(Mesh& m) {
    Point c;
    for (int i = 0; i < m.nvertices(); i++) c = c + m.get_coord(i);
    c = c * (1.0 / std::max(1, m.nvertices()));
    for (int i = 0; i < m.nvertices(); i++) m.set_coord(i, m.get_coord(i) - c);
}

class Octree {  // A3: closest triangle of a closed spherical mesh = the face whose radial cone holds the point
public:
    Octree() = default;
    explicit Octree(const Mesh& M) : geom_(std::make_shared<orc::Mesh>(M.synced())), loc_(std::make_shared<orc::SphereLocator>(*geom_)) {}
    Triangle get_closest_triangle(const Point& p) const {
        const int t = loc_->closest_triangle(p);
        if (t < 0) throw MeshException("Octree: no triangle found");
        const orc::Tri T = geom_->get_triangle(t);
        Triangle o(Point(T.v[0]), Point(T.v[1]), Point(T.v[2]), t);
        for (int j = 0; j < 3; j++) o.set_vertex_no(j, T.n[j]);
        return o;
    }
    int get_closest_vertex_ID(const Point& p) const {
        const Triangle T = get_closest_triangle(p);
        int best = 0; double bd = 1e300;
        for (int j = 0; j < 3; j++) { const double d = (T.get_vertex_coord(j) - p).norm(); if (d < bd) { bd = d; best = j; } }
        return T.get_vertex_no(best);
    }
private:
    std::shared_ptr<orc::Mesh> geom_;
    std::shared_ptr<orc::SphereLocator> loc_;
};

// A4 (oracle/geom.h:153-180)
inline void project_point(const Point& v0, const Point& v1, const Point& v2, Point& p) { p = Point(orc::project_point(p, v0, v1, v2)); }
inline double barycentric_weight(const Point& v0, const Point& v1, const Point& v2, const Point& p, double f0, double f1, double f2) {
    return orc::barycentric_weight(v0, v1, v2, p, f0, f1, f2);
}
inline std::map<int, double> calc_barycentric_weights(const Point& v0, const Point& v1, const Point& v2, const Point& p, int n0, int n1, int n2) {
    double w[3];
    orc::barycentric_weights(v0, v1, v2, p, w);
    return {{n0, w[0]}, {n1, w[1]}, {n2, w[2]}};
}
// position of p (given in triangle v0 v1 v2) carried to the triangle nv0 nv1 nv2 by its barycentric weights
inline Point barycentric(const Point& v0, const Point& v1, const Point& v2, const Point& p, const Point& nv0, const Point& nv1, const Point& nv2) {
    double w[3];
    orc::barycentric_weights(v0, v1, v2, p, w);
    return nv0 * w[0] + nv1 * w[1] + nv2 * w[2];
}

// A6: move every vertex of `up` with the deformation START -> END of a lower-resolution mesh; defined in
// oracle/ref_cost.cpp over orc::barycentric_mesh_interpolation (oracle/oracle.cpp)
void barycentric_mesh_interpolation(Mesh& up, const Mesh& START, const Mesh& END, int nthreads = 1);
// A8: metric_resample(in, ref): per-vertex values of `in` resampled at the vertices of `ref` with FSL's ADAPTIVE
// barycentric kernel, restated in oracle/geom.h (adaptive_barycentric_weights) from the published algorithm it ports.
Mesh metric_resample(const Mesh& in, const Mesh& ref, int nthreads = 1, std::shared_ptr<Mesh> EXCL = nullptr);
// A9 / A11 (oracle/feat.h): Gaussian smoothing of per-vertex data, exclusion mask from a value band
Mesh smooth_data(const Mesh& in, const Mesh& ref, double sigma, int nthreads = 1, std::shared_ptr<Mesh> EXCL = nullptr);
Mesh create_exclusion(const Mesh& data_mesh, float lower, float upper);
// project_mesh is only reached with anatomical meshes (out of scope): loud if called.
inline Mesh project_mesh(const Mesh&, const Mesh&, const Mesh&, int = 1) { throw MeshException("fsl_ext: project_mesh is not provided"); }

}  // namespace newresampler
