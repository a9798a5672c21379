// mesh_warp.h — the mesh-warp step either side of the cost evaluation, host-side mirror:
//   newresampler::barycentric_mesh_interpolation(SPH_up, SPH_low_init, SPH_low_final, threads)   [EXT] A8
//       (call sites src/mesh_registration.cpp:390,305-306,222; src/DiscreteModel.h:123) -> GPU (msm_mesh_interpolate)
//   newmeshreg::unfold(mesh, verbosity)   src/reg_tools.cpp:134-181 (+ helpers :62-132) -> host: the reference's sweep
//       updates vertices in place one after another (each fix sees the previous ones), i.e. it is sequential by
//       definition; only the fold DETECTION is data parallel, and meshes are almost never folded.
#pragma once
#include <cstdlib>
#include <mutex>

#include "../../../include/msm_b200.h"
#include "reg_tools.h"

namespace newresampler {

// process-wide utility context for stateless geometry calls
inline msm_ctx* utility_context() {
    static msm_ctx* ctx = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        int dev = 0;
        if (const char* e = std::getenv("MSM_DEVICE")) dev = std::atoi(e);
        if (msm_create(dev, &ctx) != 0) ctx = nullptr;
    });
    if (!ctx) throw newmeshreg::MeshregException("barycentric_mesh_interpolation: no CUDA device (no CPU fallback)");
    return ctx;
}

inline void barycentric_mesh_interpolation(Mesh& SPH_up, const Mesh& SPH_low_init, const Mesh& SPH_low_final, int /*numthreads*/ = 1) {
    static std::mutex mtx;
    std::lock_guard<std::mutex> lock(mtx);
    msm_ctx* ctx = utility_context();
    auto pts = SPH_up.coords();
    auto from = SPH_low_init.coords(); auto tris = SPH_low_init.faces();
    auto to = SPH_low_final.coords();
    if (msm_mesh_interpolate(ctx, pts.data(), SPH_up.nvertices(), from.data(), SPH_low_init.nvertices(), tris.data(), SPH_low_init.ntriangles(),
                             to.data(), 100.0, pts.data(), nullptr, nullptr) != 0) {
        static thread_local std::string keep;
        keep = msm_last_error(ctx);
        throw newmeshreg::MeshregException(keep.c_str());
    }
    for (int i = 0; i < SPH_up.nvertices(); i++) SPH_up.set_coord(i, Point(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
}


// Point location in an arbitrary (deformed, even torn) spherical triangle mesh on the GPU — Octree::get_closest_triangle
// + barycentric weights ([EXT] A3/A4) for every query at once: tri[i] (-1 when the mesh has no faces) and w[3i..3i+2] in
// that triangle's vertex order.
inline void locate_in_mesh(const double* pts, int M, const Mesh& mesh, std::vector<int>& tri, std::vector<double>& w) {
    tri.assign(M, -1); w.assign((size_t)M * 3, 0.0);
    if (M == 0 || mesh.ntriangles() == 0) return;
    static std::mutex mtx;
    std::lock_guard<std::mutex> lock(mtx);
    msm_ctx* ctx = utility_context();
    auto xyz = mesh.coords(); auto tris = mesh.faces();
    std::vector<double> out((size_t)M * 3);
    if (msm_mesh_interpolate(ctx, pts, M, xyz.data(), mesh.nvertices(), tris.data(), mesh.ntriangles(), xyz.data(), 100.0, out.data(), tri.data(),
                             w.data()) != 0) {
        static thread_local std::string keep;
        keep = msm_last_error(ctx);
        throw newmeshreg::MeshregException(keep.c_str());
    }
}

// newresampler::metric_resample(in, ref) ([EXT] A8; call sites src/DiscreteGroupModel.cpp:92, src/DiscreteCostFunction.cpp:381,
// src/featurespace.cc:46): ADAPTIVE BARYCENTRIC resampling of per-vertex data (values[d*Vin + j]) from mesh `in` onto the
// vertices of `ref`, restated from the published algorithm FSL ports (Connectome Workbench ADAP_BARY_AREA; INTEGRATION.md
// lists it among the assumptions to check against a real FSL build):
//   forward = barycentric weights of every ref vertex in the `in` triangle under it; reverse = those of every in vertex in
//   the ref triangle under it, regrouped per ref vertex; a ref vertex takes the reverse set when it has MORE sources;
//   w *= area_ref[k];  corr[j] = sum_k w;  w *= area_in[j] / corr[j];  normalise per ref vertex.
// The two batches of point locations — all the geometry — run on the GPU; the weight bookkeeping is a linear host pass
// with sources in ascending order (so that sums accumulate as in std::map order).
struct ResampleWeights { std::vector<int> off, src; std::vector<double> w; };
inline std::vector<double> vertex_areas(const Mesh& M) {
    std::vector<double> a(M.nvertices(), 0.0);
    for (int t = 0; t < M.ntriangles(); t++) {
        const double A = M.get_triangle(t).get_area() / 3.0;
        for (int j = 0; j < 3; j++) a[M.get_triangle_vertexID(t, j)] += A;
    }
    return a;
}
inline ResampleWeights adaptive_barycentric_weights(const Mesh& in, const Mesh& ref) {
    const int Vin = in.nvertices(), Vref = ref.nvertices();
    std::vector<int> ft, rt; std::vector<double> fw, rw;
    locate_in_mesh(ref.coord_data(), Vref, in, ft, fw);
    locate_in_mesh(in.coord_data(), Vin, ref, rt, rw);      // a face-less `ref` (bare point set): forward weights only
    const std::vector<double> ai = vertex_areas(in);
    const std::vector<double> ar = ref.ntriangles() > 0 ? vertex_areas(ref) : std::vector<double>(Vref, 1.0);
    // reverse sets per ref vertex: counting sort of (ref vertex <- in vertex) contributions; in vertices ascend within a row
    std::vector<int> roff(Vref + 1, 0);
    for (int j = 0; j < Vin; j++)
        if (rt[j] >= 0) for (int c = 0; c < 3; c++) roff[ref.get_triangle_vertexID(rt[j], c) + 1]++;
    for (int k = 0; k < Vref; k++) roff[k + 1] += roff[k];
    std::vector<int> rsrc(roff[Vref]), fill(roff.begin(), roff.end() - 1);
    std::vector<double> rwt(roff[Vref]);
    for (int j = 0; j < Vin; j++)
        if (rt[j] >= 0)
            for (int c = 0; c < 3; c++) {
                const int k = ref.get_triangle_vertexID(rt[j], c);
                // a degenerate ref triangle may list a vertex twice: the later corner overwrites the earlier (map semantics)
                if (fill[k] > roff[k] && rsrc[fill[k] - 1] == j) { rwt[fill[k] - 1] = rw[3 * j + c]; continue; }
                rsrc[fill[k]] = j; rwt[fill[k]] = rw[3 * j + c]; fill[k]++;
            }
    ResampleWeights R;
    R.off.assign(Vref + 1, 0);
    std::vector<double> corr(Vin, 0.0);
    for (int k = 0; k < Vref; k++) {
        int fs[3]; double fv[3]; int nf = 0;
        if (ft[k] >= 0) {
            for (int c = 0; c < 3; c++) {   // insert ascending, overwriting duplicates
                const int j = in.get_triangle_vertexID(ft[k], c); const double v = fw[3 * k + c];
                int p = 0;
                while (p < nf && fs[p] < j) p++;
                if (p < nf && fs[p] == j) { fv[p] = v; continue; }
                for (int q = nf; q > p; q--) { fs[q] = fs[q - 1]; fv[q] = fv[q - 1]; }
                fs[p] = j; fv[p] = v; nf++;
            }
        }
        const int nr = fill[k] - roff[k];
        if (nr > nf) for (int e = roff[k]; e < fill[k]; e++) { R.src.push_back(rsrc[e]); R.w.push_back(rwt[e]); }
        else for (int e = 0; e < nf; e++) { R.src.push_back(fs[e]); R.w.push_back(fv[e]); }
        R.off[k + 1] = (int)R.src.size();
        for (int e = R.off[k]; e < R.off[k + 1]; e++) { R.w[e] *= ar[k]; corr[R.src[e]] += R.w[e]; }
    }
    for (int k = 0; k < Vref; k++) {
        double sum = 0.0;
        for (int e = R.off[k]; e < R.off[k + 1]; e++) { R.w[e] *= ai[R.src[e]] / corr[R.src[e]]; sum += R.w[e]; }
        if (sum != 0.0) for (int e = R.off[k]; e < R.off[k + 1]; e++) R.w[e] /= sum;
    }
    return R;
}
inline std::vector<double> metric_resample(const Mesh& in, const double* values, int dims, const Mesh& ref) {
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

// The reference's own signature (src/mesh_registration.cpp:326, src/DiscreteGroupModel.cpp:92): values travel as the meshes'
// pvalues ([dims][V]); returns `ref` carrying the resampled rows.
inline Mesh metric_resample(const Mesh& in, const Mesh& ref, int /*nthreads*/ = 1) {
    const NEWMAT::Matrix pv = in.get_pvalues();
    if (pv.Ncols() != in.nvertices()) throw newmeshreg::MeshregException("metric_resample: per-vertex values do not match the mesh");
    const std::vector<double> r = metric_resample(in, pv.data(), pv.Nrows(), ref);
    NEWMAT::Matrix out(pv.Nrows(), ref.nvertices());
    std::copy(r.begin(), r.end(), out.data());
    Mesh o = ref;
    o.set_pvalues(out);
    return o;
}

// metric_resample(in, ref, threads, EXCL) with an exclusion mask over the `in` vertices (src/featurespace.cc:46): excluded sources
// (mask <= 0) drop out of every weight list, the rest is renormalised, a vertex left without sources gets 0.  mask == nullptr:
// the plain call above.
inline std::vector<double> metric_resample_excl(const Mesh& in, const double* values, int dims, const Mesh& ref, const double* mask) {
    if (!mask) return metric_resample(in, values, dims, ref);
    const ResampleWeights R = adaptive_barycentric_weights(in, ref);
    const int Vin = in.nvertices(), Vref = ref.nvertices();
    std::vector<double> out((size_t)dims * Vref, 0.0);
    for (int k = 0; k < Vref; k++) {
        double sum = 0.0;
        for (int e = R.off[k]; e < R.off[k + 1]; e++) if (mask[R.src[e]] > 0.0) sum += R.w[e];
        if (sum == 0.0) continue;
        for (int d = 0; d < dims; d++) {
            double v = 0.0;
            for (int e = R.off[k]; e < R.off[k + 1]; e++) if (mask[R.src[e]] > 0.0) v += (R.w[e] / sum) * values[(size_t)d * Vin + R.src[e]];
            out[(size_t)d * Vref + k] = v;
        }
    }
    return out;
}

// within_controlpt_range for many centres at once (src/DiscreteCostFunction.cpp:104-109, src/DiscreteGroupModel.cpp:88,107):
// for every centre k the points i with 2 RAD asin(|c_k - p_i| / 2 RAD) < range * spacing[k], ascending — the patch
// construction kernels of the cost function (exact at ties), on a context of their own.  CSR: off[ncentres+1], idx.
// new_points: upload `pts` (false: the point set of the previous call is still on the device, e.g. the template once per label).
inline void points_in_range(const double* pts, int npts, bool new_points, const double* centres, int ncentres, const double* spacing, double range,
                            std::vector<int>& off, std::vector<int>& idx) {
    static msm_ctx* ctx = nullptr;
    static std::mutex mtx;
    std::lock_guard<std::mutex> lock(mtx);
    auto check = [&](int st) {
        if (st != 0) {
            static thread_local std::string keep;
            keep = msm_last_error(ctx);
            throw newmeshreg::MeshregException(keep.c_str());
        }
    };
    if (!ctx) {
        int dev = 0;
        if (const char* e = std::getenv("MSM_DEVICE")) dev = std::atoi(e);
        if (msm_create(dev, &ctx) != 0) { ctx = nullptr; throw newmeshreg::MeshregException("points_in_range: no CUDA device (no CPU fallback)"); }
    }
    static int have_npts = -1;
    if (new_points || npts != have_npts) {
        std::vector<double> zeros(npts, 0.0);
        check(msm_set_source(ctx, pts, npts, zeros.data(), 1, nullptr, 0, 0));
        have_npts = npts;
    }
    msm_params p; msm_default_params(&p); p.range = range;
    check(msm_set_params(ctx, &p));
    check(msm_set_cpgrid(ctx, centres, ncentres, nullptr, 0, spacing, nullptr, nullptr));
    check(msm_build_patches(ctx));
    long long total = 0;
    check(msm_patch_count(ctx, &total));
    off.assign(ncentres + 1, 0); idx.assign((size_t)std::max<long long>(total, 1), 0);
    check(msm_get_patches(ctx, off.data(), idx.data()));
    idx.resize((size_t)total);
}

}  // namespace newresampler

namespace newmeshreg {

namespace detail {
using newresampler::Mesh;
using newresampler::Point;

// direction, inside the triangle's plane, perpendicular to edge v0v1 and pointing towards v2's side, scaled by half the
// edge length: the gradient of the triangle's area w.r.t. moving v2 (src/reg_tools.cpp:62-95)
inline Point area_gradient(const Point& v0, const Point& v1, const Point& v2) {
    Point towards = v2 - v0, along = v1 - v0;
    if (towards.norm() > 1e-10) towards.normalize(); else towards = Point(0, 0, 0);
    if (along.norm() > 1e-10) along.normalize(); else along = Point(0, 0, 0);
    Point nrm = towards * along;
    if (nrm.norm() > 1e-10) nrm.normalize(); else nrm = Point(0, 0, 0);
    Point perp = along * nrm;
    if ((towards | perp) < 0) perp = perp * -1;
    return perp * 0.5 * (v1 - v0).norm();
}

inline Point spatialgradient(int index, const Mesh& M) {  // :97-119
    Point grad;
    const Point ci = M.get_coord(index);
    for (auto j = M.tIDbegin(index); j != M.tIDend(index); j++) {
        const Point v0 = M.get_triangle_vertex(*j, 0), v1 = M.get_triangle_vertex(*j, 1), v2 = M.get_triangle_vertex(*j, 2);
        if ((ci - v0).norm() == 0) grad = grad + area_gradient(v1, v2, v0);
        else if ((ci - v1).norm() == 0) grad = grad + area_gradient(v2, v0, v1);
        else grad = grad + area_gradient(v0, v1, v2);
    }
    return grad;
}

inline bool is_folded(int ind, const Mesh& M) {  // check_for_intersections :121-132
    const Point N = M.get_triangle(*M.tIDbegin(ind)).normal();
    for (auto j = M.tIDbegin(ind); j != M.tIDend(ind); j++)
        if ((N | M.get_triangle(*j).normal()) <= 0.5) return true;
    return false;
}
}  // namespace detail

inline void unfold(newresampler::Mesh& SOURCE, bool verbosity = false) {  // :134-181
    using newresampler::Point;
    const double RAD_UNFOLD = 100;  // :139
    for (int sweep = 0; sweep < 1000; sweep++) {
        std::vector<int> bad;
        for (int i = 0; i < SOURCE.nvertices(); i++)
            if (detail::is_folded(i, SOURCE)) bad.push_back(i);
        if (bad.empty()) return;
        if (sweep % 100 == 0 && verbosity) std::cout << "Mesh is folded, total folded vertices: " << bad.size() << " iter: " << sweep << std::endl;
        std::vector<Point> grads;
        grads.reserve(bad.size());
        for (int v : bad) grads.push_back(detail::spatialgradient(v, SOURCE));   // all gradients first, then the moves
        for (size_t i = 0; i < bad.size(); i++) {
            const Point ci = SOURCE.get_coord(bad[i]);
            double step = 1.0;
            Point pp;
            do {  // move against the area gradient, halving the step until the vertex is no longer folded
                pp = ci - grads[i] * step;
                pp.normalize();
                SOURCE.set_coord(bad[i], pp * RAD_UNFOLD);
                step *= 0.5;
            } while (detail::is_folded(bad[i], SOURCE) && step > 1e-3);
        }
    }
}

}  // namespace newmeshreg
