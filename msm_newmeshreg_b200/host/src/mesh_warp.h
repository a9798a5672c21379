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
