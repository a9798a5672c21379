// meshwarp.cuh — the mesh-warp step on either side of the cost evaluation:
// newresampler::barycentric_mesh_interpolation(SPH_up, SPH_low_init, SPH_low_final) ([EXT] A8; call sites
// src/mesh_registration.cpp:390 (source <- control-grid update), :305-306 (level transition), src/DiscreteModel.h:123).
// For every vertex p of the upsampled mesh: find the triangle of the (deformed, arbitrary) low-resolution START mesh
// whose radial cone contains p, take the barycentric weights of the radially projected point, move p to the same
// combination of the END mesh's vertices and push it back onto the sphere of radius RAD.
//
// The START mesh is NOT a regular icosphere (it is the current, deformed control grid), so the icosahedral hierarchy of
// locate.cuh does not apply; a uniform voxel grid over unit directions gives each query a handful of candidate triangles.
// The grid is built on the device (k_tri_cells count / fill around one scan, cell lists then put in ascending triangle order
// so that exact ties between neighbouring faces are decided the same way every run): the same call also serves the data-mesh
// sized point locations of metric_resample (ico6 / ico7 meshes, millions of cells).
#pragma once
#include <cuda_runtime.h>

namespace msm {

struct WarpGrid {
    int G;                 // cells per dimension over [-1,1]^3
    double inv_h;
    const int* cell_start; // [G^3+1]
    const int* items;      // triangle ids
};

struct WarpArgs {
    int M, V, F;
    const double* pts;     // [M][3] query points
    const double* from;    // [V][3] START mesh vertices
    const double* to;      // [V][3] END mesh vertices
    const int* tris;       // [F][3]
    WarpGrid grid;
    double radius;         // output radius (RAD)
    double orient;         // +1: faces wound outward, -1: inward (the mesh's OVERALL orientation, decided once on the host)
    double* out;           // [M][3]
    int* tri_out;          // optional [M]
    double* w_out;         // optional [M][3]
};

// One thread per triangle: the cells its (bulge-expanded) bounding box of unit directions touches.  FILL = false: count per
// cell; FILL = true: write the triangle id at cell_start[cell] + (running fill count).
template <bool FILL>
__global__ void __launch_bounds__(128) k_tri_cells(int F, int G, double inv_h, const double* __restrict__ xyz, const int* __restrict__ tris,
                                                   int* __restrict__ counts, const int* __restrict__ cell_start, int* __restrict__ fill,
                                                   int* __restrict__ items) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= F) return;
    double u[3][3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        const double* p = xyz + 3 * tris[3 * t + j];
        const double n = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
        u[j][0] = p[0] / n; u[j][1] = p[1] / n; u[j][2] = p[2] / n;
    }
    auto d2 = [](const double* x, const double* y) { return (x[0] - y[0]) * (x[0] - y[0]) + (x[1] - y[1]) * (x[1] - y[1]) + (x[2] - y[2]) * (x[2] - y[2]); };
    const double emax2 = fmax(d2(u[0], u[1]), fmax(d2(u[1], u[2]), d2(u[2], u[0])));
    const double sag = 1.0 - sqrt(fmax(0.0, 1.0 - emax2 / 3.0)) + 1e-9;   // how far the spherical triangle bulges out of its flat box
    int lo[3], hi[3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
        const double mn = fmin(u[0][d], fmin(u[1][d], u[2][d])) - sag, mx = fmax(u[0][d], fmax(u[1][d], u[2][d])) + sag;
        lo[d] = max(0, min(G - 1, (int)floor((mn + 1.0) * inv_h)));
        hi[d] = max(0, min(G - 1, (int)floor((mx + 1.0) * inv_h)));
    }
    for (int x = lo[0]; x <= hi[0]; x++)
        for (int y = lo[1]; y <= hi[1]; y++)
            for (int z = lo[2]; z <= hi[2]; z++) {
                const int c = (x * G + y) * G + z;
                if (!FILL) atomicAdd(counts + c, 1);
                else items[cell_start[c] + atomicAdd(fill + c, 1)] = t;
            }
}

__device__ __forceinline__ void w_unit(const double* p, double* u) {
    const double n = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
    u[0] = p[0] / n; u[1] = p[1] / n; u[2] = p[2] / n;
}
__device__ __forceinline__ double w_triple(const double* p, const double* a, const double* b) {  // p . (a x b)
    return p[0] * (a[1] * b[2] - a[2] * b[1]) + p[1] * (a[2] * b[0] - a[0] * b[2]) + p[2] * (a[0] * b[1] - a[1] * b[0]);
}
// margin of p inside the radial cone of triangle t (min of the three edge-plane tests); -inf for the far hemisphere
__device__ __forceinline__ double w_cone_margin(const WarpArgs& a, int t, const double* u) {
    double A[3], B[3], C[3];
    w_unit(a.from + 3 * a.tris[3 * t], A); w_unit(a.from + 3 * a.tris[3 * t + 1], B); w_unit(a.from + 3 * a.tris[3 * t + 2], C);
    if (u[0] * (A[0] + B[0] + C[0]) + u[1] * (A[1] + B[1] + C[1]) + u[2] * (A[2] + B[2] + C[2]) <= 0.0) return -1e300;
    // a.orient = -1 for a mesh whose faces are wound inward as a whole; a single flipped face of a folded mesh keeps a negative
    // margin either way (it does not contain points the way its neighbours do: oracle/geom.h SphereLocator)
    return fmin(a.orient * w_triple(u, A, B), fmin(a.orient * w_triple(u, B, C), a.orient * w_triple(u, C, A)));
}
__device__ __forceinline__ double w_area(const double* a, const double* b, const double* c) {
    const double e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    const double x = e1[1] * e2[2] - e1[2] * e2[1], y = e1[2] * e2[0] - e1[0] * e2[2], z = e1[0] * e2[1] - e1[1] * e2[0];
    return 0.5 * sqrt(x * x + y * y + z * z);
}

__global__ void __launch_bounds__(128) k_mesh_interpolate(WarpArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.M) return;
    const double p[3] = {a.pts[3 * i], a.pts[3 * i + 1], a.pts[3 * i + 2]};
    double u[3];
    w_unit(p, u);
    const WarpGrid& g = a.grid;
    int c[3];
#pragma unroll
    for (int d = 0; d < 3; d++) c[d] = min(g.G - 1, max(0, (int)floor((u[d] + 1.0) * g.inv_h)));
    const int cell = (c[0] * g.G + c[1]) * g.G + c[2];
    int best = -1;
    double bestv = -1e300;
    for (int k = g.cell_start[cell]; k < g.cell_start[cell + 1]; k++) {
        const int t = g.items[k];
        const double mn = w_cone_margin(a, t, u);
        if (mn > bestv) { bestv = mn; best = t; }
    }
    if (best < 0 || bestv < -1e-9) {  // folded / degenerate neighbourhood: fall back to the definition (all triangles)
        for (int t = 0; t < a.F; t++) {
            const double mn = w_cone_margin(a, t, u);
            if (mn > bestv) { bestv = mn; best = t; }
        }
    }
    const int n0 = a.tris[3 * best], n1 = a.tris[3 * best + 1], n2 = a.tris[3 * best + 2];
    const double* v0 = a.from + 3 * n0; const double* v1 = a.from + 3 * n1; const double* v2 = a.from + 3 * n2;
    // radial projection onto the triangle's plane ([EXT] A4), then sub-triangle areas
    double s1[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]}, s2[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]};
    double n1_ = sqrt(s1[0] * s1[0] + s1[1] * s1[1] + s1[2] * s1[2]), n2_ = sqrt(s2[0] * s2[0] + s2[1] * s2[1] + s2[2] * s2[2]);
    for (int d = 0; d < 3; d++) { s1[d] /= n1_; s2[d] /= n2_; }
    double s3[3] = {s1[1] * s2[2] - s1[2] * s2[1], s1[2] * s2[0] - s1[0] * s2[2], s1[0] * s2[1] - s1[1] * s2[0]};
    const double n3 = sqrt(s3[0] * s3[0] + s3[1] * s3[1] + s3[2] * s3[2]);
    for (int d = 0; d < 3; d++) s3[d] /= n3;
    const double si = (s3[0] * v0[0] + s3[1] * v0[1] + s3[2] * v0[2]) / (s3[0] * p[0] + s3[1] * p[1] + s3[2] * p[2]);
    const double PP[3] = {p[0] * si, p[1] * si, p[2] * si};
    const double Aa = w_area(PP, v1, v2), Ab = w_area(PP, v0, v2), Ac = w_area(PP, v0, v1);
    const double At = Aa + Ab + Ac;
    const double w0 = Aa / At, w1 = Ab / At, w2 = Ac / At;
    double q[3];
    for (int d = 0; d < 3; d++) q[d] = a.to[3 * n0 + d] * w0 + a.to[3 * n1 + d] * w1 + a.to[3 * n2 + d] * w2;
    const double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    for (int d = 0; d < 3; d++) a.out[3 * i + d] = q[d] / nq * a.radius;
    if (a.tri_out) a.tri_out[i] = best;
    if (a.w_out) { a.w_out[3 * i] = w0; a.w_out[3 * i + 1] = w1; a.w_out[3 * i + 2] = w2; }
}

}  // namespace msm
