// ico_hierarchy.h — host-side construction of the icosahedral face hierarchy that replaces the reference's
// Octree (newresampler::Octree, built at src/DiscreteModel.cpp:88, queried at src/DiscreteCostFunction.cpp:421,492).
//
// The target of the discrete cost function is always a regular icosphere (src/featurespace.cc:27-31,46), so the
// triangle whose radial cone contains a point can be found by: base face (20 candidates) + d four-way steps.
// Writing a point in the (non-orthogonal) basis of a face's three corner directions, p = alpha*a + beta*b + gamma*c,
// the four-way step becomes pure scalar arithmetic (see locate.cuh):
//     corner child at a  <=> alpha > beta + gamma            new coords (alpha-beta-gamma, beta*n_ab, gamma*n_ca)
//     centre child       otherwise                           new coords (n_bc(b+g-a), n_ca(a+g-b), n_ab(a+b-g))
// where n_xy = |x + y| is the only geometry needed per hierarchy node.  Because all 20 base faces of a regular
// icosphere are congruent, ONE table of (n_ab, n_bc, n_ca) per node of one base face serves the whole sphere
// (sum_{l<d} 4^l nodes: 5461 for ico7).  At the leaf, (alpha,beta,gamma)/(alpha+beta+gamma) ARE the barycentric
// weights of the radially projected point ([EXT] A4), so no vertex coordinates are read at all.
//
// The hierarchy is rebuilt from the geometry of the mesh handed to msm_set_target, so any vertex numbering, face
// numbering and orientation of the icosahedron is accepted; a mesh that is not a regular icosphere is rejected.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace msm {

struct IcoHierarchy {
    int depth = -1;
    int V = 0, F = 0;
    double radius = 0;
    // canonical leaf index = base*4^depth + path  ->  vertex ids (corner order a,b,c) and the caller's triangle id
    std::vector<std::array<int, 3>> leaf_vid;
    std::vector<int> leaf_user_face;
    std::vector<int> user_to_leaf;
    // per base face: dual basis rows A,B,C (p.A = alpha ...) and centre direction
    double dual[20][9];
    double centre[20][3];
    // per hierarchy node (heap order: node 0 = root, children of n are 4n+1..4n+4): (n_ab, n_bc, n_ca)
    std::vector<std::array<double, 3>> ntab;
    // corner vertex ids of every face of every level (level l: index base*4^l + path) and the unit vertex directions:
    // kept for the per-level hint tables (build_hint_level)
    std::vector<std::vector<std::array<int, 3>>> lvl_vid;
    std::vector<std::array<double, 3>> unit;
    std::string error;

    static void norm3(double* v) {
        double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        v[0] /= n; v[1] /= n; v[2] /= n;
    }

    bool build(const double* xyz, int nv, const int* tris, int nf) {
        V = nv; F = nf;
        depth = -1;
        long f = 20;
        for (int d = 0; d <= 10; d++, f *= 4)
            if (f == nf) { depth = d; break; }
        if (depth < 0 || nv != nf / 2 + 2) { error = "target mesh is not an icosphere (F != 20*4^d or V != F/2+2)"; return false; }
        radius = std::sqrt(xyz[0] * xyz[0] + xyz[1] * xyz[1] + xyz[2] * xyz[2]);
        std::vector<std::array<double, 3>> u(nv);
        for (int i = 0; i < nv; i++) {
            double v[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
            double n = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            if (std::fabs(n - radius) > 1e-6 * radius) { error = "target vertices are not on one sphere"; return false; }
            u[i] = {v[0] / n, v[1] / n, v[2] / n};
        }
        // valence
        std::vector<int> valence(nv, 0);
        for (int t = 0; t < nf; t++)
            for (int j = 0; j < 3; j++) {
                int id = tris[3 * t + j];
                if (id < 0 || id >= nv) { error = "target triangle index out of range"; return false; }
                valence[id]++;
            }
        std::vector<int> base;
        for (int i = 0; i < nv; i++)
            if (valence[i] == 5) base.push_back(i);
        if (base.size() != 12) { error = "target mesh does not have exactly 12 valence-5 vertices"; return false; }

        // geometric hash of unit vertices
        const double edge = 1.1071487177940904 / double(1 << depth);  // icosahedron edge angle / 2^d
        const double cellsz = edge * 0.5;
        const double tol = edge * 0.05;
        auto cell = [&](double x) { return (int64_t)std::floor((x + 1.5) / cellsz); };
        auto key = [&](int64_t a, int64_t b, int64_t c) { return (a * 2000003LL + b) * 2000003LL + c; };
        std::unordered_map<int64_t, std::vector<int>> grid;
        grid.reserve(nv * 2);
        for (int i = 0; i < nv; i++) grid[key(cell(u[i][0]), cell(u[i][1]), cell(u[i][2]))].push_back(i);
        auto find_vertex = [&](const double* p) -> int {
            int64_t c0 = cell(p[0]), c1 = cell(p[1]), c2 = cell(p[2]);
            int best = -1;
            double bd = tol * tol;
            for (int64_t a = c0 - 1; a <= c0 + 1; a++)
                for (int64_t b = c1 - 1; b <= c1 + 1; b++)
                    for (int64_t c = c2 - 1; c <= c2 + 1; c++) {
                        auto it = grid.find(key(a, b, c));
                        if (it == grid.end()) continue;
                        for (int i : it->second) {
                            double dx = u[i][0] - p[0], dy = u[i][1] - p[1], dz = u[i][2] - p[2];
                            double d2 = dx * dx + dy * dy + dz * dz;
                            if (d2 < bd) { bd = d2; best = i; }
                        }
                    }
            return best;
        };

        // base faces: mutually adjacent triples of base vertices (cos angle = 1/sqrt(5)), oriented outward
        auto dot = [&](int i, int j) { return u[i][0] * u[j][0] + u[i][1] * u[j][1] + u[i][2] * u[j][2]; };
        const double cadj = 1.0 / std::sqrt(5.0);
        std::vector<std::array<int, 3>> bf;
        for (int i = 0; i < 12; i++)
            for (int j = i + 1; j < 12; j++)
                for (int k = j + 1; k < 12; k++) {
                    int a = base[i], b = base[j], c = base[k];
                    if (std::fabs(dot(a, b) - cadj) < 1e-6 && std::fabs(dot(b, c) - cadj) < 1e-6 && std::fabs(dot(a, c) - cadj) < 1e-6) {
                        double det = u[a][0] * (u[b][1] * u[c][2] - u[b][2] * u[c][1]) - u[a][1] * (u[b][0] * u[c][2] - u[b][2] * u[c][0]) +
                                     u[a][2] * (u[b][0] * u[c][1] - u[b][1] * u[c][0]);
                        if (det < 0) std::swap(b, c);
                        bf.push_back({a, b, c});
                    }
                }
        if (bf.size() != 20) { error = "could not recover the 20 base faces of the icosahedron"; return false; }

        // user's faces by sorted vertex triple
        auto tkey = [&](int a, int b, int c) {
            int s[3] = {a, b, c};
            std::sort(s, s + 3);
            return ((int64_t)s[0] * nv + s[1]) * nv + s[2];
        };
        std::unordered_map<int64_t, int> facemap;
        facemap.reserve(nf * 2);
        for (int t = 0; t < nf; t++) facemap[tkey(tris[3 * t], tris[3 * t + 1], tris[3 * t + 2])] = t;

        const long nleaf_per_base = 1L << (2 * depth);
        leaf_vid.assign(nf, {0, 0, 0});
        leaf_user_face.assign(nf, -1);
        user_to_leaf.assign(nf, -1);
        long nnodes = 0;
        for (int l = 0; l < depth; l++) nnodes += 1L << (2 * l);
        ntab.assign(nnodes, {0, 0, 0});
        std::vector<char> ntab_set(nnodes, 0);

        struct Item { int a, b, c; long node; long path; int level; };
        lvl_vid.assign(depth + 1, {});
        for (int l = 0; l <= depth; l++) lvl_vid[l].assign((size_t)20 << (2 * l), {0, 0, 0});
        bool ok = true;
        for (int b = 0; b < 20 && ok; b++) {
            // dual basis and centre
            const double* A = u[bf[b][0]].data(); const double* B = u[bf[b][1]].data(); const double* Cc = u[bf[b][2]].data();
            double bxc[3] = {B[1] * Cc[2] - B[2] * Cc[1], B[2] * Cc[0] - B[0] * Cc[2], B[0] * Cc[1] - B[1] * Cc[0]};
            double cxa[3] = {Cc[1] * A[2] - Cc[2] * A[1], Cc[2] * A[0] - Cc[0] * A[2], Cc[0] * A[1] - Cc[1] * A[0]};
            double axb[3] = {A[1] * B[2] - A[2] * B[1], A[2] * B[0] - A[0] * B[2], A[0] * B[1] - A[1] * B[0]};
            double D = A[0] * bxc[0] + A[1] * bxc[1] + A[2] * bxc[2];
            for (int j = 0; j < 3; j++) { dual[b][j] = bxc[j] / D; dual[b][3 + j] = cxa[j] / D; dual[b][6 + j] = axb[j] / D; }
            double cen[3] = {A[0] + B[0] + Cc[0], A[1] + B[1] + Cc[1], A[2] + B[2] + Cc[2]};
            norm3(cen);
            for (int j = 0; j < 3; j++) centre[b][j] = cen[j];

            std::vector<Item> stack;
            stack.push_back({bf[b][0], bf[b][1], bf[b][2], 0, 0, 0});
            while (!stack.empty() && ok) {
                Item it = stack.back();
                stack.pop_back();
                lvl_vid[it.level][((size_t)b << (2 * it.level)) + it.path] = {it.a, it.b, it.c};
                if (it.level == depth) {
                    long leaf = (long)b * nleaf_per_base + it.path;
                    auto fm = facemap.find(tkey(it.a, it.b, it.c));
                    if (fm == facemap.end()) { error = "target mesh faces do not follow the 4-split icosphere hierarchy"; ok = false; break; }
                    leaf_vid[leaf] = {it.a, it.b, it.c};
                    leaf_user_face[leaf] = fm->second;
                    user_to_leaf[fm->second] = (int)leaf;
                    continue;
                }
                double mab[3], mbc[3], mca[3];
                for (int j = 0; j < 3; j++) {
                    mab[j] = u[it.a][j] + u[it.b][j];
                    mbc[j] = u[it.b][j] + u[it.c][j];
                    mca[j] = u[it.c][j] + u[it.a][j];
                }
                double nab = std::sqrt(mab[0] * mab[0] + mab[1] * mab[1] + mab[2] * mab[2]);
                double nbc = std::sqrt(mbc[0] * mbc[0] + mbc[1] * mbc[1] + mbc[2] * mbc[2]);
                double nca = std::sqrt(mca[0] * mca[0] + mca[1] * mca[1] + mca[2] * mca[2]);
                if (!ntab_set[it.node]) { ntab[it.node] = {nab, nbc, nca}; ntab_set[it.node] = 1; }
                else if (std::fabs(ntab[it.node][0] - nab) > 1e-6 || std::fabs(ntab[it.node][1] - nbc) > 1e-6 ||
                         std::fabs(ntab[it.node][2] - nca) > 1e-6) {
                    error = "target mesh is not a REGULAR icosphere (base faces are not congruent)"; ok = false; break;
                }
                for (int j = 0; j < 3; j++) { mab[j] /= nab; mbc[j] /= nbc; mca[j] /= nca; }
                int iab = find_vertex(mab), ibc = find_vertex(mbc), ica = find_vertex(mca);
                if (iab < 0 || ibc < 0 || ica < 0) { error = "target mesh vertices do not follow the icosphere midpoint rule"; ok = false; break; }
                long n4 = 4 * it.node + 1, p4 = 4 * it.path;
                stack.push_back({it.a, iab, ica, n4 + 0, p4 + 0, it.level + 1});
                stack.push_back({iab, it.b, ibc, n4 + 1, p4 + 1, it.level + 1});
                stack.push_back({ica, ibc, it.c, n4 + 2, p4 + 2, it.level + 1});
                stack.push_back({ibc, ica, iab, n4 + 3, p4 + 3, it.level + 1});
            }
        }
        if (!ok) return false;
        unit = u;
        for (int t = 0; t < nf; t++)
            if (user_to_leaf[t] < 0) { error = "target mesh has faces outside the icosphere hierarchy"; return false; }
        return true;
    }

    // Hint tables for level h: per face the dual basis of its (unit) corner directions, and per edge the neighbouring
    // face with the 3x3 change of corner coordinates (x_nbr = T * x_face).  Edge e is the one opposite corner e.
    void build_hint_level(int h, std::vector<double>& hdual, std::vector<int>& hnbr, std::vector<double>& hT) const {
        const auto& fv = lvl_vid[h];
        const size_t nfh = fv.size();
        hdual.assign(nfh * 9, 0.0);
        hnbr.assign(nfh * 3, -1);
        hT.assign(nfh * 27, 0.0);
        auto dual_of = [&](const std::array<int, 3>& f, double* d) {
            const double* A = unit[f[0]].data(); const double* B = unit[f[1]].data(); const double* Cc = unit[f[2]].data();
            double bxc[3] = {B[1] * Cc[2] - B[2] * Cc[1], B[2] * Cc[0] - B[0] * Cc[2], B[0] * Cc[1] - B[1] * Cc[0]};
            double cxa[3] = {Cc[1] * A[2] - Cc[2] * A[1], Cc[2] * A[0] - Cc[0] * A[2], Cc[0] * A[1] - Cc[1] * A[0]};
            double axb[3] = {A[1] * B[2] - A[2] * B[1], A[2] * B[0] - A[0] * B[2], A[0] * B[1] - A[1] * B[0]};
            double D = A[0] * bxc[0] + A[1] * bxc[1] + A[2] * bxc[2];
            for (int j = 0; j < 3; j++) { d[j] = bxc[j] / D; d[3 + j] = cxa[j] / D; d[6 + j] = axb[j] / D; }
        };
        for (size_t f = 0; f < nfh; f++) dual_of(fv[f], &hdual[9 * f]);
        std::unordered_map<int64_t, std::pair<int, int>> edge_face;  // sorted edge -> first (face, edge slot)
        edge_face.reserve(nfh * 3);
        auto ekey = [&](int a, int b) { return (int64_t)std::min(a, b) * V + std::max(a, b); };
        for (size_t f = 0; f < nfh; f++)
            for (int e = 0; e < 3; e++) {
                int64_t k = ekey(fv[f][(e + 1) % 3], fv[f][(e + 2) % 3]);
                auto it = edge_face.find(k);
                if (it == edge_face.end()) edge_face[k] = {(int)f, e};
                else {
                    int g = it->second.first, ge = it->second.second;
                    hnbr[3 * f + e] = g;
                    hnbr[3 * g + ge] = (int)f;
                }
            }
        for (size_t f = 0; f < nfh; f++)
            for (int e = 0; e < 3; e++) {
                int g = hnbr[3 * f + e];
                if (g < 0) continue;
                // T = Dual_g (rows) * M_f (columns = unit corners of f)
                for (int r = 0; r < 3; r++)
                    for (int c = 0; c < 3; c++) {
                        const double* col = unit[fv[f][c]].data();
                        hT[27 * f + 9 * e + 3 * r + c] = hdual[9 * g + 3 * r] * col[0] + hdual[9 * g + 3 * r + 1] * col[1] + hdual[9 * g + 3 * r + 2] * col[2];
                    }
            }
    }
};

}  // namespace msm
