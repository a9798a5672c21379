// cpu_hierarchy.cpp — TEST INFRASTRUCTURE (CPU): drives the product's host-side hierarchy builder
// (msm_newmeshreg_b200/csrc/ico_hierarchy.h, plain C++) without a GPU, so that `-m "not gpu"` covers it.
// The scalar four-way step and the edge walk are restated here in fp64 exactly as csrc/locate.cuh states them for the
// device (ico_descend / ico_walk); tests/test_host_cpu.py compares the result with the oracle's brute-force closest
// triangle ([EXT] A3/A4 semantics, oracle/geom.h).
#include <cstring>
#include <string>
#include <vector>

#include "../msm_newmeshreg_b200/csrc/ico_hierarchy.h"

namespace {
msm::IcoHierarchy H;
std::string g_err;
std::vector<unsigned char> leaf_perm;  // leaf corner j sits at position leaf_perm[3*leaf+j] of the caller's triangle

int base_face(const double* p) {
    int best = 0;
    double bv = -1e300;
    for (int b = 0; b < 20; b++) {
        const double v = p[0] * H.centre[b][0] + p[1] * H.centre[b][1] + p[2] * H.centre[b][2];
        if (v > bv) { bv = v; best = b; }
    }
    return best;
}

// `nlev` four-way steps from heap node n (csrc/locate.cuh ico_descend)
void descend(int nlev, long& n, double& al, double& be, double& ga) {
    for (int l = 0; l < nlev; l++) {
        const double nab = H.ntab[n][0], nbc = H.ntab[n][1], nca = H.ntab[n][2];
        const double s = al + be + ga;
        const double sa = (al + al) - s, sb = (be + be) - s, sc = (ga + ga) - s;
        int child;
        double x, y, z;
        if (sa > 0) { child = 0; x = sa; y = be * nab; z = ga * nca; }
        else if (sb > 0) { child = 1; x = al * nab; y = sb; z = ga * nbc; }
        else if (sc > 0) { child = 2; x = al * nca; y = be * nbc; z = sc; }
        else { child = 3; x = -sa * nbc; y = -sb * nca; z = -sc * nab; }
        al = x; be = y; ga = z;
        n = 4 * n + 1 + child;
    }
}
}  // namespace

extern "C" {

const char* ich_error() { return g_err.c_str(); }

int ich_build(const double* xyz, int nv, const int* tris, int nf) {
    H = msm::IcoHierarchy();
    if (!H.build(xyz, nv, tris, nf)) { g_err = H.error; return 1; }
    leaf_perm.assign((size_t)nf * 3, 0);
    for (int f = 0; f < nf; f++) {
        const int uf = H.leaf_user_face[f];
        for (int j = 0; j < 3; j++)
            for (int q = 0; q < 3; q++)
                if (tris[3 * uf + q] == H.leaf_vid[f][j]) leaf_perm[3 * f + j] = (unsigned char)q;
    }
    return 0;
}

int ich_depth() { return H.depth; }

// full location: caller's triangle id and barycentric weights in the caller's corner order
void ich_locate(const double* pts, int M, int* tri, double* w) {
    const long nleaf = 1L << (2 * H.depth), leaf_off = (nleaf - 1) / 3;
    for (int i = 0; i < M; i++) {
        const double* p = pts + 3 * i;
        const int b = base_face(p);
        const double* d = H.dual[b];
        double al = d[0] * p[0] + d[1] * p[1] + d[2] * p[2], be = d[3] * p[0] + d[4] * p[1] + d[5] * p[2],
               ga = d[6] * p[0] + d[7] * p[1] + d[8] * p[2];
        long n = 0;
        descend(H.depth, n, al, be, ga);
        const long leaf = b * nleaf + (n - leaf_off);
        const double inv = 1.0 / (al + be + ga);
        const double wl[3] = {al * inv, be * inv, ga * inv};
        tri[i] = H.leaf_user_face[leaf];
        for (int j = 0; j < 3; j++) w[3 * i + leaf_perm[3 * leaf + j]] = wl[j];
    }
}

// Level-h route of k_unary: express p in the frame of the level-h face under `start` (another point — the control
// point of a patch), walk across edges with the hT tables until no coordinate is negative, then descend the remaining
// levels.  Returns the caller's triangle ids; steps[i] = number of edge crossings.
int ich_locate_via_hint(int h, const double* starts, const double* pts, int M, int* tri, double* w, int* steps) {
    if (h < 0 || h > H.depth) { g_err = "hint level out of range"; return 1; }
    std::vector<double> hdual, hT;
    std::vector<int> hnbr;
    H.build_hint_level(h, hdual, hnbr, hT);
    const long nleaf = 1L << (2 * H.depth), leaf_off = (nleaf - 1) / 3;
    const long nfh = 1L << (2 * h), node_off = (nfh - 1) / 3;
    for (int i = 0; i < M; i++) {
        const double* s = starts + 3 * i;
        const int b0 = base_face(s);
        const double* d0 = H.dual[b0];
        double al = d0[0] * s[0] + d0[1] * s[1] + d0[2] * s[2], be = d0[3] * s[0] + d0[4] * s[1] + d0[5] * s[2],
               ga = d0[6] * s[0] + d0[7] * s[1] + d0[8] * s[2];
        long n = 0;
        descend(h, n, al, be, ga);
        long face = b0 * nfh + (n - node_off);
        const double* p = pts + 3 * i;
        const double* A = &hdual[9 * face];
        al = A[0] * p[0] + A[1] * p[1] + A[2] * p[2]; be = A[3] * p[0] + A[4] * p[1] + A[5] * p[2]; ga = A[6] * p[0] + A[7] * p[1] + A[8] * p[2];
        int it = 0;
        for (; it < 64; it++) {
            const double mn = std::min(al, std::min(be, ga));
            if (mn >= 0) break;
            const int e = (al == mn) ? 0 : ((be == mn) ? 1 : 2);
            const double* T = &hT[27 * face + 9 * e];
            const double x = T[0] * al + T[1] * be + T[2] * ga, y = T[3] * al + T[4] * be + T[5] * ga, z = T[6] * al + T[7] * be + T[8] * ga;
            face = hnbr[3 * face + e];
            al = x; be = y; ga = z;
        }
        steps[i] = it;
        const long b = face / nfh;
        n = node_off + (face - b * nfh);
        descend(H.depth - h, n, al, be, ga);
        const long leaf = b * nleaf + (n - leaf_off);
        const double inv = 1.0 / (al + be + ga);
        const double wl[3] = {al * inv, be * inv, ga * inv};
        tri[i] = H.leaf_user_face[leaf];
        for (int j = 0; j < 3; j++) w[3 * i + leaf_perm[3 * leaf + j]] = wl[j];
    }
    return 0;
}

}  // extern "C"
