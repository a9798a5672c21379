// locate.cuh — device-side spherical point location on a regular icosphere (north_star kernel (c)).
// Replaces newresampler::Octree::get_closest_triangle + barycentric_weight ([EXT] A3/A4; call sites
// src/DiscreteCostFunction.cpp:421-433,492-505).  See ico_hierarchy.h for the derivation.
#pragma once
#include <cuda_runtime.h>

namespace msm {

constexpr int kMtabPad = 3;   // leading pad entries of the multiplier table (see msm_set_target)

// 20 base faces: dual basis rows (alpha = A.p, beta = B.p, gamma = C.p) and centre directions.
struct IcoBase {
    float dual[20][9];
    float centre[20][3];
    double dual_d[20][9];
    double centre_d[20][3];
};

struct IcoDev {
    const IcoBase* base;         // device memory, one per context: contexts over differently oriented spheres coexist
    int depth;
    int nleaf_per_base;          // 4^depth
    int leaf_off;                // (4^depth - 1)/3 : heap index of the first leaf-level node
    const float4* ntab;          // [nodes] (n_ab, n_bc, n_ca, -)
    const double4* ntab_d;       // same in fp64 (msm_locate, precise paths)
    const float4* mtab;          // [nodes incl. leaf level] multipliers applied on ENTERING a node from its parent:
                                 // corner-a child (1,nab,nca) | corner-b (nab,1,nbc) | corner-c (nca,nbc,1) | centre (nbc,nca,nab)
};

// Per-level "hint" tables (ico_hierarchy.h build_hint_level): the faces of one intermediate level h act as local
// coordinate frames.  A control point's whole patch lives within a few faces of the level-h face under the control
// point, so samples are expressed in that face's corner basis (precisely: offsets from the control point are small
// numbers), walked across edges where needed, and only the remaining depth-h levels are descended.
struct HintDev {
    int h;                  // hint level
    int nf_per_base;        // 4^h
    int node_off;           // (4^h - 1)/3
    const double* hdual;    // [20*4^h][9]
    const int* hnbr;        // [20*4^h][3]
    const float4* hT;       // [20*4^h][3 edges][3 rows] : (T row, w of row 0 = neighbour face id as int bits)
};

// Destination tables of a fused exchange: device pointers, valid on THIS GPU, to the same buffer on every rank of a
// control-point / triplet sharded run (CUDA IPC or symmetric-memory mappings over NVLink).  world == 0: not used.
constexpr int kMaxPeers = 16;
struct PeerSet {
    float* p[kMaxPeers];
    float* mc;      // optional NVSwitch multicast address of the same buffer: ONE store reaches every rank
    int world;
};
// multimem.st: the store is replicated to all GPUs of the multicast object by the switch (NVLS)
__device__ __forceinline__ void multicast_store(float* addr, float v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void multicast_store(float4* addr, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <typename T> struct NodeN;
template <> struct NodeN<float> {
    static __device__ __forceinline__ void load(const IcoDev& ico, int n, float& nab, float& nbc, float& nca) {
        float4 v = __ldg(ico.ntab + n);
        nab = v.x; nbc = v.y; nca = v.z;
    }
};
template <> struct NodeN<double> {
    static __device__ __forceinline__ void load(const IcoDev& ico, int n, double& nab, double& nbc, double& nca) {
        double4 v = ico.ntab_d[n];
        nab = v.x; nbc = v.y; nca = v.z;
    }
};

template <typename T> __device__ __forceinline__ const T* base_dual(const IcoDev& ico, int b);
template <> __device__ __forceinline__ const float* base_dual<float>(const IcoDev& ico, int b) { return ico.base->dual[b]; }
template <> __device__ __forceinline__ const double* base_dual<double>(const IcoDev& ico, int b) { return ico.base->dual_d[b]; }
template <typename T> __device__ __forceinline__ const T* base_centre(const IcoDev& ico, int b);
template <> __device__ __forceinline__ const float* base_centre<float>(const IcoDev& ico, int b) { return ico.base->centre[b]; }
template <> __device__ __forceinline__ const double* base_centre<double>(const IcoDev& ico, int b) { return ico.base->centre_d[b]; }

// Base face containing p: the face cones of a regular icosahedron are the Voronoi cells of the face centres.
template <typename T>
__device__ __forceinline__ int ico_base_face(const IcoDev& ico, T px, T py, T pz) {
    int best = 0;
    T bv = -1e30f;
#pragma unroll
    for (int b = 0; b < 20; b++) {
        const T* c = base_centre<T>(ico, b);
        T v = px * c[0] + py * c[1] + pz * c[2];
        if (v > bv) { bv = v; best = b; }
    }
    return best;
}

// Four-way descent.  In: coordinates (al,be,ga) of p in the corner basis of base face `base`.
// Out: canonical leaf index, and (al,be,ga) in the leaf's corner basis (un-normalised barycentric weights).
template <typename T>
__device__ __forceinline__ int ico_descend(const IcoDev& ico, int base, T& al, T& be, T& ga) {
    int n = 0;
    for (int l = 0; l < ico.depth; l++) {
        T nab, nbc, nca;
        NodeN<T>::load(ico, n, nab, nbc, nca);
        const T s = al + be + ga;
        const T sa = (al + al) - s, sb = (be + be) - s, sc = (ga + ga) - s;  // alpha-beta-gamma etc.
        int child;
        T x, y, z;
        if (sa > T(0)) { child = 0; x = sa;       y = be * nab; z = ga * nca; }
        else if (sb > T(0)) { child = 1; x = al * nab; y = sb;       z = ga * nbc; }
        else if (sc > T(0)) { child = 2; x = al * nca; y = be * nbc; z = sc; }
        else { child = 3; x = -sa * nbc; y = -sb * nca; z = -sc * nab; }
        al = x; be = y; ga = z;
        n = 4 * n + 1 + child;
    }
    return base * ico.nleaf_per_base + (n - ico.leaf_off);
}

// Descent of `nlev` levels starting at heap node n (fp32): returns the heap node reached.
// Branch-free: the child is picked from the signs of (2x - s), the child's incoming multipliers come from one
// 16-byte table read, and the value entering each product is |2x - s| for the child's own corner / the centre child
// and x otherwise.
__device__ __forceinline__ void ico_descend_level(const float4* __restrict__ mtab, unsigned& n, float& al, float& be, float& ga) {
    const float s = al + be + ga;
    const float sa = fmaf(al, 2.f, -s), sb = fmaf(be, 2.f, -s), sc = fmaf(ga, 2.f, -s);
    const bool ca = sa > 0.f, cb = sb > 0.f, cc = sc > 0.f;
    // child: centre unless one of (2x - s) is positive; on (rounding-only) double positives the earlier corner wins
    unsigned idx = 4u * n + 4u;
    idx = cc ? idx - 1u : idx;
    idx = cb ? 4u * n + 2u : idx;
    idx = ca ? 4u * n + 1u : idx;
    n = idx;
    const float4 m = __ldg(mtab + idx);
    const bool cb_ = cb && !ca, cc_ = cc && !ca && !cb;
    const float va = (cb_ || cc_) ? al : fabsf(sa);
    const float vb = (ca || cc_) ? be : fabsf(sb);
    const float vc = (ca || cb_) ? ga : fabsf(sc);
    // __fmul_rn: never contracted into the next level's additions — whether ptxas fuses depends on the surrounding code,
    // and the same item must give the same bits at every call site (label-group size / shard independence)
    al = __fmul_rn(va, m.x); be = __fmul_rn(vb, m.y); ga = __fmul_rn(vc, m.z);
}

template <int NLEV>
__device__ __forceinline__ unsigned ico_descend_fixed(const float4* __restrict__ mtab, unsigned n, float& al, float& be, float& ga) {
#pragma unroll
    for (int l = 0; l < NLEV; l++) ico_descend_level(mtab, n, al, be, ga);
    return n;
}

__device__ __forceinline__ int ico_descend_from(const float4* __restrict__ mtab, int nlev, int n0, float& al, float& be, float& ga) {
    unsigned n = (unsigned)n0;
    switch (nlev) {  // uniform across the launch: fully unrolled bodies
        case 1: n = ico_descend_fixed<1>(mtab, n, al, be, ga); break;
        case 2: n = ico_descend_fixed<2>(mtab, n, al, be, ga); break;
        case 3: n = ico_descend_fixed<3>(mtab, n, al, be, ga); break;
        case 4: n = ico_descend_fixed<4>(mtab, n, al, be, ga); break;
        case 5: n = ico_descend_fixed<5>(mtab, n, al, be, ga); break;
        case 6: n = ico_descend_fixed<6>(mtab, n, al, be, ga); break;
        case 7: n = ico_descend_fixed<7>(mtab, n, al, be, ga); break;
        default:
#pragma unroll 1
            for (int l = 0; l < nlev; l++) ico_descend_level(mtab, n, al, be, ga);
    }
    return (int)n;
}

// Visibility walk on the level-h faces: cross the edge opposite the most negative corner coordinate.
__device__ __forceinline__ int ico_walk(const HintDev& hd, int face, float& al, float& be, float& ga) {
    for (int it = 0; it < 64; it++) {
        const float mn = fminf(al, fminf(be, ga));
        if (mn >= 0.f) break;
        const int e = (al == mn) ? 0 : ((be == mn) ? 1 : 2);
        // three 16-byte rows per (face, edge): the 3x3 change of corner coordinates; row 0 carries the neighbour id
        const float4* T = hd.hT + (unsigned)(face * 3 + e) * 3u;
        const float4 r0 = __ldg(T), r1 = __ldg(T + 1), r2 = __ldg(T + 2);
        const float x = fmaf(r0.z, ga, fmaf(r0.y, be, r0.x * al));   // contraction spelled out: same bits at every call site
        const float y = fmaf(r1.z, ga, fmaf(r1.y, be, r1.x * al));
        const float z = fmaf(r2.z, ga, fmaf(r2.y, be, r2.x * al));
        face = __float_as_int(r0.w);
        al = x; be = y; ga = z;
    }
    return face;
}

// Full location: hint base face first (all samples of one control point's patch almost always share it).
template <typename T>
__device__ __forceinline__ int ico_locate(const IcoDev& ico, int hint_base, T px, T py, T pz, T& al, T& be, T& ga) {
    int base = hint_base;
    const T* d = base_dual<T>(ico, base);
    al = d[0] * px + d[1] * py + d[2] * pz;
    be = d[3] * px + d[4] * py + d[5] * pz;
    ga = d[6] * px + d[7] * py + d[8] * pz;
    if (fminf(float(al), fminf(float(be), float(ga))) < 0.f) {
        base = ico_base_face<T>(ico, px, py, pz);
        d = base_dual<T>(ico, base);
        al = d[0] * px + d[1] * py + d[2] * pz;
        be = d[3] * px + d[4] * py + d[5] * pz;
        ga = d[6] * px + d[7] * py + d[8] * pz;
        // rounding at a cone boundary can leave a tiny negative coordinate: clamp (weights stay continuous)
        al = al < T(0) ? T(0) : al; be = be < T(0) ? T(0) : be; ga = ga < T(0) ? T(0) : ga;
    }
    return ico_descend<T>(ico, base, al, be, ga);
}

// Minimal (Rodrigues) rotation taking direction a onto direction b, returned as K = R - I, fp64.
// Restates estimate_rotation_matrix ([EXT] A2): identity below 1e-8 rad.
__device__ __forceinline__ void rodrigues_minus_identity(const double a_[3], const double b_[3], double K[9]) {
    double na = sqrt(a_[0] * a_[0] + a_[1] * a_[1] + a_[2] * a_[2]);
    double nb = sqrt(b_[0] * b_[0] + b_[1] * b_[1] + b_[2] * b_[2]);
    double a[3] = {a_[0] / na, a_[1] / na, a_[2] / na}, b[3] = {b_[0] / nb, b_[1] / nb, b_[2] / nb};
    double c = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    c = fmin(1.0, fmax(-1.0, c));
    // no FMA contraction here: for exactly (anti)parallel directions the cross product must be exactly zero, as in the
    // reference's plain C++ (the antipode of the label-grid centre is such a control point on a regular icosphere)
    double k[3] = {__dsub_rn(__dmul_rn(a[1], b[2]), __dmul_rn(a[2], b[1])), __dsub_rn(__dmul_rn(a[2], b[0]), __dmul_rn(a[0], b[2])),
                   __dsub_rn(__dmul_rn(a[0], b[1]), __dmul_rn(a[1], b[0]))};
    double s = sqrt(k[0] * k[0] + k[1] * k[1] + k[2] * k[2]);  // sin(theta), well conditioned for small theta
    double theta = atan2(s, c);
#pragma unroll
    for (int i = 0; i < 9; i++) K[i] = 0.0;
    if (s < 1e-12 && c < 0.0) {
        // antiparallel: rotation by pi about a x e, e = coordinate axis of a's smallest component (the convention shared
        // with the host mirror and the oracle; the reference formula degenerates here).  R - I = 2 (u u^T - I).
        const double ax = fabs(a[0]), ay = fabs(a[1]), az = fabs(a[2]);
        double u[3];
        if (ax <= ay && ax <= az) { u[0] = 0.0; u[1] = a[2]; u[2] = -a[1]; }        // a x (1,0,0)
        else if (ay <= az) { u[0] = -a[2]; u[1] = 0.0; u[2] = a[0]; }               // a x (0,1,0)
        else { u[0] = a[1]; u[1] = -a[0]; u[2] = 0.0; }                            // a x (0,0,1)
        const double nu = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
        u[0] /= nu; u[1] /= nu; u[2] /= nu;
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) K[3 * i + j] = 2.0 * u[i] * u[j] - (i == j ? 2.0 : 0.0);
        return;
    }
    if (fabs(theta) < 1e-8 || s == 0.0) return;
    double kx = k[0] / s, ky = k[1] / s, kz = k[2] / s;
    // 1 - cos(theta) = s^2 / (1 + c) avoids cancellation for small angles
    double omc = (c > 0.0) ? s * s / (1.0 + c) : 1.0 - c;
    // K = s*[k]x + omc*[k]x^2 ; [k]x^2 = k k^T - I
    K[0] = omc * (kx * kx - 1.0);      K[1] = -s * kz + omc * kx * ky;   K[2] = s * ky + omc * kx * kz;
    K[3] = s * kz + omc * kx * ky;     K[4] = omc * (ky * ky - 1.0);     K[5] = -s * kx + omc * ky * kz;
    K[6] = -s * ky + omc * kx * kz;    K[7] = s * kx + omc * ky * kz;    K[8] = omc * (kz * kz - 1.0);
}

}  // namespace msm
