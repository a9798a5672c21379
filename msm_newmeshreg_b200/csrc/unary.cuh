// unary.cuh — the unary data term (src/DiscreteCostFunction.cpp:306-313,412-448,483-526; similarities.cc:122-173).
//
// k_cp_hint (fp64, one thread per control point k): the level-h "hint" face under CP_k (and the target value under CP_k);
// k_unary_setup (fp64, one warp per control point k): CP_k's corner coordinates in that face, and from those
//     per sample i :  c_i = x(CP_k) + A r_i           (r_i = SRC_i - CP_k, the packed offsets)
//     per label  l :  M_l = A (R_kl - I),  g_l = A (R_kl - I) CP_k
//   where A is the dual basis of the hint face (x = A p are the coordinates of p in the face's corner basis); plus
//   tref_k = target feature under CP_k (Pearson correlation is shift invariant: the univariate kernel works on
//   target values relative to tref_k and source values relative to the patch's first sample, so that fp32 storage
//   resolves the small variation inside a patch instead of the field's absolute level).
// k_unary (fp32, one CTA per (control point, label group)):
//   Phase 1: stage the patch (r, c, source features, weights) and the label tables in shared memory — all coalesced.
//   Phase 2: every thread takes (label, sample) items.  Corner coordinates of the rotated sample R_kl (CP_k + r_i):
//            x = c_i + g_l + M_l r_i   (12 FMA), edge walk on level h where a coordinate is negative (rare: the hint
//            level is chosen so that a face is much larger than a patch), branch-free 4-way scalar descent of the
//            remaining levels, ONE 16-byte gather of the face-packed target features, interpolation.
//   Phase 3: 8 lanes per label reduce the weighted Pearson moments (fp64 accumulators, fixed order => deterministic).
// CQ = number of float4 channel quads of the multivariate cost function (0 = univariate).
#pragma once
#include "locate.cuh"

namespace msm {

// fp64 descent of a point from its base face: `nlev` levels; returns base face, heap node, corner coordinates
__device__ __forceinline__ void descend_d(const IcoDev& ico, int nlev, double px, double py, double pz, int& base, int& n, double& al,
                                          double& be, double& ga) {
    base = ico_base_face<double>(ico, px, py, pz);
    const double* d = base_dual<double>(ico, base);
    al = d[0] * px + d[1] * py + d[2] * pz; be = d[3] * px + d[4] * py + d[5] * pz; ga = d[6] * px + d[7] * py + d[8] * pz;
    n = 0;
    for (int l = 0; l < nlev; l++) {
        const double4 nn = ico.ntab_d[n];
        const double s = al + be + ga;
        const double sa = (al + al) - s, sb = (be + be) - s, sc = (ga + ga) - s;
        int child; double x, y, z;
        if (sa > 0.0) { child = 0; x = sa; y = be * nn.x; z = ga * nn.z; }
        else if (sb > 0.0) { child = 1; x = al * nn.x; y = sb; z = ga * nn.y; }
        else if (sc > 0.0) { child = 2; x = al * nn.z; y = be * nn.y; z = sc; }
        else { child = 3; x = -sa * nn.y; y = -sb * nn.z; z = -sc * nn.x; }
        al = x; be = y; ga = z;
        n = 4 * n + 1 + child;
    }
}

// one thread per source vertex: its leaf of the target hierarchy (canonical nested numbering) — the spatial sort key of
// k_patch_pack
__global__ void __launch_bounds__(128) k_vertex_key(int V, const double* __restrict__ src, IcoDev ico, int* __restrict__ key) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    int base, n; double al, be, ga;
    descend_d(ico, ico.depth, src[3 * v], src[3 * v + 1], src[3 * v + 2], base, n, al, be, ga);
    key[v] = base * ico.nleaf_per_base + (n - ico.leaf_off);
}

struct UnarySetupArgs {
    int k_begin, L;
    const int* off; const int* end;     // samples of row r: [off[r], end[r]) of the packed streams (CSR: end = off + 1; fixed-capacity
                                        // rows of the one-pass build, patchfast.cuh: off[r] = r cap, end[r] = off[r] + count)
    const float* px; const float* py; const float* pz;  // packed offsets r
    const double* Kd;                   // [N][L][9]  R_kl - I
    const double* cp;                   // [N][3]
    const float* face_feat;             // univariate face-packed features (tref) or nullptr
    IcoDev ico;
    HintDev hint;
    int* hint_face;                     // [Ns]
    float* tref;                        // [Ns]
    float* lab12;                       // [Ns][L][12] : M (9) then g (3)
    float* pcx; float* pcy; float* pcz; // [total] corner coordinates of the unrotated samples
};

// One THREAD per control point: the fp64 walk down the hierarchy — hint face, then on to the leaf for the reference target
// value.  A chain of dependent table reads per control point, so control points go side by side (it used to be lane 0 of the
// warp below, 31 lanes waiting).
__global__ void __launch_bounds__(128) k_cp_hint(UnarySetupArgs a, int Ns) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= Ns) return;
    const int k = a.k_begin + row;
    int base, n; double al, be, ga;
    descend_d(a.ico, a.hint.h, a.cp[3 * k], a.cp[3 * k + 1], a.cp[3 * k + 2], base, n, al, be, ga);
    a.hint_face[row] = base * a.hint.nf_per_base + (n - a.hint.node_off);
    float tr = 0.f;
    if (a.face_feat) {  // continue the same descent down to the leaf
        for (int l = a.hint.h; l < a.ico.depth; l++) {
            const double4 nn = a.ico.ntab_d[n];
            const double s = al + be + ga;
            const double sa = (al + al) - s, sb = (be + be) - s, sc = (ga + ga) - s;
            int child; double x, y, z;
            if (sa > 0.0) { child = 0; x = sa; y = be * nn.x; z = ga * nn.z; }
            else if (sb > 0.0) { child = 1; x = al * nn.x; y = sb; z = ga * nn.y; }
            else if (sc > 0.0) { child = 2; x = al * nn.z; y = be * nn.y; z = sc; }
            else { child = 3; x = -sa * nn.y; y = -sb * nn.z; z = -sc * nn.x; }
            al = x; be = y; ga = z;
            n = 4 * n + 1 + child;
        }
        tr = reinterpret_cast<const float4*>(a.face_feat)[base * a.ico.nleaf_per_base + (n - a.ico.leaf_off)].x;
    }
    a.tref[row] = tr;
}

// One WARP per control point (4 per CTA, no block barrier): from the hint face's dual basis A (k_cp_hint found the face),
//     per sample i :  c_i = A CP_k + A r_i     and     per label l :  M_l = A K_kl,  g_l = A K_kl CP_k.
__global__ void __launch_bounds__(128, 8) k_unary_setup(UnarySetupArgs a, int Ns, int staged) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= Ns) return;   // whole warps leave together
    const int k = a.k_begin + row;
    const double ox = a.cp[3 * k], oy = a.cp[3 * k + 1], oz = a.cp[3 * k + 2];
    const int f = a.hint_face[row];
    double A[9];   // same address in every lane: one broadcast transaction each
#pragma unroll
    for (int j = 0; j < 9; j++) A[j] = a.hint.hdual[(size_t)f * 9 + j];
    const double x0 = A[0] * ox + A[1] * oy + A[2] * oz, x1 = A[3] * ox + A[4] * oy + A[5] * oz, x2 = A[6] * ox + A[7] * oy + A[8] * oz;
    const int sb = a.off[row], se = a.end[row];
    for (int s = sb + lane; s < se; s += 32) {
        const double x = a.px[s], y = a.py[s], z = a.pz[s];
        a.pcx[s] = (float)(x0 + (A[0] * x + A[1] * y + A[2] * z));
        a.pcy[s] = (float)(x1 + (A[3] * x + A[4] * y + A[5] * z));
        a.pcz[s] = (float)(x2 + (A[6] * x + A[7] * y + A[8] * z));
    }
    // Per label: the control point's block of K (L x 9 doubles) and its output block (L x 12 floats) are contiguous, but a lane
    // working on label l touches them with strides of 72 / 48 bytes (12 sectors per request, ncu r6).  With `staged` the warp
    // moves both blocks through its slice of shared memory with coalesced requests; the arithmetic is the same either way.
    extern __shared__ double s_setup[];
    double* sK = s_setup + (size_t)(threadIdx.x >> 5) * a.L * 15;      // [L][9] doubles, then [L][12] floats (= 6 doubles per label)
    float* sO = reinterpret_cast<float*>(sK + (size_t)a.L * 9);
    const double* Kg = a.Kd + (size_t)k * a.L * 9;
    float* og = a.lab12 + (size_t)row * a.L * 12;
    if (staged) {
        for (int i = lane; i < a.L * 9; i += 32) sK[i] = Kg[i];
        __syncwarp();
    }
    for (int l = lane; l < a.L; l += 32) {
        const double* K = (staged ? sK : Kg) + l * 9;
        // the control point itself moves by (R-I) CP_k (no cancellation: K holds R-I at full precision)
        const double dx = K[0] * ox + K[1] * oy + K[2] * oz, dy = K[3] * ox + K[4] * oy + K[5] * oz, dz = K[6] * ox + K[7] * oy + K[8] * oz;
        float* o = (staged ? sO : og) + l * 12;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const double A0 = A[3 * r], A1 = A[3 * r + 1], A2 = A[3 * r + 2];
            o[3 * r + 0] = (float)(A0 * K[0] + A1 * K[3] + A2 * K[6]);
            o[3 * r + 1] = (float)(A0 * K[1] + A1 * K[4] + A2 * K[7]);
            o[3 * r + 2] = (float)(A0 * K[2] + A1 * K[5] + A2 * K[8]);
            o[9 + r] = (float)(A0 * dx + A1 * dy + A2 * dz);
        }
    }
    if (staged) {
        __syncwarp();
        for (int i = lane; i < a.L * 12; i += 32) og[i] = sO[i];
    }
}

struct UnaryArgs {
    int k_begin, Ns, L, G, Lg;          // shard rows, labels, label groups, labels per group
    const int* off; const int* end;     // samples of row r: [off[r], end[r]) (see UnarySetupArgs)
    const float* px; const float* py; const float* pz;     // packed offsets r = SRC - CP_k
    const float* pcx; const float* pcy; const float* pcz;  // corner coordinates c
    const float* pa; const float* pw;   // packed source features / weights [Cp][total]
    int total;                          // stride of pa/pw channels
    int cq_rt;                          // channel quads of the generic (CQ == 5) variant
    const int* hint_face;               // [Ns]
    const float* tref;                  // [Ns]
    const float* lab12;                 // [Ns][L][12]
    const float* absw;                  // [N]
    const float* face_feat;             // face-packed target features
    IcoDev ico;
    HintDev hint;
    int simmeasure;
    float* out;                         // [L][Ns]
    PeerSet peers;                      // optional (world > 0): label-major [L][n_full] tables, one per rank (peer mappings)
    int n_full;
};

__device__ __forceinline__ double group8_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// The single-item tail and the two-in-flight body of unary_items_uni must give the SAME bits for an item (which of the two
// handles it depends on the label-group size, i.e. on the shard), so every contraction is spelled out.
__device__ __forceinline__ void item_coords(const float4& rc, const float2& cc, const float4& m0, const float4& m1, const float4& m2,
                                            float& x, float& y, float& z) {
    x = rc.w + (m2.y + fmaf(m0.z, rc.z, fmaf(m0.y, rc.y, m0.x * rc.x)));
    y = cc.x + (m2.z + fmaf(m1.y, rc.z, fmaf(m1.x, rc.y, m0.w * rc.x)));
    z = cc.y + (m2.w + fmaf(m2.x, rc.z, fmaf(m1.w, rc.y, m1.z * rc.x)));
}
__device__ __forceinline__ float item_value(const float4& f, float tref, float wb, float wc) {
    return ((f.x - tref) + f.y) + fmaf(wb, f.z, wc * f.w);
}

// Univariate phase 2 with the number of remaining levels known at compile time and TWO items per thread in flight:
// the two descents are independent dependency chains (each level waits on a 16-byte table read), so interleaving them
// hides that latency without needing more resident warps.
template <int NLEV>
__device__ __forceinline__ void unary_items_uni(const UnaryArgs& a, const float* __restrict__ sK, const float4* __restrict__ rc4,
                                                const float2* __restrict__ c2, float* __restrict__ tb, int items, int P, int face0,
                                                float tref, int tid) {
    const float invP = 1.0f / (float)max(P, 1);
    const int hshift = 2 * a.hint.h;
    const float4* __restrict__ mtab = a.ico.mtab;
    const int leaf_off = a.ico.leaf_off, nleaf = a.ico.nleaf_per_base, node_off = a.hint.node_off;
    const float4* __restrict__ ff = reinterpret_cast<const float4*>(a.face_feat);
    for (int jb = 0; jb < items; jb += 512) {   // block-uniform trip count: the warp votes below see every lane
        const int j0 = jb + tid, j1 = j0 + 256;
        const bool has0 = j0 < items, has1 = j1 < items;
        if (!__any_sync(0xffffffffu, has1)) {
            // tail: nobody in this warp has a second item — plain single-item path
            if (!has0) continue;
            const int l = (int)(((float)j0 + 0.5f) * invP);
            const int i = j0 - l * P;
            const float4 rc = rc4[i];
            const float2 cc = c2[i];
            const float4* M = reinterpret_cast<const float4*>(sK + l * 12);
            const float4 m0 = M[0], m1 = M[1], m2 = M[2];
            float x, y, z;
            item_coords(rc, cc, m0, m1, m2, x, y, z);
            int face = face0;
            if (fminf(x, fminf(y, z)) < 0.f) face = ico_walk(a.hint, face0, x, y, z);
            const int b = face >> hshift;
            unsigned nn = (unsigned)(node_off + (face - (b << hshift)));
#pragma unroll
            for (int lv = 0; lv < NLEV; lv++) ico_descend_level(mtab, nn, x, y, z);
            const float inv = __fdividef(1.0f, x + y + z);
            const float4 f = __ldg(ff + (b * nleaf + ((int)nn - leaf_off)));
            tb[j0] = item_value(f, tref, y * inv, z * inv);
            continue;
        }
        const int jj[2] = {has0 ? j0 : 0, has1 ? j1 : 0};   // inactive lanes recompute item 0 and do not store
        float al[2], be[2], ga[2];
        unsigned n[2];
        int base[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int l = (int)(((float)jj[u] + 0.5f) * invP);
            const int i = jj[u] - l * P;
            const float4 rc = rc4[i];
            const float2 cc = c2[i];
            const float4* M = reinterpret_cast<const float4*>(sK + l * 12);
            const float4 m0 = M[0], m1 = M[1], m2 = M[2];
            item_coords(rc, cc, m0, m1, m2, al[u], be[u], ga[u]);
            int face = face0;
            if (fminf(al[u], fminf(be[u], ga[u])) < 0.f) face = ico_walk(a.hint, face0, al[u], be[u], ga[u]);
            base[u] = face >> hshift;
            n[u] = (unsigned)(node_off + (face - (base[u] << hshift)));
        }
#pragma unroll
        for (int lv = 0; lv < NLEV; lv++) {
            ico_descend_level(mtab, n[0], al[0], be[0], ga[0]);
            ico_descend_level(mtab, n[1], al[1], be[1], ga[1]);
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int leaf = base[u] * nleaf + ((int)n[u] - leaf_off);
            const float inv = __fdividef(1.0f, al[u] + be[u] + ga[u]);
            const float4 f = __ldg(ff + leaf);
            const float t = item_value(f, tref, be[u] * inv, ga[u] * inv);
            if (u == 0 ? has0 : has1) tb[jj[u]] = t;
        }
    }
}

template <int CQ>
__device__ __forceinline__ void unary_body(const UnaryArgs& a) {
    // padded channel count staged in shared memory; CQ == 5 is the generic variant (any channel count, a.cq_rt quads):
    // source features / weights are then read straight from the packed global streams (coalesced, L1 resident)
    constexpr int CP_ = CQ == 0 ? 1 : (CQ <= 4 ? 4 * CQ : 0);
    extern __shared__ float smem[];
    const int row = blockIdx.x / a.G, g = blockIdx.x - row * a.G;
    const int k = a.k_begin + row;
    const int l0 = g * a.Lg, l1 = min(a.L, l0 + a.Lg), nl = l1 - l0;
    const int s0 = a.off[row], P = a.end[row] - s0;
    const int tid = threadIdx.x;
    // smem carve (16-byte aligned pieces first): lab[Lg][12] | rc4[P] = (rx,ry,rz,cx) | c2[P] = (cy,cz) | a[CP_][P] |
    // w[CP_][P] | tbuf[Lg][P]
    float* sK = smem;
    float4* rc4 = reinterpret_cast<float4*>(sK + a.Lg * 12);
    float2* c2 = reinterpret_cast<float2*>(rc4 + P);
    float* sa = reinterpret_cast<float*>(c2 + P); float* sw = sa + CP_ * P;
    float* tb = sw + CP_ * P;

    for (int i = tid; i < P; i += 256) {
        rc4[i] = make_float4(a.px[s0 + i], a.py[s0 + i], a.pz[s0 + i], a.pcx[s0 + i]);
        c2[i] = make_float2(a.pcy[s0 + i], a.pcz[s0 + i]);
#pragma unroll
        for (int c = 0; c < CP_; c++) {
            sa[c * P + i] = a.pa[(size_t)c * a.total + s0 + i];
            sw[c * P + i] = a.pw[(size_t)c * a.total + s0 + i];
        }
    }
    for (int i = tid; i < nl * 12; i += 256) sK[i] = a.lab12[((size_t)row * a.L + l0) * 12 + i];
    const int face0 = a.hint_face[row];
    const float tref = a.tref[row];
    __syncthreads();
    const int items = nl * P;
    const float invP = 1.0f / (float)max(P, 1);
    const int hshift = 2 * a.hint.h;
    const int nlev = a.ico.depth - a.hint.h;
    const float4* __restrict__ mtab = a.ico.mtab;
    const int leaf_off = a.ico.leaf_off, nleaf = a.ico.nleaf_per_base, node_off = a.hint.node_off;

    bool done = false;
    if constexpr (CQ == 0) {
        done = true;
        switch (nlev) {  // uniform across the launch
            case 1: unary_items_uni<1>(a, sK, rc4, c2, tb, items, P, face0, tref, tid); break;
            case 2: unary_items_uni<2>(a, sK, rc4, c2, tb, items, P, face0, tref, tid); break;
            case 3: unary_items_uni<3>(a, sK, rc4, c2, tb, items, P, face0, tref, tid); break;
            case 4: unary_items_uni<4>(a, sK, rc4, c2, tb, items, P, face0, tref, tid); break;
            case 5: unary_items_uni<5>(a, sK, rc4, c2, tb, items, P, face0, tref, tid); break;
            case 6: unary_items_uni<6>(a, sK, rc4, c2, tb, items, P, face0, tref, tid); break;
            case 7: unary_items_uni<7>(a, sK, rc4, c2, tb, items, P, face0, tref, tid); break;
            default: done = false;
        }
    }
    for (int j = done ? items : tid; j < items; j += 256) {
        const int l = (int)(((float)j + 0.5f) * invP);
        const int i = j - l * P;
        const float4 rc = rc4[i];
        const float2 cc = c2[i];
        const float x = rc.x, y = rc.y, z = rc.z;
        const float4* M = reinterpret_cast<const float4*>(sK + l * 12);
        const float4 m0 = M[0], m1 = M[1], m2 = M[2];   // (M00 M01 M02 M10) (M11 M12 M20 M21) (M22 g0 g1 g2)
        float al = rc.w + (m2.y + (m0.x * x + m0.y * y + m0.z * z));
        float be = cc.x + (m2.z + (m0.w * x + m1.x * y + m1.y * z));
        float ga = cc.y + (m2.w + (m1.z * x + m1.w * y + m2.x * z));
        int face = face0;
        if (fminf(al, fminf(be, ga)) < 0.f) face = ico_walk(a.hint, face0, al, be, ga);
        const int base = face >> hshift;
        const int n = ico_descend_from(mtab, nlev, node_off + (face - (base << hshift)), al, be, ga);
        const int leaf = base * nleaf + (n - leaf_off);
        const float inv = __fdividef(1.0f, al + be + ga);
        if constexpr (CQ == 0) {
            // face-packed (f0_hi, f0_lo, f1-f0, f2-f0): t - tref = (f0_hi - tref) + f0_lo + wb (f1-f0) + wc (f2-f0)
            const float4 f = __ldg(reinterpret_cast<const float4*>(a.face_feat) + leaf);
            tb[j] = ((f.x - tref) + f.y) + ((be * inv) * f.z + (ga * inv) * f.w);
        } else if constexpr (CQ > 4) {
            // generic multivariate (more than 16 channels): two passes over the channels, re-gathering the target
            const int cq = a.cq_rt;
            const float4* ff = reinterpret_cast<const float4*>(a.face_feat) + (size_t)leaf * 3 * cq;
            const float wa = al * inv, wb = be * inv, wc = ga * inv;
            const float* paj = a.pa + s0 + i;
            const float* pwj = a.pw + s0 + i;
            float W = 0.f, Sa = 0.f, St = 0.f;
            for (int q = 0; q < cq; q++) {
                const float4 f0 = __ldg(ff + q), f1 = __ldg(ff + cq + q), f2 = __ldg(ff + 2 * cq + q);
                const float t4[4] = {wa * f0.x + wb * f1.x + wc * f2.x, wa * f0.y + wb * f1.y + wc * f2.y,
                                     wa * f0.z + wb * f1.z + wc * f2.z, wa * f0.w + wb * f1.w + wc * f2.w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const float w = pwj[(size_t)(4 * q + e) * a.total];
                    W += w; Sa += w * paj[(size_t)(4 * q + e) * a.total]; St += w * t4[e];
                }
            }
            if (W > 0.f) { Sa /= W; St /= W; }
            float pr = 0.f, va = 0.f, vt = 0.f;
            for (int q = 0; q < cq; q++) {
                const float4 f0 = __ldg(ff + q), f1 = __ldg(ff + cq + q), f2 = __ldg(ff + 2 * cq + q);
                const float t4[4] = {wa * f0.x + wb * f1.x + wc * f2.x, wa * f0.y + wb * f1.y + wc * f2.y,
                                     wa * f0.z + wb * f1.z + wc * f2.z, wa * f0.w + wb * f1.w + wc * f2.w};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const float w = pwj[(size_t)(4 * q + e) * a.total];
                    const float da = paj[(size_t)(4 * q + e) * a.total] - Sa, dt = t4[e] - St;
                    pr += w * da * dt; va += w * da * da; vt += w * dt * dt;
                }
            }
            if (W > 0.f) { pr /= W; va /= W; vt /= W; }
            const float rho = (va == 0.f || vt == 0.f) ? 0.f : pr / (sqrtf(va) * sqrtf(vt));
            tb[j] = 1.0f - (1.0f + rho) * 0.5f;
        } else {
            // multivariate: Pearson across the channels of this one sample (src/DiscreteCostFunction.cpp:515-518)
            const float4* ff = reinterpret_cast<const float4*>(a.face_feat) + (size_t)leaf * 3 * CQ;
            float t[CP_ > 0 ? CP_ : 1];
            const float wa = al * inv, wb = be * inv, wc = ga * inv;
#pragma unroll
            for (int q = 0; q < CQ; q++) {
                const float4 f0 = __ldg(ff + q), f1 = __ldg(ff + CQ + q), f2 = __ldg(ff + 2 * CQ + q);
                t[4 * q + 0] = wa * f0.x + wb * f1.x + wc * f2.x;
                t[4 * q + 1] = wa * f0.y + wb * f1.y + wc * f2.y;
                t[4 * q + 2] = wa * f0.z + wb * f1.z + wc * f2.z;
                t[4 * q + 3] = wa * f0.w + wb * f1.w + wc * f2.w;
            }
            float W = 0.f, Sa = 0.f, St = 0.f;
#pragma unroll
            for (int c = 0; c < CP_; c++) {
                const float w = sw[c * P + i];
                W += w; Sa += w * sa[c * P + i]; St += w * t[c];
            }
            if (W > 0.f) { Sa /= W; St /= W; }
            float pr = 0.f, va = 0.f, vt = 0.f;
#pragma unroll
            for (int c = 0; c < CP_; c++) {
                const float w = sw[c * P + i];
                const float da = sa[c * P + i] - Sa, dt = t[c] - St;
                pr += w * da * dt; va += w * da * da; vt += w * dt * dt;
            }
            if (W > 0.f) { pr /= W; va /= W; vt /= W; }
            const float rho = (va == 0.f || vt == 0.f) ? 0.f : pr / (sqrtf(va) * sqrtf(vt));
            tb[j] = 1.0f - (1.0f + rho) * 0.5f;
        }
    }
    __syncthreads();

    // Phase 3: 8 lanes per label.
    const float AW = a.absw[k];
    const bool pos = (a.simmeasure == 1 || a.simmeasure == 2);
    const int grp = tid >> 3, sub = tid & 7;
    for (int lb = 0; lb < nl; lb += 32) {   // warp-uniform trip count
        const int l = lb + grp;
        const bool act = l < nl;
        const float* t = tb + (act ? l : 0) * P;
        double cost;
        if (CQ == 0) {
            // weighted Pearson (src/similarities.cc:122-173) from single-pass moments of the SHIFTED data
            double W = 0, Sa = 0, Saa = 0, St = 0, Stt = 0, Sat = 0;
            if (act)
                for (int i = sub; i < P; i += 8) {
                    const double w = sw[i], av = sa[i], tv = t[i];
                    const double wa = w * av, wt = w * tv;
                    W += w; Sa += wa; St += wt; Saa += wa * av; Stt += wt * tv; Sat += wa * tv;
                }
            W = group8_sum(W); Sa = group8_sum(Sa); St = group8_sum(St);
            Saa = group8_sum(Saa); Stt = group8_sum(Stt); Sat = group8_sum(Sat);
            double ma = Sa, mt = St;                       // :139-142 means are divided only when sum > 0
            if (W > 0) { const double iw = 1.0 / W; ma *= iw; mt *= iw; }
            double pr = Sat - ma * St - mt * Sa + ma * mt * W;   // = sum w (a-ma)(t-mt)
            double va = Saa - 2.0 * ma * Sa + ma * ma * W;
            double vt = Stt - 2.0 * mt * St + mt * mt * W;
            // (the common 1/W of :164-169 cancels in the ratio)
            // :171 "either variance is zero -> 0": a variance below 1e-14 of the raw second moment is rounding noise
            double rho = 0.0;
            if (va > 1e-14 * Saa && vt > 1e-14 * Stt) rho = pr * rsqrt(va) * rsqrt(vt);
            cost = (double)AW * (1.0 - (1.0 + rho) * 0.5);
        } else {
            double s = 0;
            if (act)
                for (int i = sub; i < P; i += 8) s += (double)t[i];
            s = group8_sum(s);
            if (P > 0) s /= (double)P;
            cost = (double)AW * s;
        }
        if (!pos) cost = -cost;
        if (act && sub == 0) {
            if (a.peers.world > 0) {
                // fused exchange (control-point sharding): the finished entry goes straight into the label-major [L][N]
                // table of EVERY rank through its peer mapping (NVLink stores) — no gather, no re-layout afterwards
                const size_t at = (size_t)(l0 + l) * a.n_full + k;
                if (a.peers.mc) multicast_store(a.peers.mc + at, (float)cost);
                else
                    for (int r = 0; r < a.peers.world; r++) a.peers.p[r][at] = (float)cost;
            } else {
                a.out[(size_t)(l0 + l) * a.Ns + row] = (float)cost;
            }
        }
    }
}

template <int CQ>
__global__ void __launch_bounds__(256) k_unary(UnaryArgs a) { unary_body<CQ>(a); }
// univariate: 6 CTAs per SM (40 registers, no spills) instead of the 5 the default heuristics settle on
template <>
__global__ void __launch_bounds__(256, 6) k_unary<0>(UnaryArgs a) { unary_body<0>(a); }

}  // namespace msm
