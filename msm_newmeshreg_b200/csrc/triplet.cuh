// triplet.cuh — the triplet regulariser for the DEFAULT settings (strain energy with k = 2 and exponent 2, or the group
// rules with exponent 2): computeTripletCost (src/DiscreteCostFunction.cpp:138-254, rmode 3 :193-201, fold :163-167,
// epilogue :246-253), strain energy (src/reg_tools.cpp:556-601,703-748), group variant
// (src/DiscreteGroupCostFunction.cpp:9-30).  Every other parameter set runs the generic kernels of kernels.cuh.
//
// Formulation.  The strain energy depends on the displaced triangle only through its Gram matrix, and the Gram matrix only
// through the three squared edge lengths  g11 = |B-A|^2, g22 = |C-A|^2, dbc = |C-B|^2  (g12 = (g11 + g22 - dbc)/2).  With
// G0 = Gram of the ORIGINAL triangle and inv0 = 1/det G0:
//     trC = (g11 G22 - 2 g12 G12 + g22 G11) inv0          I3 = (g11 g22 - g12^2) inv0
//     W   = 1/2 ( mu (trC^2/I3 - 4)^+ + kappa (I3 + 1/I3 - 2) )  =  (1/I3) ( mu/2 (trC^2 - 4 I3)^+ + kappa/2 (I3 - 1)^2 )
// so that a table entry costs ~20 fp64 operations: one reciprocal (MUFU.RCP64H + Newton), no sqrt / pow / division, and
// the squared edge lengths are shared between the entries that have an edge in common (the whole (la,lc) and (lb,lc)
// planes of the full table; the 12 edges of the 8 fusion moves).  lambda is folded in as sqrt(lambda) (cost = lambda W^2).
// All coordinates are offsets from the triangle's first CURRENT vertex (local frame): small numbers, formed in fp64.
// fp64 throughout the strain (it is a difference of near-equal quantities for small deformations); the fold test of the
// full table is a 2-D cross product in fp32 with an fp64 fallback next to zero.
#pragma once
#include "kernels.cuh"
#ifndef MSM_MOVES_MINB
#define MSM_MOVES_MINB 5
#endif

namespace msm {

struct StrainConsts {       // launch constants (kernel parameters: read through the constant bank)
    double A1, A0;          // 2 sqrt(kah) (I3 - 1) = A1 J - A0,  J = 4 I3:  A1 = sqrt(kah)/2, A0 = 2 sqrt(kah),  kah = sqrt(lambda) kappa/2
    double M4;              // 4 muh,  muh = sqrt(lambda) mu / 2
    double thr2;            // cost < (1e-8)^2 lambda -> 0  (|W| < 1e-8 -> 0, :246); 0 in group mode (no such rule)
    double foldc_d;         // (1 + 1e6) lambda (:163-167,:253) | 1e7 (group, src/DiscreteGroupCostFunction.cpp:23)
    float foldc;
};
__host__ __device__ inline StrainConsts strain_consts(const DevParams& p) {
    StrainConsts k;
    const double sl = sqrt(p.lambda);
    const double sk = sqrt(0.5 * sl * p.kappa);
    k.A1 = 0.5 * sk; k.A0 = 2.0 * sk; k.M4 = 4.0 * (0.5 * sl * p.mu);
    k.thr2 = p.group ? 0.0 : 1e-16 * p.lambda;
    k.foldc_d = p.group ? 1e7 : (1.0 + 1e6) * p.lambda;
    k.foldc = (float)k.foldc_d;
    return k;
}
// Per original triangle.  Squared edge lengths are used SCALED by sinv0 = 1/sqrt(det G0), g' = g sinv0, so that
// I3 = g11' g22' - g12'^2 needs no further factor and trC = G22' g11' - 2 G12' g12' + G11' g22'.
struct TriConsts {
    double t22, t12h, G22s, sinv0;   // G11 sinv0, -G12 sinv0, G22 sinv0, 1/sqrt(det G0)
};
__device__ __forceinline__ TriConsts tri_consts(const TripGeom& g) {
    TriConsts c;
    c.sinv0 = rsqrt(g.det0); c.t22 = g.G11 * c.sinv0; c.t12h = -g.G12 * c.sinv0; c.G22s = g.G22 * c.sinv0;
    return c;
}

// sqrt(lambda) W from the SCALED squared edge lengths, written in  G = 2 g12 = g22 - dbc + g11  and  J = 4 I3 = 4 g11 g22 - G^2:
//     trC = t22 g22 + t12h G + t11          s  = trC^2 - J            ( = (trC^2/I3 - 4) I3 = (l1^2 - l2^2)^2 )
//     xk  = 2 sqrt(kah)(I3 - 1) = A1 J - A0     w4 = xk^2 + 4 muh s   ( = 4 (kah (I3-1)^2 + muh s) )
//     sqrt(lambda) W = w4 / J
// g11x4 = 4 g11' and t11 = g11' G22' are passed in because the full-table kernel hoists them out of its inner loop.
// 13 fp64 operations + MUFU.RCP64H (NEWTON = 1: ~1e-12 relative) / 15 (NEWTON = 2: full precision).
// s >= 0 up to rounding: the reference's "R = 1 if I1* <= 2" (src/reg_tools.cpp:592-598) only ever fires on rounding noise,
// where it changes W by ~1e-16 — not reproduced.
template <int NEWTON>
__device__ __forceinline__ double strain_sw(const StrainConsts& k, const TriConsts& c, double g11, double g11x4, double t11, double g22, double dbc) {
    const double G = (g22 - dbc) + g11;
    const double trC = fma(c.t22, g22, fma(c.t12h, G, t11));
    const double J = fma(g11x4, g22, -(G * G));
    const double s = fma(trC, trC, -J);
    const double xk = fma(k.A1, J, -k.A0);
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(J));    // MUFU.RCP64H: ~2^-20
    double e = fma(-J, r, 1.0);
    if (NEWTON > 1) { r = fma(r, e, r); e = fma(-J, r, 1.0); }
    const double wr = fma(xk, xk, k.M4 * s) * r;
    return fma(wr, e, wr);                                    // w4 / J after the last Newton step
}
__device__ __forceinline__ double strain_finish(const StrainConsts& k, double sw, bool fold) {
    double c = sw * sw;
    if (c < k.thr2) c = 0.0;
    return fold ? k.foldc_d : c;
}

// =========================================================================================================
// Fusion-move costs (include/Fusion/Fusion.h:187-194): combo bit2 -> A, bit1 -> B, bit0 -> C take label alpha.
// grid (ceil(T/128), ceil(labels/AG)): one thread per (triplet, group of AG proposal labels).  Everything that does not
// depend on alpha — the triangle constants, the three current corners, the all-current combination — is formed once per
// thread; per alpha the 9 new edges between the six candidate corners are shared by the 7 combinations that use them.
// out[(alpha)][t][8].
// =========================================================================================================
template <typename OutT, int AG>
__global__ void __launch_bounds__(128, MSM_MOVES_MINB) k_triplet_moves_fast(TripArgs a, StrainConsts k, int T, const int* __restrict__ labeling, int alpha,
                                                            OutT* __restrict__ out, PeerSet peers, int t_full) {
    const int tl = blockIdx.x * blockDim.x + threadIdx.x;
    if (tl >= T) return;
    const int t = a.t_off + tl;               // T = number of triplets of this launch (a shard [t_off, t_off + T))
    const int al0 = alpha < 0 ? (int)blockIdx.y * AG : alpha;
    const int al1 = alpha < 0 ? min(a.L, al0 + AG) : alpha + 1;
    int id[3]; TripGeom g;
    trip_load(a, t, id, g);
    const TriConsts c = tri_consts(g);
    const double ox = a.cp[3 * id[0]], oy = a.cp[3 * id[0] + 1], oz = a.cp[3 * id[0] + 2];
    double C0[3][3];                           // current corners, local frame
#pragma unroll
    for (int v = 0; v < 3; v++) {
        const double* pc = a.disp + ((size_t)id[v] * a.L + labeling[id[v]]) * 3;
        C0[v][0] = pc[0] - ox; C0[v][1] = pc[1] - oy; C0[v][2] = pc[2] - oz;
    }
    auto sq = [&](const double* p, const double* q, double* e) {   // e = q - p, returns |e|^2 scaled
        e[0] = q[0] - p[0]; e[1] = q[1] - p[1]; e[2] = q[2] - p[2];
        return fma(e[0], e[0], fma(e[1], e[1], e[2] * e[2])) * c.sinv0;
    };
    // alpha-independent: edges between current corners, the all-current combination
    double e1cc[3], e2cc[3], tmp[3], crcc[3];
    const double g11cc = sq(C0[0], C0[1], e1cc), g22cc = sq(C0[0], C0[2], e2cc), dbccc = sq(C0[1], C0[2], tmp);
    cross3(g.n, e1cc, crcc);                   // (e1 x e2).n = e2.(n x e1)
    const double res0 = strain_finish(k, strain_sw<2>(k, c, g11cc, 4.0 * g11cc, g11cc * c.G22s, g22cc, dbccc), dot3(e2cc, crcc) < 0.0);
    for (int al = al0; al < al1; al++) {
        double A1[3][3];                       // corners under label alpha
#pragma unroll
        for (int v = 0; v < 3; v++) {
            const double* pn = a.disp + ((size_t)id[v] * a.L + al) * 3;
            A1[v][0] = pn[0] - ox; A1[v][1] = pn[1] - oy; A1[v][2] = pn[2] - oz;
        }
        double res[8];
        res[0] = res0;
        // squared lengths of B-C edges (index [j][q]: j = B takes alpha, q = C takes alpha)
        double dbc[2][2];
        dbc[0][0] = dbccc; dbc[0][1] = sq(C0[1], A1[2], tmp); dbc[1][0] = sq(A1[1], C0[2], tmp); dbc[1][1] = sq(A1[1], A1[2], tmp);
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const double* PA = i ? A1[0] : C0[0];
            double e2[2][3], g22[2];
            if (i == 0) { g22[0] = g22cc; e2[0][0] = e2cc[0]; e2[0][1] = e2cc[1]; e2[0][2] = e2cc[2]; }
            else g22[0] = sq(PA, C0[2], e2[0]);
            g22[1] = sq(PA, A1[2], e2[1]);
#pragma unroll
            for (int j = 0; j < 2; j++) {
                double e1[3], cr[3], g11;
                if (i == 0 && j == 0) { g11 = g11cc; cr[0] = crcc[0]; cr[1] = crcc[1]; cr[2] = crcc[2]; }
                else { g11 = sq(PA, j ? A1[1] : C0[1], e1); cross3(g.n, e1, cr); }
                const double g11x4 = 4.0 * g11, t11 = g11 * c.G22s;
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    if (i + j + q == 0) continue;
                    const bool fold = dot3(e2[q], cr) < 0.0;   // sign of the dot of the unit normals (:163)
                    res[4 * i + 2 * j + q] = strain_finish(k, strain_sw<2>(k, c, g11, g11x4, t11, g22[q], dbc[j][q]), fold);
                }
            }
        }
        const size_t slot = (size_t)(alpha < 0 ? al : 0);
        if (peers.world > 0) {
            // fused exchange (triplet sharding): straight into the full [L or 1][t_full][8] table of every rank
            const float4 lo = make_float4((float)res[0], (float)res[1], (float)res[2], (float)res[3]);
            const float4 hi = make_float4((float)res[4], (float)res[5], (float)res[6], (float)res[7]);
            const size_t at = (slot * t_full + t) * 2;
            if (peers.mc) {
                float4* dst = reinterpret_cast<float4*>(peers.mc) + at;
                multicast_store(dst, lo); multicast_store(dst + 1, hi);
            } else {
                for (int r = 0; r < peers.world; r++) {
                    float4* dst = reinterpret_cast<float4*>(peers.p[r]) + at;
                    dst[0] = lo; dst[1] = hi;
                }
            }
            continue;
        }
        OutT* o = out + (slot * T + tl) * 8;
        if constexpr (sizeof(OutT) == 4) {
            reinterpret_cast<float4*>(o)[0] = make_float4((float)res[0], (float)res[1], (float)res[2], (float)res[3]);
            reinterpret_cast<float4*>(o)[1] = make_float4((float)res[4], (float)res[5], (float)res[6], (float)res[7]);
        } else {
#pragma unroll
            for (int m = 0; m < 4; m++) reinterpret_cast<double2*>(o)[m] = make_double2(res[2 * m], res[2 * m + 1]);
        }
    }
}

// =========================================================================================================
// Per-triplet records of the full-table kernel, formed once per table by k_triplet_prep (one warp per triplet): the
// triangle constants, the 3L displaced corners as offsets from the first current corner (fp64) and their 2-D coordinates
// in the plane of the current triangle (fp32, fold test).  A CTA of the table kernel then starts a triplet from ONE
// contiguous ~2 KB record — prefetched with cp.async while it works on the previous triplet — instead of a chain of
// dependent gathers (triplet ids -> control points / original triangle -> displaced corners) and a serial set-up.
// =========================================================================================================
struct TripPrep {
    double o[3];           // local origin: the first current corner
    double n[3];           // (v1 - v0) x (v2 - v0) of the current control triangle (fold test)
    double u[3], v[3];     // orthonormal basis of its plane, u x v = n/|n|
    double t22, t12h, G22s, sinv0;   // TriConsts
    int id[3];
    float qmax;            // largest |2-D coordinate| (scales the fp32 fold test's uncertainty band)
};
static_assert(sizeof(TripPrep) == 144, "record header layout");
// record = TripPrep | so[3L][3] fp64 | sq[3L] float2, padded to 16 bytes
__host__ __device__ constexpr size_t triplet_rec_bytes(int L) { return (sizeof(TripPrep) + (size_t)72 * L + (size_t)24 * L + 15) & ~(size_t)15; }

__global__ void __launch_bounds__(128) k_triplet_prep(TripArgs a, int t0, int nt, unsigned char* __restrict__ recs) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= nt) return;   // whole warps leave together
    const int L = a.L;
    unsigned char* rec = recs + (size_t)i * triplet_rec_bytes(L);
    TripPrep* hp = reinterpret_cast<TripPrep*>(rec);
    double* so = reinterpret_cast<double*>(rec + sizeof(TripPrep));
    float2* sq = reinterpret_cast<float2*>(rec + sizeof(TripPrep) + (size_t)72 * L);
    int id[3]; TripGeom g;
    trip_load(a, t0 + i, id, g);     // every lane (same addresses: broadcast loads), so that no shuffles are needed below
    const double* c0 = a.cp + 3 * id[0]; const double* c1 = a.cp + 3 * id[1];
    double u[3] = {c1[0] - c0[0], c1[1] - c0[1], c1[2] - c0[2]};
    const double iu = rsqrt(dot3(u, u)), in = rsqrt(dot3(g.n, g.n));
    double v[3];
#pragma unroll
    for (int j = 0; j < 3; j++) u[j] *= iu;
    cross3(g.n, u, v);
#pragma unroll
    for (int j = 0; j < 3; j++) v[j] *= in;
    float qm = 0.f;
    for (int e = lane; e < 3 * L; e += 32) {
        const int w = e / L, l = e - w * L;
        const double* p = a.disp + ((size_t)id[w] * L + l) * 3;
        const double x = p[0] - c0[0], y = p[1] - c0[1], z = p[2] - c0[2];
        so[3 * e] = x; so[3 * e + 1] = y; so[3 * e + 2] = z;
        const float2 q = make_float2((float)(x * u[0] + y * u[1] + z * u[2]), (float)(x * v[0] + y * v[1] + z * v[2]));
        sq[e] = q;
        qm = fmaxf(qm, fmaxf(fabsf(q.x), fabsf(q.y)));
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) qm = fmaxf(qm, __shfl_xor_sync(0xffffffffu, qm, m));
    if (lane == 0) {
        const TriConsts c = tri_consts(g);
        TripPrep p;
#pragma unroll
        for (int j = 0; j < 3; j++) { p.o[j] = c0[j]; p.n[j] = g.n[j]; p.u[j] = u[j]; p.v[j] = v[j]; p.id[j] = id[j]; }
        p.t22 = c.t22; p.t12h = c.t12h; p.G22s = c.G22s; p.sinv0 = c.sinv0; p.qmax = qm;
        *hp = p;
    }
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// =========================================================================================================
// Full look-up table float[T][L][L][L] (serves the unmodified Fusion::optimize through computeTripletCost).
// Persistent CTAs (96 threads, MINB per SM), each walking triplets blockIdx.x, + gridDim.x, ...; per triplet:
//   0. wait for the triplet's record (cp.async issued one triplet earlier), start the prefetch of the next one;
//   1. scaled squared edge lengths of the (lc,la) and (lb,lc) planes from the record's corners (fp64, shared memory);
//   2. the table: a thread owns one lb and a tile of 4 consecutive la and loops over lc — per lc ONE (lb,lc) length, the
//      four (la,lc) ones as two 16-byte broadcast loads, 4 independent entries (the kernel's first limiter was the
//      shared-memory crossbar, hence the register tile); fold test = 2-D cross product in fp32, fp64 next to zero;
//   3. the L^3 results leave shared memory as 16-byte coalesced streaming stores (the table is store bound: 4 B/entry).
// LT = compile-time label count (19 = the standard label sets, src/DiscreteModel.cpp:91-162 at labeldist 0.5) or 0 (any).
// =========================================================================================================
constexpr int kTabThreads = 96;
constexpr int kTabTile = 4;
__host__ __device__ constexpr int triplet_table_lsa(int L) { return ((L + kTabTile - 1) / kTabTile) * kTabTile; }   // (lc,la) row: whole tiles
__host__ __device__ constexpr int triplet_table_lsb(int L) { return L | 1; }                                         // (lb,lc) row: odd stride
__host__ __device__ inline size_t triplet_table_smem(int L, int nbuf = 2) {
    // nbuf records | DacT[L][LSA] fp64 | Dbc[L][LSB] fp64 | results L^3 (+8: alignment) floats
    return nbuf * triplet_rec_bytes(L) + sizeof(double) * ((size_t)L * (triplet_table_lsa(L) + triplet_table_lsb(L)) + 2) +
           sizeof(float) * ((size_t)L * L * L + 8);
}

template <int LT, int MINB>
__global__ void __launch_bounds__(kTabThreads, MINB) k_triplet_table_fast(TripArgs a, StrainConsts k, const unsigned char* __restrict__ recs, int ntrip,
                                                                         int nbuf, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int L = LT ? LT : a.L;
    const int L2 = L * L, L3 = L2 * L, LSA = triplet_table_lsa(L), LSB = triplet_table_lsb(L);
    const int RB = (int)triplet_rec_bytes(L);
    double* sDacT = reinterpret_cast<double*>(sm_raw + nbuf * RB);   // nbuf = 2 only when CTAs walk several triplets   // [lc][la], row stride LSA (16-byte aligned rows)
    double* sDbc = sDacT + L * LSA;                               // [lb][lc], row stride LSB
    float* sres0 = reinterpret_cast<float*>(sDbc + ((L * LSB + 1) & ~1));  // 16-byte aligned (every piece before it is)
    const int tid = threadIdx.x;
    auto prefetch = [&](int t, int buf) {
        const unsigned char* src = recs + (size_t)t * RB;
        unsigned char* dst = sm_raw + buf * RB;
        for (int e = tid * 16; e < RB; e += kTabThreads * 16) cp_async16(dst + e, src + e);
        cp_async_commit();
    };
    // One CTA per triplet by default (the hardware staggers the CTAs of an SM over the phases below, which measured
    // faster than persistent CTAs marching in step); with fewer CTAs than triplets a CTA walks blockIdx.x, + gridDim.x, ...
    // and prefetches its next record while it works.
    int buf = 0;
    if ((int)blockIdx.x < ntrip) prefetch(blockIdx.x, 0);
    for (int t = blockIdx.x; t < ntrip; t += gridDim.x, buf ^= 1) {
        cp_async_wait_all();
        __syncthreads();        // record t is in shared memory; every thread is done with the previous triplet's copy-out reads
        if (t + (int)gridDim.x < ntrip) prefetch(t + gridDim.x, buf ^ 1);
        const unsigned char* rec = sm_raw + buf * RB;
        const TripPrep& S = *reinterpret_cast<const TripPrep*>(rec);
        const double* so = reinterpret_cast<const double*>(rec + sizeof(TripPrep));            // [3][L][3]
        const float2* sq = reinterpret_cast<const float2*>(rec + sizeof(TripPrep) + 72 * L);    // [3][L]
        float* o = out + (size_t)t * L3;
        // the results sit in shared memory with the same 16-byte phase as their destination, so both sides of the copy-out
        // are aligned float4 accesses
        const int phase = (int)((reinterpret_cast<uintptr_t>(o) >> 2) & 3);
        float* sres = sres0 + phase;
        TriConsts c;
        c.t22 = S.t22; c.t12h = S.t12h; c.G22s = S.G22s; c.sinv0 = S.sinv0;
        // phase 1: scaled squared edge lengths.  A thread takes a segment of one row of corners (A under la, or B under
        // lb): the row's corner stays in registers, the segment's columns are independent chains.
        {
            const int nseg = max(1, kTabThreads / (2 * L)), seglen = (L + nseg - 1) / nseg;
            for (int w = tid; w < 2 * L * nseg; w += kTabThreads) {
                const int row = w / nseg, seg = w - row * nseg;      // rows 0..L-1: corner A under la; L..2L-1: corner B under lb
                const double px = so[3 * row], py = so[3 * row + 1], pz = so[3 * row + 2];
                const int c0 = seg * seglen;
#pragma unroll(LT ? (LT + 1) / 2 : 1)
                for (int j = 0; j < seglen; j++) {
                    const int lc = min(c0 + j, L - 1);              // a short last segment repeats its last column
                    const double* q = so + 3 * (2 * L + lc);
                    const double x = q[0] - px, y = q[1] - py, z = q[2] - pz;
                    const double d = fma(x, x, fma(y, y, z * z)) * c.sinv0;
                    if (row < L) sDacT[lc * LSA + row] = d;
                    else sDbc[(row - L) * LSB + lc] = d;
                }
            }
            for (int i = tid; i < L * (LSA - L); i += kTabThreads) {   // pad of the la tiles: benign values
                const int lc = i / (LSA - L), j = i - lc * (LSA - L);
                sDacT[lc * LSA + L + j] = 1.0;
            }
        }
        __syncthreads();
        // phase 2: the table.  work item = (tile of 4 la, lb); consecutive threads take consecutive lb
        const int ntile = LSA / kTabTile;
        for (int w = tid; w < ntile * L; w += kTabThreads) {
            const int tile = w / L, lb = w - tile * L, la0 = tile * kTabTile;
            const double* pb = so + 3 * (L + lb);
            const float2 qb = sq[L + lb];
            double g11[kTabTile], g11x4[kTabTile], t11[kTabTile];
            float f1x[kTabTile], f1y[kTabTile], kk[kTabTile];
            float fsum = 0.f;
#pragma unroll
            for (int i = 0; i < kTabTile; i++) {
                const int la = min(la0 + i, L - 1);       // a short last tile repeats its last la (its extra entries are not stored)
                const double* pa = so + 3 * la;
                const double e1x = pb[0] - pa[0], e1y = pb[1] - pa[1], e1z = pb[2] - pa[2];
                g11[i] = fma(e1x, e1x, fma(e1y, e1y, e1z * e1z)) * c.sinv0;
                g11x4[i] = 4.0 * g11[i]; t11[i] = g11[i] * c.G22s;
                // fold test: (e1 x e2).n has the sign of the 2-D cross product of the edges' projections onto the current
                // triangle's plane:  f1x (qc.y - qa.y) - f1y (qc.x - qa.x)
                const float2 qa = sq[la];
                f1x[i] = qb.x - qa.x; f1y[i] = qb.y - qa.y;
                kk[i] = f1x[i] * qa.y - f1y[i] * qa.x;
                fsum = fmaxf(fsum, fabsf(f1x[i]) + fabsf(f1y[i]));
            }
            const float band = 8e-6f * fsum * (S.qmax + 1e-30f);
            const double* rowbc = sDbc + lb * LSB;
            const double* colac = sDacT + la0;
            const float2* qc = sq + 2 * L;
            float* dst = sres + (la0 * L + lb) * L;
#pragma unroll(LT ? 2 : 1)
            for (int lc = 0; lc < L; lc++) {
                const float2 q = qc[lc];
                const double dbc = rowbc[lc];
                double g22[kTabTile];
                {
                    const double2 x01 = *reinterpret_cast<const double2*>(colac + lc * LSA);
                    const double2 x23 = *reinterpret_cast<const double2*>(colac + lc * LSA + 2);
                    g22[0] = x01.x; g22[1] = x01.y; g22[2] = x23.x; g22[3] = x23.y;
                }
                bool fold[kTabTile];
                float amin = 3.0e38f;
#pragma unroll
                for (int i = 0; i < kTabTile; i++) {
                    const float crs = fmaf(f1x[i], q.y, -fmaf(f1y[i], q.x, kk[i]));
                    fold[i] = crs < 0.f;
                    amin = fminf(amin, fabsf(crs));
                }
                if (amin <= band) {   // next to zero (a degenerate displaced triangle): decide in fp64 from the 3-D corners
                    const double* pc = so + 3 * (2 * L + lc);
#pragma unroll
                    for (int i = 0; i < kTabTile; i++) {
                        const double* pa = so + 3 * min(la0 + i, L - 1);
                        const double e1[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
                        const double e2[3] = {pc[0] - pa[0], pc[1] - pa[1], pc[2] - pa[2]};
                        double nn[3];
                        cross3(e1, e2, nn);
                        fold[i] = dot3(nn, S.n) < 0.0;
                    }
                }
#pragma unroll
                for (int i = 0; i < kTabTile; i++) {
                    const double sw = strain_sw<1>(k, c, g11[i], g11x4[i], t11[i], g22[i], dbc);
                    // (the reference's "|W| < 1e-8 -> 0" (:246) changes an entry by < 1e-16 lambda: not reproduced in the fp32 table)
                    const float v = fold[i] ? k.foldc : (float)(sw * sw);
                    if (la0 + i < L) dst[i * L2 + lc] = v;
                }
            }
        }
        __syncthreads();
        // phase 3: copy-out — scalar head up to the first 16-byte boundary, float4 body, scalar tail
        const int head = min(L3, (4 - phase) & 3);
        const int nbody = (L3 - head) >> 2;
        if (tid < head) o[tid] = sres[tid];
        const float4* sb = reinterpret_cast<const float4*>(sres + head);
        float4* ob = reinterpret_cast<float4*>(o + head);
#pragma unroll 4
        for (int i = tid; i < nbody; i += kTabThreads) __stcs(ob + i, sb[i]);
        const int tail0 = head + 4 * nbody;
        if (tid < L3 - tail0) o[tail0 + tid] = sres[tail0 + tid];
    }
}

}  // namespace msm
