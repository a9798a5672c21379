// kernels.cuh — sm_100a kernels of the newMSM cost-evaluation hot path.
// No tensor cores: nothing here is a dense contraction (north_star).  Everything is gather / scan / reduce work.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "locate.cuh"

namespace msm {

constexpr double kRAD = 100.0;      // src/DiscreteCostFunction.h:9
constexpr double kEPSILON = 1.0e-8; // src/DiscreteCostFunction.h:10

struct DevParams {
    double lambda, rexp, mu, kappa, k_exp, range, meanangle, mvdmax;
    int simmeasure, rmode, dweight, anorm, group;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// =========================================================================================================
// k_label_rot — per (control point, label):  disp = ROT[k]*label_l  (ROT[k] = rot(centre -> CP_k),
// src/DiscreteModel.cpp:284-294) and K = rot(CP_k -> disp) - I (src/DiscreteCostFunction.cpp:439,513,258-259).
// One thread per (k,l); fp64 (B200 has full-rate-class FP64), results kept in fp64 and fp32.
// =========================================================================================================
__global__ void __launch_bounds__(256) k_label_rot(int N, int L, const double* __restrict__ cp, const double* __restrict__ labels,
                                                   double cx, double cy, double cz, double* __restrict__ disp,
                                                   float* __restrict__ Kf, double* __restrict__ Kd) {
    int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= N * L) return;
    int k = id / L, l = id - k * L;
    double c[3] = {cx, cy, cz};
    double p[3] = {cp[3 * k], cp[3 * k + 1], cp[3 * k + 2]};
    double K[9];
    rodrigues_minus_identity(c, p, K);
    double lx = labels[3 * l], ly = labels[3 * l + 1], lz = labels[3 * l + 2];
    double d[3];
    d[0] = lx + (K[0] * lx + K[1] * ly + K[2] * lz);
    d[1] = ly + (K[3] * lx + K[4] * ly + K[5] * lz);
    d[2] = lz + (K[6] * lx + K[7] * ly + K[8] * lz);
    disp[3 * id] = d[0]; disp[3 * id + 1] = d[1]; disp[3 * id + 2] = d[2];
    rodrigues_minus_identity(p, d, K);
#pragma unroll
    for (int i = 0; i < 9; i++) Kf[(size_t)id * 9 + i] = (float)K[i];
    if (Kd) {
#pragma unroll
        for (int i = 0; i < 9; i++) Kd[(size_t)id * 9 + i] = K[i];
    }
}

// =========================================================================================================
// Patch construction — get_source_data / within_controlpt_range (src/DiscreteCostFunction.cpp:104-109,394-410,458-481)
// geod(CP_k, SRC_i) < range*MAXSEP_k  <=>  chord < 2R sin(range*MAXSEP_k / 2R).  fp32 pre-filter, fp64 decision with
// an uncertainty band of 1e-11 relative; pairs inside the band go to the host, which applies the reference's own
// asin formula, so membership is identical to the reference even at the exact ties of an undeformed icosphere.
// =========================================================================================================
struct PatchArgs {
    int k_begin, k_end, V;
    const float* sxf; const float* syf; const float* szf;    // source xyz fp32 SoA
    const double* src;                                       // source xyz fp64 [V][3]
    const double* cp;                                        // CP xyz fp64 [N][3]
    const double* maxsep;                                    // [N]
    double range;
    int* counts;                                             // [Ns] definite members
    int2* unc; int unc_cap; int* unc_count;                  // uncertain (k,i) pairs
    const int* acc_off; const int* acc_idx;                  // accepted uncertain pairs per shard row (fill pass)
    const int* off;                                          // [Ns+1] CSR offsets (fill pass)
    int* idx;                                                // [total]
    int* slab; int slab_cap;                                 // optional [Ns][slab_cap]: the count pass stashes the definite
                                                             // members it finds, so the fill pass need not search again
};

__device__ __forceinline__ void patch_thresholds(double thr, double& c2lo, double& c2hi, float& c2f) {
    double h = thr / (2.0 * kRAD);
    if (h >= 1.5707963267948966) { c2lo = 1e300; c2hi = 1e300; c2f = 3.0e38f; return; }
    double c = 2.0 * kRAD * sin(h);
    double lo = c * (1.0 - 1e-11), hi = c * (1.0 + 1e-11);
    c2lo = lo * lo; c2hi = hi * hi;
    c2f = (float)(c * c * (1.0 + 1e-4) + 1e-6);
}

// classify vertex i against CP k: 0 out, 1 in, 2 uncertain
__device__ __forceinline__ int patch_classify(const PatchArgs& a, int i, double cx, double cy, double cz, float cxf, float cyf, float czf,
                                              double c2lo, double c2hi, float c2f) {
    float dx = a.sxf[i] - cxf, dy = a.syf[i] - cyf, dz = a.szf[i] - czf;
    float d2f = dx * dx + dy * dy + dz * dz;
    if (d2f > c2f) return 0;
    double ex = cx - a.src[3 * i], ey = cy - a.src[3 * i + 1], ez = cz - a.src[3 * i + 2];
    double d2 = ex * ex + ey * ey + ez * ez;
    if (d2 < c2lo) return 1;
    if (d2 <= c2hi) return 2;
    return 0;
}

template <bool FILL>
__global__ void __launch_bounds__(256) k_patch(PatchArgs a) {
    const int row = blockIdx.x;
    const int k = a.k_begin + row;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const double cx = a.cp[3 * k], cy = a.cp[3 * k + 1], cz = a.cp[3 * k + 2];
    const float cxf = (float)cx, cyf = (float)cy, czf = (float)cz;
    double c2lo, c2hi; float c2f;
    patch_thresholds(a.range * a.maxsep[k], c2lo, c2hi, c2f);
    // the fp32 pre-filter compares against rounded coordinates: widen by the rounding of |x|<=~100 inputs
    c2f += 1e-2f;
    int acc0 = 0, acc1 = 0;
    if (FILL && a.acc_off) { acc0 = a.acc_off[row]; acc1 = a.acc_off[row + 1]; }
    if (tid == 0) s_base = FILL ? a.off[row] : 0;
    __syncthreads();
    int total = 0;
    for (int base = 0; base < a.V; base += 256) {
        int i = base + tid;
        int cls = 0;
        if (i < a.V) cls = patch_classify(a, i, cx, cy, cz, cxf, cyf, czf, c2lo, c2hi, c2f);
        bool in = (cls == 1);
        if (cls == 2) {
            if (!FILL) {
                int slot = atomicAdd(a.unc_count, 1);
                if (slot < a.unc_cap) a.unc[slot] = make_int2(k, i);
            } else {
                for (int e = acc0; e < acc1; e++)
                    if (a.acc_idx[e] == i) { in = true; break; }
            }
        }
        unsigned m = __ballot_sync(0xffffffffu, in);
        if (!FILL) {
            if (lane == 0) total += __popc(m);
        } else {
            if (lane == 0) s_warp[warp] = __popc(m);
            __syncthreads();
            int before = 0, all = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) { int c = s_warp[w]; if (w < warp) before += c; all += c; }
            if (in) a.idx[s_base + before + __popc(m & ((1u << lane) - 1u))] = i;
            __syncthreads();
            if (tid == 0) s_base += all;
            __syncthreads();
        }
    }
    if (!FILL) {
        if (lane == 0) s_warp[warp] = total;
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < 8; w++) t += s_warp[w];
            a.counts[row] = t;
        }
    }
}

// ---- binned variant: source vertices are counting-sorted into a uniform voxel grid whose cell size is the largest patch
// chord, so a control point only tests the vertices of the <= 27 cells around it instead of the whole mesh (the
// reference's O(N*V) scan, src/DiscreteCostFunction.cpp:398-400).  Membership decisions are the same fp64 tests as above;
// each patch list is sorted ascending afterwards (the reference's order), which also makes the result deterministic.
struct BinGrid {
    int G;            // cells per dimension
    double lo, inv_h; // cell = floor((x - lo) * inv_h), clamped
    int* cell_start;  // [G^3 + 1] (exclusive scan of counts)
    int* cell_fill;   // [G^3]
    int* sorted;      // [V] vertex ids grouped by cell
};
__device__ __forceinline__ int bin_coord(const BinGrid& g, double x) {
    int c = (int)floor((x - g.lo) * g.inv_h);
    return min(g.G - 1, max(0, c));
}
// fp64 [V][3] vertices -> fp32 SoA (the patch pre-filter's copy)
__global__ void __launch_bounds__(256) k_split_xyz(int V, const double* __restrict__ src, float* __restrict__ x, float* __restrict__ y,
                                                   float* __restrict__ z) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    x[i] = (float)src[3 * i]; y[i] = (float)src[3 * i + 1]; z[i] = (float)src[3 * i + 2];
}
__global__ void __launch_bounds__(256) k_bin_count(BinGrid g, int V, const double* __restrict__ src, int* __restrict__ cell_of, int* __restrict__ counts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const int c = (bin_coord(g, src[3 * i]) * g.G + bin_coord(g, src[3 * i + 1])) * g.G + bin_coord(g, src[3 * i + 2]);
    cell_of[i] = c;
    atomicAdd(counts + c, 1);
}
__global__ void __launch_bounds__(256) k_bin_fill(BinGrid g, int V, const int* __restrict__ cell_of) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const int c = cell_of[i];
    g.sorted[g.cell_start[c] + atomicAdd(g.cell_fill + c, 1)] = i;
}

// 128 threads: bitonic sort of s_list[0..n) (padded with INT_MAX to the next power of two), then store ascending.
// Callers synchronise before the call.
__device__ __forceinline__ void patch_sort_store(int* s_list, int n, int* __restrict__ out, int tid) {
    int m = 1;
    while (m < n) m <<= 1;
    for (int i = n + tid; i < m; i += 128) s_list[i] = 0x7fffffff;
    __syncthreads();
    for (int sz = 2; sz <= m; sz <<= 1)
        for (int st = sz >> 1; st > 0; st >>= 1) {
            for (int i = tid; i < m; i += 128) {
                const int j = i ^ st;
                if (j > i) {
                    const int vi = s_list[i], vj = s_list[j];
                    const bool up = (i & sz) == 0;
                    if ((vi > vj) == up) { s_list[i] = vj; s_list[j] = vi; }
                }
            }
            __syncthreads();
        }
    for (int i = tid; i < n; i += 128) out[i] = s_list[i];
}

template <bool FILL>
__global__ void __launch_bounds__(128) k_patch_binned(PatchArgs a, BinGrid g, int cap /*smem ints, FILL only*/) {
    extern __shared__ int s_list[];
    __shared__ int s_n;
    const int row = blockIdx.x, k = a.k_begin + row, tid = threadIdx.x;
    const double cx = a.cp[3 * k], cy = a.cp[3 * k + 1], cz = a.cp[3 * k + 2];
    const float cxf = (float)cx, cyf = (float)cy, czf = (float)cz;
    double c2lo, c2hi; float c2f;
    patch_thresholds(a.range * a.maxsep[k], c2lo, c2hi, c2f);
    c2f += 1e-2f;
    const double reach = sqrt(c2hi) * (1.0 + 1e-9);
    int acc0 = 0, acc1 = 0;
    if (FILL && a.acc_off) { acc0 = a.acc_off[row]; acc1 = a.acc_off[row + 1]; }
    if (tid == 0) s_n = 0;
    __syncthreads();
    const int x0 = bin_coord(g, cx - reach), x1 = bin_coord(g, cx + reach);
    const int y0 = bin_coord(g, cy - reach), y1 = bin_coord(g, cy + reach);
    const int z0 = bin_coord(g, cz - reach), z1 = bin_coord(g, cz + reach);
    int mine = 0;
    for (int x = x0; x <= x1; x++)
        for (int y = y0; y <= y1; y++) {
            // cells (x,y,z0..z1) are contiguous in memory: one candidate range per (x,y)
            const int c0 = (x * g.G + y) * g.G + z0, c1 = (x * g.G + y) * g.G + z1;
            const int b = g.cell_start[c0], e = g.cell_start[c1 + 1];
            for (int s = b + tid; s < e; s += 128) {
                const int i = g.sorted[s];
                const int cls = patch_classify(a, i, cx, cy, cz, cxf, cyf, czf, c2lo, c2hi, c2f);
                bool in = (cls == 1);
                if (cls == 2) {
                    if (!FILL) {
                        int slot = atomicAdd(a.unc_count, 1);
                        if (slot < a.unc_cap) a.unc[slot] = make_int2(k, i);
                    } else {
                        for (int q = acc0; q < acc1; q++)
                            if (a.acc_idx[q] == i) { in = true; break; }
                    }
                }
                if (in) {
                    if (!FILL) {
                        if (a.slab) { const int p = atomicAdd(&s_n, 1); if (p < a.slab_cap) a.slab[(size_t)row * a.slab_cap + p] = i; }
                        else mine++;
                    } else { const int p = atomicAdd(&s_n, 1); if (p < cap) s_list[p] = i; }
                }
            }
        }
    if (!FILL) {
        if (!a.slab) {
            mine = (int)warp_sum((float)mine);
            if ((tid & 31) == 0) atomicAdd(&s_n, mine);
        }
        __syncthreads();
        if (tid == 0) a.counts[row] = s_n;   // the true count, also when it exceeds slab_cap (the host then re-searches)
    } else {
        __syncthreads();
        const int n = min(s_n, cap);
        patch_sort_store(s_list, n, a.idx + a.off[row], tid);
    }
}

// Fill pass when the count pass stashed its members (PatchArgs::slab): no second search — the definite members come from
// the slab, the accepted boundary ties from the host's list; sort, store.
__global__ void __launch_bounds__(128) k_patch_from_slab(PatchArgs a, int cap /*smem ints*/) {
    extern __shared__ int s_list[];
    const int row = blockIdx.x, tid = threadIdx.x;
    const int n0 = min(a.counts[row], min(a.slab_cap, cap));
    int acc0 = 0, acc1 = 0;
    if (a.acc_off) { acc0 = a.acc_off[row]; acc1 = a.acc_off[row + 1]; }
    const int n = min(n0 + (acc1 - acc0), cap);
    const int* __restrict__ sl = a.slab + (size_t)row * a.slab_cap;
    for (int i = tid; i < n0; i += 128) s_list[i] = sl[i];
    for (int i = n0 + tid; i < n; i += 128) s_list[i] = a.acc_idx[acc0 + (i - n0)];
    __syncthreads();
    patch_sort_store(s_list, n, a.idx + a.off[row], tid);
}

// exclusive scan of (counts + accepted) over the shard rows; also total and max.  Single CTA.
__global__ void __launch_bounds__(1024) k_patch_scan(int n, const int* __restrict__ counts, const int* __restrict__ acc_off,
                                                     int* __restrict__ off, int* __restrict__ total_max) {
    // each thread owns one contiguous segment: local sum, one block-wide scan of the 1024 partials, local rewrite
    __shared__ int s[1024];
    __shared__ int s_max;
    const int tid = threadIdx.x;
    const int seg = (n + 1023) / 1024;
    const int b = min(n, tid * seg), e = min(n, b + seg);
    if (tid == 0) s_max = 0;
    int sum = 0, mx = 0;
    for (int i = b; i < e; i++) {
        const int c = counts[i] + (acc_off ? (acc_off[i + 1] - acc_off[i]) : 0);
        sum += c; mx = max(mx, c);
    }
    s[tid] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = (tid >= o) ? s[tid - o] : 0;
        __syncthreads();
        s[tid] += v;
        __syncthreads();
    }
    atomicMax(&s_max, mx);
    int run = s[tid] - sum;
    for (int i = b; i < e; i++) {
        const int c = counts[i] + (acc_off ? (acc_off[i + 1] - acc_off[i]) : 0);
        off[i] = run; run += c;
    }
    __syncthreads();
    if (tid == 1023) { off[n] = s[1023]; total_max[0] = s[1023]; total_max[1] = s_max; }
}

// Two-pass exclusive scan for LARGE arrays (the voxel grid's cell counts, up to 64^3): per-chunk sums, then every chunk
// scans itself on top of the sum of the chunks before it.  Coalesced reads, one CTA per chunk.
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__global__ void __launch_bounds__(256) k_scan_chunk_sums(int n, int chunk, const int* __restrict__ counts, int* __restrict__ sums) {
    __shared__ int s_w[8];
    const int b = blockIdx.x * chunk, e = min(n, b + chunk);
    int sum = 0;
    for (int i = b + threadIdx.x; i < e; i += 256) sum += counts[i];
    sum = warp_sum_i(sum);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; w++) t += s_w[w];
        sums[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(256) k_scan_chunk_apply(int n, int chunk, const int* __restrict__ counts, const int* __restrict__ sums,
                                                          int* __restrict__ off) {
    __shared__ int s_w[8];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) {
        int acc = 0;
        for (int j = lane; j < (int)blockIdx.x; j += 32) acc += sums[j];
        acc = warp_sum_i(acc);
        if (lane == 0) s_base = acc;
    }
    __syncthreads();
    int run = s_base;
    const int b = blockIdx.x * chunk, e = min(n, b + chunk);
    for (int tile = b; tile < e; tile += 256) {
        const int i = tile + tid;
        const int c = i < e ? counts[i] : 0;
        int inc = c;   // inclusive scan within the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (lane == 31) s_w[warp] = inc;
        __syncthreads();
        int before = 0, all = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) { const int t = s_w[w]; if (w < warp) before += t; all += t; }
        if (i < e) off[i] = run + before + inc - c;
        run += all;
        __syncthreads();
    }
    if (blockIdx.x == gridDim.x - 1 && tid == 0) off[n] = run;
}

// gather the packed per-patch sample streams (SoA over the global sample index) — coalesced reads in k_unary.
// Coordinates are stored as offsets from the patch's control point, r = SRC_i - CP_k, formed in fp64: small numbers
// keep full relative precision in fp32 (absolute coordinates at |x|~100 would not resolve an ico7 triangle to 1e-5).
// Univariate: source values are stored relative to the patch's first sample (Pearson is shift invariant), so the
// fp32 copy resolves the variation inside the patch rather than the field's absolute level.
// `vkey` (optional): a spatial key per source vertex (its leaf of the target hierarchy, a nested = space-filling
// numbering).  The samples of a patch are then PACKED in key order, so that the 32 samples a warp of k_unary works on are
// neighbours on the sphere: their descents read the same few table lines (the L1 wavefronts of those 16-byte gathers are
// k_unary's top limiter) and take the same branches.  The public patch list (idx, ascending ids) is not reordered.
__global__ void __launch_bounds__(128) k_patch_pack(int k_begin, int total, int Cp, int C, int wrows, int V, int univariate,
                                                    const int* __restrict__ off, const int* __restrict__ idx, const int* __restrict__ vkey,
                                                    const double* __restrict__ src, const double* __restrict__ cp,
                                                    const double* __restrict__ feat /*[C][V]*/, const float* __restrict__ wgt /*[wrows][V]*/,
                                                    float* __restrict__ px, float* __restrict__ py, float* __restrict__ pz,
                                                    float* __restrict__ pa /*[Cp][total]*/, float* __restrict__ pw) {
    extern __shared__ int s_key[];
    const int row = blockIdx.x, k = k_begin + row;
    const double ox = cp[3 * k], oy = cp[3 * k + 1], oz = cp[3 * k + 2];
    const int s_begin = off[row], s_end = off[row + 1], P = s_end - s_begin;
    const double aref = (univariate && s_end > s_begin) ? feat[idx[s_begin]] : 0.0;
    if (vkey) {
        for (int j = threadIdx.x; j < P; j += blockDim.x) s_key[j] = vkey[idx[s_begin + j]];
        __syncthreads();
    }
    for (int j = threadIdx.x; j < P; j += blockDim.x) {
        const int i = idx[s_begin + j];
        int pos = j;
        if (vkey) {   // rank of (key, j) among the patch's samples
            const int kj = s_key[j];
            pos = 0;
            for (int t = 0; t < P; t++) {
                const int kt = s_key[t];
                pos += (kt < kj || (kt == kj && t < j)) ? 1 : 0;
            }
        }
        const int s = s_begin + pos;
        px[s] = (float)(src[3 * i] - ox); py[s] = (float)(src[3 * i + 1] - oy); pz[s] = (float)(src[3 * i + 2] - oz);
        for (int c = 0; c < Cp; c++) {
            float a = 0.f, w = 0.f;  // padded channels carry zero weight: they drop out of every weighted sum
            if (c < C) {
                a = (float)(feat[(size_t)c * V + i] - aref);
                w = (c < wrows) ? wgt[(size_t)c * V + i] : 1.0f;  // src/DiscreteCostFunction.cpp:404-407,474-477
            }
            pa[(size_t)c * total + s] = a;
            pw[(size_t)c * total + s] = w;
        }
    }
}

// face-packed target features: leaf -> (f[a][0..Cp), f[b][..], f[c][..]).
// Univariate: ONE float4 per face, (f_a hi, f_a lo, f_b - f_a, f_c - f_a) formed in fp64: the interpolated value
// f_a + w_b (f_b-f_a) + w_c (f_c-f_a) then keeps ~1e-15 relative accuracy of the input although stored as 4 floats.
__global__ void __launch_bounds__(256) k_face_pack(int F, int C, int Cp, int V, int univariate, const int* __restrict__ leaf_vid /*[F][3]*/,
                                                   const double* __restrict__ feat /*[C][V]*/, float* __restrict__ out) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    int v0 = leaf_vid[3 * f], v1 = leaf_vid[3 * f + 1], v2 = leaf_vid[3 * f + 2];
    if (univariate) {
        const double f0 = feat[v0], f1 = feat[v1], f2 = feat[v2];
        const float hi = (float)f0;
        float4 o = make_float4(hi, (float)(f0 - (double)hi), (float)(f1 - f0), (float)(f2 - f0));
        reinterpret_cast<float4*>(out)[f] = o;
    } else {
        float* o = out + (size_t)f * 3 * Cp;
        for (int c = 0; c < Cp; c++) {
            o[c] = c < C ? (float)feat[(size_t)c * V + v0] : 0.f;
            o[Cp + c] = c < C ? (float)feat[(size_t)c * V + v1] : 0.f;
            o[2 * Cp + c] = c < C ? (float)feat[(size_t)c * V + v2] : 0.f;
        }
    }
}

// =========================================================================================================
// Triplet regulariser — computeTripletCost (src/DiscreteCostFunction.cpp:138-254), strain energy
// (src/reg_tools.cpp:556-601,703-748, Gram form: SURVEY Appendix A), group variant
// (src/DiscreteGroupCostFunction.cpp:9-30).  fp64: the cost is a difference of near-equal quantities for small
// deformations and coordinates live at |x|~100.
// =========================================================================================================
struct TripGeom {            // per-triplet constants
    double n[3];             // (v1-v0)x(v2-v0) of the CURRENT control triangle (fold test)
    double G11, G12, G22;    // Gram of the ORIGINAL triangle
    double det0, idet0;
    double oang[3], oarea;   // rmode 2
};

__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

__device__ __forceinline__ void tri_angles(const double* v0, const double* v1, const double* v2, double* ang) {
    const double* v[3] = {v0, v1, v2};
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double e1[3], e2[3];
        for (int j = 0; j < 3; j++) { e1[j] = v[(i + 1) % 3][j] - v[i][j]; e2[j] = v[(i + 2) % 3][j] - v[i][j]; }
        double c = dot3(e1, e2) / (sqrt(dot3(e1, e1)) * sqrt(dot3(e2, e2)));
        c = fmin(1.0, fmax(-1.0, c));
        ang[i] = acos(c);
    }
}

__device__ __forceinline__ void trip_geom(const DevParams& p, const double* v0, const double* v1, const double* v2, const double* o0,
                                          const double* o1, const double* o2, TripGeom& g) {
    double e1[3], e2[3];
    for (int j = 0; j < 3; j++) { e1[j] = v1[j] - v0[j]; e2[j] = v2[j] - v0[j]; }
    cross3(e1, e2, g.n);
    double E1[3], E2[3];
    for (int j = 0; j < 3; j++) { E1[j] = o1[j] - o0[j]; E2[j] = o2[j] - o0[j]; }
    g.G11 = dot3(E1, E1); g.G12 = dot3(E1, E2); g.G22 = dot3(E2, E2);
    g.det0 = g.G11 * g.G22 - g.G12 * g.G12;
    g.idet0 = 1.0 / g.det0;
    if (p.rmode == 2 && !p.group) {
        tri_angles(o0, o1, o2, g.oang);
        double c[3]; cross3(E1, E2, c);
        g.oarea = 0.5 * sqrt(dot3(c, c));
    }
}

// strain energy from Gram entries; k_exp == 2 avoids sqrt/pow: R^2+R^-2-2 = I1*^2-4, J^2 = I3
__device__ __forceinline__ double strain_energy(double g11, double g12, double g22, const TripGeom& g, double mu, double kappa, double k_exp) {
    const double trC = (g11 * g.G22 - 2.0 * g12 * g.G12 + g22 * g.G11) * g.idet0;   // one reciprocal per triplet, not
    const double I3 = (g11 * g22 - g12 * g12) * g.idet0;                            // two divisions per evaluation
    if (k_exp == 2.0) {
        const double r = 1.0 / I3;
        const double i1s2 = trC * trC * r;  // (I1*)^2
        // reference: R = 1 if I1* <= 2 (J = sqrt(I3) > 0 so the sign of I1* is the sign of trC > 0)
        const double shear = (i1s2 <= 4.0) ? 0.0 : (i1s2 - 4.0) + 0.0;
        // R^2 + R^-2 - 2 with R = (I1* + sqrt(I1*^2-4))/2 :  R + 1/R = I1*  =>  R^2 + R^-2 = I1*^2 - 2
        return 0.5 * (mu * shear + kappa * (I3 + r - 2.0));
    }
    const double J = sqrt(I3);
    const double I1s = trC / J;
    const double R = (I1s <= 2.0) ? 1.0 : 0.5 * (I1s + sqrt(I1s * I1s - 4.0));
    const double Rs = pow(R, k_exp), Js = pow(J, k_exp);
    return 0.5 * (mu * (Rs + 1.0 / Rs - 2.0) + kappa * (Js + 1.0 / Js - 2.0));
}

__device__ __forceinline__ double trip_cost(const DevParams& p, const TripGeom& g, const double* d0, const double* d1, const double* d2) {
    double e1[3], e2[3], nn[3];
    for (int j = 0; j < 3; j++) { e1[j] = d1[j] - d0[j]; e2[j] = d2[j] - d0[j]; }
    cross3(e1, e2, nn);
    const bool fold = dot3(nn, g.n) < 0.0;  // sign of the dot of the unit normals (:163)
    if (p.group) {                           // src/DiscreteGroupCostFunction.cpp:23,29
        if (fold) return 1e7;
        const double W = strain_energy(dot3(e1, e1), dot3(e1, e2), dot3(e2, e2), g, p.mu, p.kappa, 2.0);
        return p.lambda * pow(W, p.rexp);
    }
    double cost, weight = 1.0;
    if (fold) { cost = 1.0; weight += 1e6; }
    else if (p.rmode == 3) cost = strain_energy(dot3(e1, e1), dot3(e1, e2), dot3(e2, e2), g, p.mu, p.kappa, p.k_exp);
    else {  // rmode 2 (:172-192)
        double ang[3];
        tri_angles(d0, d1, d2, ang);
        const double area = 0.5 * sqrt(dot3(nn, nn));
        const double distortion = log2(area / g.oarea);
        double tweight = 1.0;
        if (p.dweight && distortion != 1.0) tweight = exp(fabs(distortion - 1.0));
        double diff = 0.0;
        for (int i = 0; i < 3; i++) { const double dd = ang[i] - g.oang[i]; diff += dd * dd; }
        cost = tweight * sqrt(diff);
        if (p.anorm) weight *= (1.0 / p.meanangle);
    }
    if (fabs(cost) < 1e-8) cost = 0.0;                                      // :246
    const double pw = (p.rexp == 2.0) ? cost * cost : pow(cost, p.rexp);     // :253
    return weight * p.lambda * pw;
}

struct TripArgs {
    DevParams p;
    int L;
    int group_N, group_T;            // per-subject sizes (group mode)
    int t_off;                       // first triplet of a sharded moves launch
    const int* trip;                 // [T][3]
    const double* cp;                // [N][3] current
    const double* orig;              // [N or N_subj][3]
    const double* disp;              // [N][L][3]
};

__device__ __forceinline__ void trip_load(const TripArgs& a, int t, int id[3], TripGeom& g) {
    id[0] = a.trip[3 * t]; id[1] = a.trip[3 * t + 1]; id[2] = a.trip[3 * t + 2];
    int m[3] = {id[0], id[1], id[2]};
    if (a.p.group) {
        const int mesh = t / a.group_T;
        for (int j = 0; j < 3; j++) m[j] = id[j] - mesh * a.group_N;
    }
    trip_geom(a.p, a.cp + 3 * id[0], a.cp + 3 * id[1], a.cp + 3 * id[2], a.orig + 3 * m[0], a.orig + 3 * m[1], a.orig + 3 * m[2], g);
}

// arbitrary queries (t, la, lb, lc)
__global__ void __launch_bounds__(128) k_triplet_queries(TripArgs a, int n, const int* __restrict__ q, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = q[4 * i];
    int id[3]; TripGeom g;
    trip_load(a, t, id, g);
    out[i] = trip_cost(a.p, g, a.disp + ((size_t)id[0] * a.L + q[4 * i + 1]) * 3, a.disp + ((size_t)id[1] * a.L + q[4 * i + 2]) * 3,
                       a.disp + ((size_t)id[2] * a.L + q[4 * i + 3]) * 3);
}

// fusion-move costs (include/Fusion/Fusion.h:187-194): combo bit2->A, bit1->B, bit0->C take label alpha.
// alpha >= 0: out[t][8]; alpha < 0: all labels, out[alpha][t][8].
// The default configuration (strain regulariser with k = 2 and exponent 2, or the group rules with exponent 2) without the
// angular mode, pow() and the other rarely used branches: a fraction of the code and registers of trip_cost.  Same
// arithmetic as trip_cost for those parameters.
__device__ __forceinline__ double trip_cost_fast(const DevParams& p, const TripGeom& g, const double* d0, const double* d1, const double* d2) {
    double e1[3], e2[3], nn[3];
    for (int j = 0; j < 3; j++) { e1[j] = d1[j] - d0[j]; e2[j] = d2[j] - d0[j]; }
    cross3(e1, e2, nn);
    const bool fold = dot3(nn, g.n) < 0.0;
    const double W = strain_energy(dot3(e1, e1), dot3(e1, e2), dot3(e2, e2), g, p.mu, p.kappa, 2.0);
    if (p.group) return fold ? 1e7 : p.lambda * (W * W);
    double cost = fold ? 1.0 : W;
    const double weight = fold ? 1.0 + 1e6 : 1.0;
    if (fabs(cost) < 1e-8) cost = 0.0;
    return weight * p.lambda * (cost * cost);
}
__host__ __device__ inline bool trip_fast_ok(const DevParams& p) {
    return p.rexp == 2.0 && (p.group || (p.rmode == 3 && p.k_exp == 2.0)) && p.lambda >= 0.0 && p.mu >= 0.0 && p.kappa >= 0.0;
}

template <typename OutT, bool FAST>
__global__ void __launch_bounds__(128) k_triplet_moves(TripArgs a, int T, const int* __restrict__ labeling, int alpha, OutT* __restrict__ out,
                                                       PeerSet peers = PeerSet{}, int t_full = 0) {
    // one thread per (alpha, triplet): the triangle constants and the six candidate corners (current / alpha label of
    // each vertex) are formed once and shared by the 8 combinations
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int nalpha = alpha < 0 ? a.L : 1;
    if (gid >= (long long)nalpha * T) return;
    const int t = a.t_off + (int)(gid % T);   // T = number of triplets of this launch (a shard [t_off, t_off+T))
    const int al = alpha < 0 ? (int)(gid / T) : alpha;
    int id[3]; TripGeom g;
    trip_load(a, t, id, g);
    double pc[3][3], pa[3][3];
#pragma unroll
    for (int v = 0; v < 3; v++) {
        const double* c = a.disp + ((size_t)id[v] * a.L + labeling[id[v]]) * 3;
        const double* n = a.disp + ((size_t)id[v] * a.L + al) * 3;
#pragma unroll
        for (int j = 0; j < 3; j++) { pc[v][j] = c[j]; pa[v][j] = n[j]; }
    }
    float4 lo, hi;   // the 8 costs, also as two 16-byte stores for the fused exchange
    OutT* o = peers.world > 0 ? nullptr : out + gid * 8;
#pragma unroll
    for (int combo = 0; combo < 8; combo++) {
        const double* d0 = (combo & 4) ? pa[0] : pc[0];
        const double* d1 = (combo & 2) ? pa[1] : pc[1];
        const double* d2 = (combo & 1) ? pa[2] : pc[2];
        const OutT v = (OutT)(FAST ? trip_cost_fast(a.p, g, d0, d1, d2) : trip_cost(a.p, g, d0, d1, d2));
        if (o) o[combo] = v;
        (&(combo < 4 ? lo : hi).x)[combo & 3] = (float)v;
    }
    if (peers.world > 0) {
        // fused exchange (triplet sharding): straight into the full [L or 1][t_full][8] table of every rank
        const size_t at = ((size_t)(alpha < 0 ? al : 0) * t_full + t) * 2;
        if (peers.mc) {
            float4* dst = reinterpret_cast<float4*>(peers.mc) + at;
            multicast_store(dst, lo); multicast_store(dst + 1, hi);
        } else {
            for (int r = 0; r < peers.world; r++) {
                float4* dst = reinterpret_cast<float4*>(peers.p[r]) + at;
                dst[0] = lo; dst[1] = hi;
            }
        }
    }
}

// full table: one CTA per triplet; displaced corners staged in shared memory; thread = (la,lb), loop lc;
// results staged in shared memory and written with coalesced stores (the table is store-bound: 4 B/eval).
__global__ void __launch_bounds__(256) k_triplet_table(TripArgs a, int t0, float* __restrict__ out) {
    extern __shared__ double sm_d[];
    const int L = a.L, L2 = L * L, L3 = L2 * L;
    double* sd = sm_d;                                   // [3][L][3]
    float* sres = reinterpret_cast<float*>(sd + 9 * L);  // [L3]
    __shared__ TripGeom g;
    __shared__ int id[3];
    const int t = t0 + blockIdx.x;
    if (threadIdx.x == 0) trip_load(a, t, id, g);
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * L; i += blockDim.x) {
        int v = i / (3 * L), r = i - v * 3 * L;
        sd[i] = a.disp[(size_t)id[v] * L * 3 + r];
    }
    __syncthreads();
    const bool fast = (a.p.group || (a.p.rmode == 3 && a.p.k_exp == 2.0));   // strain energy with k = 2: no sqrt / pow
    const double inv0 = 1.0 / g.det0;
    for (int ab = threadIdx.x; ab < L2; ab += blockDim.x) {
        const int la = ab / L, lb = ab - la * L;
        const double* d0 = sd + 3 * la;
        const double* d1 = sd + 3 * L + 3 * lb;
        if (!fast) {
            for (int lc = 0; lc < L; lc++) sres[ab * L + lc] = (float)trip_cost(a.p, g, d0, d1, sd + 6 * L + 3 * lc);
            continue;
        }
        // everything that depends on (la, lb) only is hoisted out of the lc loop
        const double e1[3] = {d1[0] - d0[0], d1[1] - d0[1], d1[2] - d0[2]};
        const double g11 = dot3(e1, e1);
        double c[3];
        cross3(g.n, e1, c);                       // (e1 x e2).n = e2.(n x e1)
        const double t11 = g11 * g.G22 * inv0, t12 = -2.0 * g.G12 * inv0, t22 = g.G11 * inv0;
        for (int lc = 0; lc < L; lc++) {
            const double* d2 = sd + 6 * L + 3 * lc;
            const double e2[3] = {d2[0] - d0[0], d2[1] - d0[1], d2[2] - d0[2]};
            const double g12 = dot3(e1, e2), g22 = dot3(e2, e2);
            const bool fold = dot3(e2, c) < 0.0;
            const double trC = t11 + t12 * g12 + t22 * g22;
            const double I3 = (g11 * g22 - g12 * g12) * inv0;
            const double r = 1.0 / I3;
            const double i1s2 = trC * trC * r;
            double W = 0.5 * (a.p.mu * fmax(i1s2 - 4.0, 0.0) + a.p.kappa * (I3 + r - 2.0));
            double res;
            if (a.p.group) res = fold ? 1e7 : a.p.lambda * ((a.p.rexp == 2.0) ? W * W : pow(W, a.p.rexp));
            else {
                double weight = 1.0;
                if (fold) { W = 1.0; weight += 1e6; }
                if (fabs(W) < 1e-8) W = 0.0;
                res = weight * a.p.lambda * ((a.p.rexp == 2.0) ? W * W : pow(W, a.p.rexp));
            }
            sres[ab * L + lc] = (float)res;
        }
    }
    __syncthreads();
    float* o = out + (size_t)blockIdx.x * L3;
    for (int i = threadIdx.x; i < L3; i += blockDim.x) o[i] = sres[i];
}

// =========================================================================================================
// Pairwise rotation regulariser — computePairwiseCost (src/DiscreteCostFunction.cpp:256-296), table layout :303.
// =========================================================================================================
struct PairArgs {
    DevParams p;
    int L, P;
    const int* pairs;      // [P][2]
    const double* Kd;      // [N][L][9]  R - I
    const double* disp;    // [N][L][3]
    const double* cp;      // current CP grid
    const double* ocp;     // _oCPgrid
    const int* tris;       // [T][3] CP-grid faces
    const int* vt_off; const int* vt_idx;  // vertex -> incident faces (ascending)
};

__device__ __forceinline__ double pair_cost(const PairArgs& a, int pr, int la, int lb) {
    const int va = a.pairs[2 * pr], vb = a.pairs[2 * pr + 1];
    const double* K1 = a.Kd + ((size_t)va * a.L + la) * 9;
    const double* K2 = a.Kd + ((size_t)vb * a.L + lb) * 9;
    double fro = 0.0;
#pragma unroll
    for (int i = 0; i < 9; i++) fro += K1[i] * K2[i];
    // (trace(R1^T R2) - 1)/2 = 1 + (tr K1 + tr K2 + K1:K2)/2
    const double x = -0.5 * ((K1[0] + K1[4] + K1[8]) + (K2[0] + K2[4] + K2[8]) + fro);  // = 1 - (tr-1)/2
    if (!(fabs(x) > kEPSILON)) return 0.0;                                            // :272
    const double theta = (x < 0.5) ? 2.0 * asin(sqrt(fmax(x, 0.0) * 0.5)) : acos(1.0 - x);
    const double theta_MVD = 2.0 * asin(a.p.mvdmax / (2.0 * kRAD));                   // :262
    double weight = 1.0;
    const double* da = a.disp + ((size_t)va * a.L + la) * 3;
    const double* db = a.disp + ((size_t)vb * a.L + lb) * 3;
    for (int e = a.vt_off[va]; e < a.vt_off[va + 1]; e++) {  // :274-282 faces incident to the FIRST vertex
        const int f = a.vt_idx[e];
        const double* o[3]; const double* c[3];
        for (int j = 0; j < 3; j++) {
            const int v = a.tris[3 * f + j];
            o[j] = a.ocp + 3 * v;
            c[j] = (v == va) ? da : (v == vb) ? db : a.cp + 3 * v;
        }
        double e1[3], e2[3], n0[3], n1[3];
        for (int j = 0; j < 3; j++) { e1[j] = o[1][j] - o[0][j]; e2[j] = o[2][j] - o[0][j]; }
        cross3(e1, e2, n0);
        for (int j = 0; j < 3; j++) { e1[j] = c[1][j] - c[0][j]; e2[j] = c[2][j] - c[0][j]; }
        cross3(e1, e2, n1);
        if (dot3(n0, n1) < 0.0) { weight += 1e6; break; }
    }
    const double cost = (1.4142135623730951 * theta) / theta_MVD;  // :284
    if (a.p.rexp == 1.0) return weight * a.p.lambda * cost;        // :286-289
    return weight * a.p.lambda * pow(cost, a.p.rexp);
}

template <typename OutT>
__global__ void __launch_bounds__(128) k_pair_table(PairArgs a, OutT* __restrict__ out) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int L2 = a.L * a.L;
    if (gid >= (long long)a.P * L2) return;
    const int pr = (int)(gid / L2);
    const int r = (int)(gid - (long long)pr * L2);
    const int lb = r / a.L, la = r - lb * a.L;  // paircosts[p*L*L + labelB*L + labelA]
    out[gid] = (OutT)pair_cost(a, pr, la, lb);
}

__global__ void __launch_bounds__(128) k_pair_queries(PairArgs a, int n, const int* __restrict__ q, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = pair_cost(a, q[3 * i], q[3 * i + 1], q[3 * i + 2]);
}

// =========================================================================================================
// Group data term — DiscreteGroupCostFunction::computePairwiseCost (src/DiscreteGroupCostFunction.cpp:32-54):
// Pearson correlation over the intersection of two (template vertex -> value) patch maps, in A's key order, unweighted,
// returned as 1-(1+rho)/2 (no lambda).  Maps are CSR with ascending keys (std::map order), so the intersection is a
// merge join.  One thread per query; fp64; two passes like the reference (means, then centred sums).
// =========================================================================================================
struct GroupPairArgs {
    int L, N_subj;
    const int* pairs;            // [P][2] global node ids
    const long long* moff;       // [nmaps+1]
    const int* mkey;             // ascending within a map
    const double* mval;
};

__device__ __forceinline__ double group_pair_cost(const GroupPairArgs& a, int pr, int la, int lb) {
    const int ga = a.pairs[2 * pr], gb = a.pairs[2 * pr + 1];
    // patch_data index = subject*N*L + node*L + label = global_node*L + label
    const long long a0 = a.moff[(long long)ga * a.L + la], a1 = a.moff[(long long)ga * a.L + la + 1];
    const long long b0 = a.moff[(long long)gb * a.L + lb], b1 = a.moff[(long long)gb * a.L + lb + 1];
    double sA = 0, sB = 0; int n = 0;
    for (long long i = a0, j = b0; i < a1 && j < b1;) {
        const int ka = a.mkey[i], kb = a.mkey[j];
        if (ka == kb) { sA += a.mval[i]; sB += a.mval[j]; n++; i++; j++; }
        else if (ka < kb) i++;
        else j++;
    }
    if (n > 0) { sA /= n; sB /= n; }           // src/similarities.cc:139-142 (sum of unit weights = n)
    double pr_ = 0, vA = 0, vB = 0;
    for (long long i = a0, j = b0; i < a1 && j < b1;) {
        const int ka = a.mkey[i], kb = a.mkey[j];
        if (ka == kb) {
            const double da = a.mval[i] - sA, db = a.mval[j] - sB;
            pr_ += da * db; vA += da * da; vB += db * db; i++; j++;
        } else if (ka < kb) i++;
        else j++;
    }
    if (n > 0) { pr_ /= n; vA /= n; vB /= n; }
    const double rho = (vA == 0.0 || vB == 0.0) ? 0.0 : pr_ / (sqrt(vA) * sqrt(vB));
    return 1.0 - (1.0 + rho) / 2.0;
}

__global__ void __launch_bounds__(128) k_group_pair_queries(GroupPairArgs a, int n, const int* __restrict__ q, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = group_pair_cost(a, q[3 * i], q[3 * i + 1], q[3 * i + 2]);
}

// the whole [P][L][L] table in the reference's pairwise layout (pair*L*L + labelB*L + labelA): serves the per-element
// virtual of the unmodified Fusion::optimize as a look-up
__global__ void __launch_bounds__(128) k_group_pair_table(GroupPairArgs a, int P, float* __restrict__ out) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int L2 = a.L * a.L;
    if (gid >= (long long)P * L2) return;
    const int pr = (int)(gid / L2), r = (int)(gid - (long long)pr * L2);
    const int lb = r / a.L, la = r - lb * a.L;
    out[gid] = (float)group_pair_cost(a, pr, la, lb);
}

// fusion-move costs of include/Fusion/Fusion.h:169-172: (cur,cur) (cur,alpha) (alpha,cur) (alpha,alpha) per pair
__global__ void __launch_bounds__(128) k_group_pair_moves(GroupPairArgs a, int P, const int* __restrict__ labeling, int alpha,
                                                          double* __restrict__ out) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)P * 4) return;
    const int pr = (int)(gid >> 2), c = (int)(gid & 3);
    const int la = (c & 2) ? alpha : labeling[a.pairs[2 * pr]];
    const int lb = (c & 1) ? alpha : labeling[a.pairs[2 * pr + 1]];
    out[gid] = group_pair_cost(a, pr, la, lb);
}

// =========================================================================================================
// k_locate — stand-alone point location + barycentric weights (msm_locate), fp64.
// =========================================================================================================
__global__ void __launch_bounds__(256) k_locate(IcoDev ico, int M, const double* __restrict__ pts, const int* __restrict__ leaf_user /*[F]*/,
                                                const unsigned char* __restrict__ leaf_perm /*[F][3]*/, int* __restrict__ tri,
                                                double* __restrict__ w) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    double al, be, ga;
    const double px = pts[3 * i], py = pts[3 * i + 1], pz = pts[3 * i + 2];
    const int hint = ico_base_face<double>(ico, px, py, pz);
    const int leaf = ico_locate<double>(ico, hint, px, py, pz, al, be, ga);
    const double inv = 1.0 / (al + be + ga);
    const double wl[3] = {al * inv, be * inv, ga * inv};
    tri[i] = leaf_user[leaf];
    // leaf corner j sits at position leaf_perm[leaf][j] of the caller's triangle
    w[3 * i + leaf_perm[3 * leaf + 0]] = wl[0];
    w[3 * i + leaf_perm[3 * leaf + 1]] = wl[1];
    w[3 * i + leaf_perm[3 * leaf + 2]] = wl[2];
}

}  // namespace msm

#include "unary.cuh"
