// features.cuh — feature pre-processing per resolution level: the device side of
// featurespace::initialize (src/featurespace.cc:18-65):
//   smooth_data            [EXT] A9, call site src/featurespace.cc:49          -> k_cell_rank_sort + k_smooth
//   histogram matching     src/reg_tools.cpp:750-810 over MISCMATHS::Histogram [EXT] A10 -> k_hist_stats + k_hist_match
//   variance_normalise     src/featurespace.cc:67-108                          -> k_varnorm
//   create_exclusion       [EXT] A11, call site src/featurespace.cc:42         -> k_exclusion
// (metric_resample, [EXT] A8, is host/src/mesh_warp.h over the point-location kernels.)  Everything is fp64: these run once
// per level on at most a few hundred thousand vertices; what matters is that they agree with the reference to rounding.
#pragma once
#include "kernels.cuh"

namespace msm {

// Cells of the voxel grid are filled with atomics (k_bin_fill): put every cell's vertex ids in ascending order so that the
// sums below accumulate in one fixed order, run to run.  A warp per cell, rank by counting.
__global__ void __launch_bounds__(128) k_cell_rank_sort(int ncells, const int* __restrict__ cell_start, const int* __restrict__ in, int* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (c >= ncells) return;
    const int b = cell_start[c], n = cell_start[c + 1] - b;
    for (int j = lane; j < n; j += 32) {
        const int v = in[b + j];
        int r = 0;
        for (int t = 0; t < n; t++) r += in[b + t] < v ? 1 : 0;
        out[b + r] = v;
    }
}

struct SmoothArgs {
    BinGrid g;                       // over the `in` vertices, cell size >= chord_max
    int Vin, Vref, D;
    const double* in_xyz;            // [Vin][3]
    const double* ref_xyz;           // [Vref][3]
    const double* values;            // [D][Vin]
    const double* mask;              // [Vin] or null: sources with mask <= 0 are skipped
    double chord_max, sigma;
    double* out;                     // [D][Vref]
};

// A warp per ref vertex; lanes stride the candidates of the <= 27 cells around it.  w = exp(-d^2 / 2 sigma^2) with the geodesic
// distance d = 2 RAD asin(chord / 2 RAD); DB feature dimensions per pass share the weights.
template <int DB>
__global__ void __launch_bounds__(128) k_smooth(SmoothArgs a) {
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (k >= a.Vref) return;         // whole warps leave together
    const double RAD = 100.0;
    const double cx = a.ref_xyz[3 * k], cy = a.ref_xyz[3 * k + 1], cz = a.ref_xyz[3 * k + 2];
    const BinGrid& g = a.g;
    const int x0 = bin_coord(g, cx - a.chord_max), x1 = bin_coord(g, cx + a.chord_max);
    const int y0 = bin_coord(g, cy - a.chord_max), y1 = bin_coord(g, cy + a.chord_max);
    const int z0 = bin_coord(g, cz - a.chord_max), z1 = bin_coord(g, cz + a.chord_max);
    const double inv2s2 = 1.0 / (2.0 * a.sigma * a.sigma);
    const double c2_reject = a.chord_max * a.chord_max * (1.0 + 1e-12);
    for (int d0 = 0; d0 < a.D; d0 += DB) {
        double W = 0.0, acc[DB];
#pragma unroll
        for (int q = 0; q < DB; q++) acc[q] = 0.0;
        for (int x = x0; x <= x1; x++)
            for (int y = y0; y <= y1; y++) {
                const int b = g.cell_start[(x * g.G + y) * g.G + z0], e = g.cell_start[(x * g.G + y) * g.G + z1 + 1];
                for (int s = b + lane; s < e; s += 32) {
                    const int j = g.sorted[s];
                    if (a.mask && !(a.mask[j] > 0.0)) continue;
                    const double ex = cx - a.in_xyz[3 * j], ey = cy - a.in_xyz[3 * j + 1], ez = cz - a.in_xyz[3 * j + 2];
                    const double d2 = ex * ex + ey * ey + ez * ez;
                    if (d2 > c2_reject) continue;                         // most candidates of the 27 cells: no sqrt / asin / exp
                    const double chord = sqrt(d2);
                    if (!(chord < a.chord_max)) continue;                 // the decision itself, as the reference's restatement makes it
                    const double d = 2.0 * RAD * asin(fmin(1.0, chord / (2.0 * RAD)));
                    const double w = exp(-(d * d) * inv2s2);
                    W += w;
#pragma unroll
                    for (int q = 0; q < DB; q++)
                        if (d0 + q < a.D) acc[q] += w * a.values[(size_t)(d0 + q) * a.Vin + j];
                }
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            W += __shfl_xor_sync(0xffffffffu, W, o);
#pragma unroll
            for (int q = 0; q < DB; q++) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < DB; q++)
                if (d0 + q < a.D) a.out[(size_t)(d0 + q) * a.Vref + k] = W > 0.0 ? acc[q] / W : 0.0;
        }
    }
}

// ---- MISCMATHS::Histogram (A10).  The float arithmetic of getBin is spelled with round-to-nearest intrinsics so that no
// contraction can change a bin.
__device__ __forceinline__ int hist_bin(double value, float hmin, float hmax, int bins) {
    const float v = (float)value;                                      // getBin(float value): the argument is converted first
    const float num = __fmul_rn((float)bins, __fsub_rn(v, hmin));
    const float den = __fsub_rn(hmax, hmin);
    const float q = __fdiv_rn(num, den);
    int b = (q == q && fabsf(q) < 2.0e9f) ? (int)q : (int)0x80000000;  // x86 cvttss2si of NaN / out of range
    return max(1, min(b, bins));
}

struct HistStats { float hmin, hmax, datapoints; int pad; };

// One CTA per (mesh, dimension): range over ALL values, counts of the non-excluded ones, CDF accumulated in index order.
// blockIdx.x = which * D + d (which: 0 = data to be matched, 1 = template)
__global__ void __launch_bounds__(1024) k_hist_stats(int D, int bins, const double* __restrict__ data0, int V0, const double* __restrict__ mask0, int rows0,
                                                     const double* __restrict__ data1, int V1, const double* __restrict__ mask1, int rows1,
                                                     HistStats* __restrict__ stats /*[2][D]*/, double* __restrict__ cdf /*[2][D][bins]*/) {
    extern __shared__ int s_hist[];            // [bins]
    __shared__ double s_lo[32], s_hi[32];
    __shared__ int s_excluded[32];
    const int which = blockIdx.x / D, d = blockIdx.x % D;
    const int V = which ? V1 : V0;
    const double* x = (which ? data1 : data0) + (size_t)d * V;
    const double* mk = which ? mask1 : mask0;
    const int rows = which ? rows1 : rows0;
    if (mk) mk += (size_t)(rows > d ? d : 0) * V;   // src/reg_tools.cpp:771-776: a mask row per dimension if there are that many
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double lo = x[0], hi = x[0];
    for (int i = tid; i < V; i += 1024) { const double v = x[i]; lo = fmin(lo, v); hi = fmax(hi, v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if (lane == 0) { s_lo[warp] = lo; s_hi[warp] = hi; }
    for (int i = tid; i < bins; i += 1024) s_hist[i] = 0;
    __syncthreads();
    lo = s_lo[0]; hi = s_hi[0];
    for (int w = 1; w < 32; w++) { lo = fmin(lo, s_lo[w]); hi = fmax(hi, s_hi[w]); }
    // the reference's running float min / max end at the rounded extremes whatever the order (rounding is monotone)
    const float hmin = (float)lo, hmax = (float)hi;
    int excluded = 0;
    for (int i = tid; i < V; i += 1024) {
        if (!mk || mk[i] > 0.0) atomicAdd(&s_hist[hist_bin(x[i], hmin, hmax, bins) - 1], 1);
        else excluded++;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) excluded += __shfl_xor_sync(0xffffffffu, excluded, o);
    if (lane == 0) s_excluded[warp] = excluded;
    __syncthreads();
    if (tid == 0) {
        int ex = 0;
        for (int w = 0; w < 32; w++) ex += s_excluded[w];
        const float datapoints = (float)V - (float)ex;     // float datapoints = size, decremented per excluded value (exact below 2^24)
        HistStats st; st.hmin = hmin; st.hmax = hmax; st.datapoints = datapoints; st.pad = 0;
        stats[which * D + d] = st;
        double* c = cdf + ((size_t)which * D + d) * bins;
        double run = (double)s_hist[0] / (double)datapoints;
        c[0] = run;
        for (int i = 1; i < bins; i++) { run = run + (double)s_hist[i] / (double)datapoints; c[i] = run; }
    }
}

// Histogram::match: a thread per (dimension, vertex) of the data to be matched
__global__ void __launch_bounds__(256) k_hist_match(int D, int bins, double* __restrict__ data, int V, const double* __restrict__ mask, int rows,
                                                    const HistStats* __restrict__ stats, const double* __restrict__ cdf) {
    extern __shared__ double s_cdf[];          // own [bins] | template [bins]
    const int d = blockIdx.y;
    for (int i = threadIdx.x; i < bins; i += blockDim.x) {
        s_cdf[i] = cdf[(size_t)d * bins + i];
        s_cdf[bins + i] = cdf[((size_t)D + d) * bins + i];
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const double* mk = mask ? mask + (size_t)(rows > d ? d : 0) * V : nullptr;
    if (mk && !(mk[i] > 0.0)) return;
    const HistStats own = stats[d], tp = stats[D + d];
    const int bin = hist_bin(data[(size_t)d * V + i], own.hmin, own.hmax, bins);
    const double cdfval = s_cdf[bin - 1];
    int newbin = 1;
    if (bin == bins) newbin = bin;
    else {
        // first template bin whose CDF reaches cdfval (the CDF is nondecreasing: the reference's linear walk finds the same bin)
        int lo = 1, hi = bins;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if ((cdfval - s_cdf[bins + mid - 1]) > 1e-20) lo = mid + 1; else hi = mid;
        }
        newbin = lo;
    }
    const float width = __fdiv_rn(__fsub_rn(tp.hmax, tp.hmin), (float)bins);
    double v = (double)__fadd_rn(tp.hmin, __fmul_rn((float)(newbin - 1), width));
    if (v < (double)tp.hmin) v = (double)tp.hmin;
    if (v > (double)tp.hmax) v = (double)tp.hmax;
    data[(size_t)d * V + i] = v;
}

// variance_normalise: one CTA per dimension; mean and (n-1) variance over the non-excluded vertices (two passes, fixed
// reduction order), then (x - mean) / sqrt(var), the division only when var > 0
__device__ __forceinline__ double block_sum_1024(double v, double* s_part) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) s_part[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < 32; w++) t += s_part[w];
    return t;
}
__global__ void __launch_bounds__(1024) k_varnorm(int V, double* __restrict__ data, const double* __restrict__ mask) {
    __shared__ double s_part[32];
    double* x = data + (size_t)blockIdx.x * V;
    const int tid = threadIdx.x;
    double n = 0.0, s = 0.0;
    for (int i = tid; i < V; i += 1024)
        if (!mask || mask[i] > 0.0) { n += 1.0; s += x[i]; }
    n = block_sum_1024(n, s_part);
    s = block_sum_1024(s, s_part);
    const double mean = s / n;
    double q = 0.0;
    for (int i = tid; i < V; i += 1024)
        if (!mask || mask[i] > 0.0) { const double e = x[i] - mean; q += e * e; }
    q = block_sum_1024(q, s_part);
    const double var = q / (n - 1.0);
    const double sd = sqrt(var);
    for (int i = tid; i < V; i += 1024)
        if (!mask || mask[i] > 0.0) {
            double v = x[i] - mean;
            if (var > 0.0) v /= sd;
            x[i] = v;
        }
}

// create_exclusion (A11): 0 where every dimension lies in [lower - 1e-8, upper + 1e-8]
__global__ void __launch_bounds__(256) k_exclusion(int D, int V, const double* __restrict__ data, double lower, double upper, double* __restrict__ mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    bool all_in = D > 0;
    for (int d = 0; d < D; d++) {
        const double v = data[(size_t)d * V + i];
        if (v < lower - 1e-8 || v > upper + 1e-8) { all_in = false; break; }
    }
    mask[i] = all_in ? 0.0 : 1.0;
}

}  // namespace msm
