// patchfast.cuh — one-pass patch construction for the second and later builds of a level
// (get_source_data / within_controlpt_range, src/DiscreteCostFunction.cpp:104-109,394-410,458-481).
//
// Patches change little between outer iterations, so the largest patch of the previous build bounds this one's: every
// control point gets a fixed-capacity row, and ONE kernel (a warp per control point) searches the voxel grid, ranks the
// members it found by (spatial key, vertex id) — the order k_unary consumes, the same one k_patch_pack produces — and
// writes the packed sample streams.  No offsets scan, no sort of the public (ascending) list, no second search, and the
// host reads ONE flag word at the end instead of sizing buffers in the middle of the pipeline.  Anything unusual — a pair
// inside the 1e-11 uncertainty band of the range test (the exact ties of an undeformed icosphere), a patch that outgrew
// its row — sends the build through the exact two-pass path of kernels.cuh instead; membership decisions are the same
// fp64 tests either way, so both paths give bit-identical patches and tables (tests/test_gpu_parity.py).
#pragma once
#include "kernels.cuh"

namespace msm {

struct PatchFastArgs {
    int k_begin, Ns, V, cap;                 // cap: row capacity (samples per control point)
    BinGrid g;
    const float* sxf; const float* syf; const float* szf;   // source xyz fp32 SoA (pre-filter)
    const double* src;                       // source xyz fp64 [V][3]
    const double* cp;                        // CP xyz fp64 [N][3]
    const double* maxsep;                    // [N]
    double range;
    const int* vkey;                         // [V] spatial key of every source vertex (k_vertex_key)
    int Cp, C, wrows, univariate;
    const double* feat;                      // [C][V]
    const float* wgt;                        // [wrows][V]
    int* counts;                             // [Ns]
    int* off; int* end;                      // [Ns] row r of the packed streams: [r cap, r cap + count)
    int* idx;                                // [Ns][cap] member ids, packed order
    float* px; float* py; float* pz;         // [Ns][cap]
    float* pa; float* pw;                    // [Cp][Ns][cap]
    int* flags;                              // [0] uncertain pairs, [1] overflowing rows, [2] largest patch, [3] total
};

constexpr int kPatchWarps = 4;

template <int RMAX>   // 32 RMAX members are ranked per pass (RMAX = ceil(cap / 32) up to 4: one pass)
__global__ void __launch_bounds__(kPatchWarps * 32) k_patch_fused(PatchFastArgs a) {
    extern __shared__ int s_all[];           // per warp: ids[cap] | keys[cap]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * kPatchWarps + warp;
    if (row >= a.Ns) return;                 // whole warps leave together (no block barrier below)
    int* s_id = s_all + (size_t)warp * 2 * a.cap;
    int* s_key = s_id + a.cap;
    const int k = a.k_begin + row;
    const double cx = a.cp[3 * k], cy = a.cp[3 * k + 1], cz = a.cp[3 * k + 2];
    const float cxf = (float)cx, cyf = (float)cy, czf = (float)cz;
    double c2lo, c2hi; float c2f;
    patch_thresholds(a.range * a.maxsep[k], c2lo, c2hi, c2f);
    c2f += 1e-2f;                            // the fp32 pre-filter compares rounded coordinates (|x| <= ~100)
    const double reach = sqrt(c2hi) * (1.0 + 1e-9);
    const BinGrid& g = a.g;
    const int x0 = bin_coord(g, cx - reach), x1 = bin_coord(g, cx + reach);
    const int y0 = bin_coord(g, cy - reach), y1 = bin_coord(g, cy + reach);
    const int z0 = bin_coord(g, cz - reach), z1 = bin_coord(g, cz + reach);
    int n = 0, unc = 0;
    for (int x = x0; x <= x1; x++)
        for (int y = y0; y <= y1; y++) {
            // cells (x,y,z0..z1) are contiguous in memory: one candidate range per (x,y)
            const int b = g.cell_start[(x * g.G + y) * g.G + z0], e = g.cell_start[(x * g.G + y) * g.G + z1 + 1];
            for (int s0 = b; s0 < e; s0 += 32) {   // warp-uniform trip count
                const int s = s0 + lane;
                int i = -1, cls = 0;
                if (s < e) {
                    i = g.sorted[s];
                    const float dx = a.sxf[i] - cxf, dy = a.syf[i] - cyf, dz = a.szf[i] - czf;
                    if (dx * dx + dy * dy + dz * dz <= c2f) {
                        const double ex = cx - a.src[3 * i], ey = cy - a.src[3 * i + 1], ez = cz - a.src[3 * i + 2];
                        const double d2 = ex * ex + ey * ey + ez * ez;
                        cls = d2 < c2lo ? 1 : (d2 <= c2hi ? 2 : 0);
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, cls == 1);
                unc += __popc(__ballot_sync(0xffffffffu, cls == 2));
                if (cls == 1) {
                    const int p = n + __popc(m & ((1u << lane) - 1u));
                    if (p < a.cap) { s_id[p] = i; s_key[p] = a.vkey[i]; }
                }
                n += __popc(m);
            }
        }
    if (lane == 0) {
        if (unc) atomicAdd(a.flags + 0, unc);
        if (n > a.cap) atomicAdd(a.flags + 1, 1);
        atomicMax(a.flags + 2, n);
        atomicAdd(a.flags + 3, min(n, a.cap));
        a.counts[row] = min(n, a.cap);
        a.off[row] = row * a.cap; a.end[row] = row * a.cap + min(n, a.cap);
    }
    n = min(n, a.cap);
    __syncwarp();
    // rank by (key, id) and pack; coordinates as offsets from the control point, formed in fp64 (see k_patch_pack)
    const size_t base = (size_t)row * a.cap, plane = (size_t)a.Ns * a.cap;
    double aref = 0.0;
    if (a.univariate && n > 0) {
        // the patch's first sample in ASCENDING id order (what the two-pass path calls idx[s_begin]): Pearson is shift
        // invariant, the reference value only has to be the same on both paths
        int mn = 0x7fffffff;
        for (int j = lane; j < n; j += 32) mn = min(mn, s_id[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        aref = a.feat[mn];
    }
    // rank of every member under (key, id): a lane owns members lane, lane + 32, ... in registers and walks the list once,
    // one broadcast read per listed member
    for (int j0 = 0; j0 < n; j0 += 32 * RMAX) {
        int kj[RMAX], ij[RMAX], pos[RMAX];
#pragma unroll
        for (int r = 0; r < RMAX; r++) {
            const int j = j0 + lane + 32 * r;
            kj[r] = j < n ? s_key[j] : 0x7fffffff; ij[r] = j < n ? s_id[j] : 0x7fffffff; pos[r] = 0;
        }
        for (int t = 0; t < n; t++) {
            const int kt = s_key[t], it = s_id[t];
#pragma unroll
            for (int r = 0; r < RMAX; r++) pos[r] += (kt < kj[r] || (kt == kj[r] && it < ij[r])) ? 1 : 0;
        }
#pragma unroll
        for (int r = 0; r < RMAX; r++) {
            if (j0 + lane + 32 * r >= n) continue;
            const int id = ij[r];
            const size_t s = base + pos[r];
            a.idx[s] = id;
            a.px[s] = (float)(a.src[3 * id] - cx); a.py[s] = (float)(a.src[3 * id + 1] - cy); a.pz[s] = (float)(a.src[3 * id + 2] - cz);
            for (int c = 0; c < a.Cp; c++) {
                float av = 0.f, w = 0.f;   // padded channels carry zero weight: they drop out of every weighted sum
                if (c < a.C) {
                    av = (float)(a.feat[(size_t)c * a.V + id] - aref);
                    w = (c < a.wrows) ? a.wgt[(size_t)c * a.V + id] : 1.0f;   // src/DiscreteCostFunction.cpp:404-407,474-477
                }
                a.pa[(size_t)c * plane + s] = av;
                a.pw[(size_t)c * plane + s] = w;
            }
        }
    }
}

}  // namespace msm
