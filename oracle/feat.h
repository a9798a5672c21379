// feat.h — ORACLE (test infrastructure): CPU restatement of the reference's feature pre-processing
// (SURVEY.md §8f rank 4): featurespace::initialize src/featurespace.cc:18-65, featurespace::variance_normalise
// src/featurespace.cc:67-108, multivariate_histogram_normalization src/reg_tools.cpp:750-810.
// Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it; nothing in the product path does.
//
// variance_normalise and the driver loop are the reference's own code and are pinned against it (oracle/_ref/libref_feat.so
// = src/featurespace.cc + src/reg_tools.cpp compiled where they lie, tests/test_reference_pin.py).  What they call into is
// third-party code absent from the tree ([EXT], FSL 6.0.x `fsl-newresampler` / `fsl-miscmaths`, no version pin in
// src/Makefile:5) and is restated here from its published behaviour — PARITY UNPINNED for exactly these:
//   A8  metric_resample(in, ref, threads, EXCL)  oracle/geom.h (adaptive barycentric); with an exclusion mask the sources whose
//       mask value is 0 drop out of every weight list, the rest is renormalised, a vertex left without sources gets 0.
//   A9  smooth_data(in, ref, sigma, threads, EXCL): Gaussian kernel of the geodesic distance d = 2 RAD asin(chord / 2 RAD),
//       w = exp(-d^2 / (2 sigma^2)), over the vertices of `in` whose angular separation from the ref vertex is below
//       2 asin(sigma / RAD) (support of about two sigma), excluded sources (mask 0) skipped, value = sum(w f) / sum(w), 0 when
//       nothing is in range.  Sources are visited in ascending vertex order.  Call site src/featurespace.cc:49.
//   A10 MISCMATHS::Histogram (miscmaths/histogram.{h,cc}): float range [min,max] of ALL values, `bins` equal bins,
//       getBin(v) = max(1, min(int(bins * float(v - min) / (max - min)), bins)) in float arithmetic, counts of the non-excluded
//       values, CDF(i) = CDF(i-1) + count(i) / datapoints, match(): every non-excluded value goes to the lower edge of the
//       first template bin whose CDF reaches the value's own CDF (last bin stays last), clamped to the template's range.
//       Call site src/reg_tools.cpp:792-806.
//   A11 create_exclusion(mesh, lower, upper): one mask row, 0 where EVERY data dimension of the vertex lies in
//       [lower - 1e-8, upper + 1e-8], else 1.  Call sites src/featurespace.cc:42, src/mesh_registration.cpp:241-242.
#pragma once
#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <vector>

#include "geom.h"

namespace orc {

// ---- A11
inline std::vector<double> create_exclusion(const std::vector<double>& data /*[D][V]*/, int D, int V, float lower, float upper) {
    std::vector<double> m(V, 1.0);
    for (int i = 0; i < V; i++) {
        bool all_in = D > 0;
        for (int d = 0; d < D; d++) {
            const double v = data[(size_t)d * V + i];
            if (v < (double)lower - 1e-8 || v > (double)upper + 1e-8) { all_in = false; break; }
        }
        if (all_in) m[i] = 0.0;
    }
    return m;
}

// ---- A8 with an exclusion mask (mask == nullptr: oracle/geom.h metric_resample)
inline std::vector<double> metric_resample_excl(const Mesh& in, const std::vector<double>& values, int dims, const Mesh& ref, const double* mask) {
    if (!mask) return metric_resample(in, values, dims, ref);
    const ResampleWeights R = adaptive_barycentric_weights(in, ref);
    const int Vin = in.nvertices(), Vref = ref.nvertices();
    std::vector<double> out((size_t)dims * Vref, 0.0);
    for (int k = 0; k < Vref; k++) {
        double sum = 0.0;
        for (int e = R.off[k]; e < R.off[k + 1]; e++) if (mask[R.src[e]] > 0.0) sum += R.w[e];
        if (sum == 0.0) continue;
        for (int d = 0; d < dims; d++) {
            double v = 0.0;
            for (int e = R.off[k]; e < R.off[k + 1]; e++) if (mask[R.src[e]] > 0.0) v += (R.w[e] / sum) * values[(size_t)d * Vin + R.src[e]];
            out[(size_t)d * Vref + k] = v;
        }
    }
    return out;
}

// ---- A9
inline std::vector<double> smooth_data(const std::vector<Pt>& in_xyz, const std::vector<double>& values /*[D][Vin]*/, int D,
                                       const std::vector<Pt>& ref_xyz, double sigma, const double* mask /*[Vin] or null*/) {
    const double RAD = 100.0;
    const int Vin = (int)in_xyz.size(), Vref = (int)ref_xyz.size();
    std::vector<double> out((size_t)D * Vref, 0.0);
    const double chord_max = 2.0 * RAD * std::sin(std::asin(std::min(1.0, sigma / RAD)));   // separation 2 asin(sigma/RAD)
    // uniform grid over the sources (the result does not depend on it: candidates are visited in ascending vertex order)
    double bound = 0.0;
    for (const Pt& p : in_xyz) bound = std::max({bound, std::fabs(p.X), std::fabs(p.Y), std::fabs(p.Z)});
    bound = bound * (1.0 + 1e-9) + 1e-6;
    const int G = std::max(1, std::min(64, (int)std::floor(2.0 * bound / std::max(chord_max, 1e-9))));
    const double inv_h = G / (2.0 * bound);
    auto cell = [&](double x) { int c = (int)std::floor((x + bound) * inv_h); return std::max(0, std::min(G - 1, c)); };
    std::vector<std::vector<int>> cells((size_t)G * G * G);
    for (int j = 0; j < Vin; j++) cells[((size_t)cell(in_xyz[j].X) * G + cell(in_xyz[j].Y)) * G + cell(in_xyz[j].Z)].push_back(j);
    std::vector<int> cand;
    std::vector<double> acc(D);
    for (int k = 0; k < Vref; k++) {
        const Pt& c = ref_xyz[k];
        cand.clear();
        for (int x = cell(c.X - chord_max); x <= cell(c.X + chord_max); x++)
            for (int y = cell(c.Y - chord_max); y <= cell(c.Y + chord_max); y++)
                for (int z = cell(c.Z - chord_max); z <= cell(c.Z + chord_max); z++)
                    for (int j : cells[((size_t)x * G + y) * G + z]) cand.push_back(j);
        std::sort(cand.begin(), cand.end());
        double W = 0.0;
        std::fill(acc.begin(), acc.end(), 0.0);
        for (int j : cand) {
            if (mask && !(mask[j] > 0.0)) continue;
            const double ex = c.X - in_xyz[j].X, ey = c.Y - in_xyz[j].Y, ez = c.Z - in_xyz[j].Z;
            const double chord = std::sqrt(ex * ex + ey * ey + ez * ez);
            if (!(chord < chord_max)) continue;
            const double d = 2.0 * RAD * std::asin(std::min(1.0, chord / (2.0 * RAD)));
            const double w = std::exp(-(d * d) / (2.0 * sigma * sigma));
            W += w;
            for (int q = 0; q < D; q++) acc[q] += w * values[(size_t)q * Vin + j];
        }
        if (W > 0.0) for (int q = 0; q < D; q++) out[(size_t)q * Vref + k] = acc[q] / W;
    }
    return out;
}

// ---- A10
struct Histogram {
    std::vector<double> source, hist, excl, cdf;
    float datapoints = 0.f, hmin = 0.f, hmax = 0.f;
    int bins = 0;
    Histogram(const std::vector<double>& data, int nbins) : source(data), bins(nbins) {}
    int get_bin(float value) const {
        return std::max(1, std::min((int)((((float)bins) * ((float)(value - hmin))) / ((float)(hmax - hmin))), bins));
    }
    void generate(const std::vector<double>& exclude) {
        const int n = (int)source.size();
        hmin = hmax = (float)source[0];
        for (int i = 0; i < n; i++) {
            if (source[i] > hmax) hmax = (float)source[i];
            if (source[i] < hmin) hmin = (float)source[i];
        }
        hist.assign(bins, 0.0);
        datapoints = (float)n;
        for (int i = 0; i < n; i++) {
            if (exclude[i] > 0) hist[get_bin((float)source[i]) - 1]++;
            else datapoints--;
        }
    }
    void generate_cdf() {
        cdf.assign(bins, 0.0);
        cdf[0] = hist[0] / datapoints;
        for (int i = 1; i < bins; i++) cdf[i] = cdf[i - 1] + hist[i] / datapoints;
    }
    void match(const Histogram& target) {
        if (excl.empty()) excl.assign(source.size(), 1.0);
        const float width = (target.hmax - target.hmin) / target.bins;
        for (size_t i = 0; i < source.size(); i++) {
            if (!(excl[i] > 0)) continue;
            const int bin = get_bin((float)source[i]);
            const double cdfval = cdf[bin - 1];
            int newbin = 1;
            if (bin == target.bins) newbin = bin;
            else
                while ((cdfval - target.cdf[newbin - 1]) > 1e-20) {
                    newbin++;
                    if (newbin == target.bins) break;
                }
            source[i] = target.hmin + (newbin - 1) * width;
            if (source[i] < target.hmin) source[i] = target.hmin;
            if (source[i] > target.hmax) source[i] = target.hmax;
        }
    }
};

// multivariate_histogram_normalization, src/reg_tools.cpp:750-810: every dimension of IN matched to the same dimension of REF,
// 256 bins; mask rows: one per dimension when the mask has that many, else its first row (:771-776,786-790)
inline void histogram_normalization(std::vector<double>& in /*[D][Vin]*/, int Vin, const std::vector<double>& ref /*[D][Vref]*/, int Vref, int D,
                                    const double* excl_in, int excl_in_rows, const double* excl_ref, int excl_ref_rows) {
    const int numbins = 256;
    for (int d = 0; d < D; d++) {
        std::vector<double> a(in.begin() + (size_t)d * Vin, in.begin() + (size_t)(d + 1) * Vin);
        std::vector<double> b(ref.begin() + (size_t)d * Vref, ref.begin() + (size_t)(d + 1) * Vref);
        std::vector<double> ea(Vin, 1.0), eb(Vref, 1.0);
        if (excl_in) { const double* r = excl_in + (size_t)(excl_in_rows > d ? d : 0) * Vin; ea.assign(r, r + Vin); }
        if (excl_ref) { const double* r = excl_ref + (size_t)(excl_ref_rows > d ? d : 0) * Vref; eb.assign(r, r + Vref); }
        Histogram ha(a, numbins), hb(b, numbins);
        ha.generate(ea); ha.excl = ea;
        hb.generate(eb); hb.excl = eb;
        ha.generate_cdf(); hb.generate_cdf();
        ha.match(hb);
        std::copy(ha.source.begin(), ha.source.end(), in.begin() + (size_t)d * Vin);
    }
}

// featurespace::variance_normalise, src/featurespace.cc:67-108: per dimension, Welford mean / variance (n-1) over the
// non-excluded vertices (first mask row), then (x - mean) / sqrt(var) (division only when var > 0); excluded vertices untouched
inline void variance_normalise(std::vector<double>& data /*[D][V]*/, int D, int V, const double* mask) {
    for (int d = 0; d < D; d++) {
        double* row = data.data() + (size_t)d * V;
        double mean = 0.0, var = 0.0;
        size_t n = 0;
        for (int i = 0; i < V; i++) {
            if (mask && !(mask[i] > 0.0)) continue;
            const double delta = row[i] - mean;
            mean += delta / (double)(n + 1);
            var += delta * (row[i] - mean);
            n++;
        }
        var /= (double)(n - 1);   // n == 1: division by the wrapped size_t of the reference (:90) is a division by 0 -> inf/nan there too
        for (int i = 0; i < V; i++) {
            if (mask && !(mask[i] > 0.0)) continue;
            row[i] -= mean;
            if (var > 0.0) row[i] /= std::sqrt(var);
        }
    }
}

struct FeatureOptions {
    std::vector<double> sigma;          // per mesh, featurespace::_sigma_in (default 5.0, src/featurespace.cc:6,13)
    float thr_lower = 0.f, thr_upper = 0.f;
    bool intensitynorm = false, varnorm = false, cut = false, exclude = false;
};
struct FeatureMesh { Mesh mesh; std::vector<double> data; int D = 0; };   // data[d*V + v]

// featurespace::initialize, src/featurespace.cc:18-65.  out_data[i] = DATA[i] ([D][Vico]); out_excl[i] = EXCL[i] (empty = none)
inline Mesh featurespace_initialize(int ico, const std::vector<FeatureMesh>& IN, const FeatureOptions& opt, std::vector<std::vector<double>>& out_data,
                                    std::vector<std::vector<double>>& out_excl) {
    const double RAD = 100.0;
    Mesh icotmp;
    if (ico > 0) {
        icotmp = make_mesh_from_icosa(ico, 1.0);
        Pt c(0, 0, 0);                                    // recentre, true_rescale ([EXT] A7, as oracle/ext/fsl_ext.h states them)
        for (const Pt& p : icotmp.V) c = c + p;
        c = c * (1.0 / std::max(1, icotmp.nvertices()));
        for (Pt& p : icotmp.V) p = p - c;
        for (Pt& p : icotmp.V) { p.normalize(); p = p * RAD; }
    }
    out_data.assign(IN.size(), {}); out_excl.assign(IN.size(), {});
    for (size_t i = 0; i < IN.size(); i++) {
        const FeatureMesh& m = IN[i];
        const int V = m.mesh.nvertices();
        if (ico == 0) icotmp = m.mesh;
        if (opt.exclude || opt.cut) out_excl[i] = create_exclusion(m.data, m.D, V, opt.thr_lower, opt.thr_upper);
        const double* mask = out_excl[i].empty() ? nullptr : out_excl[i].data();
        // the reference indexes the SAME mask by data-mesh vertex (metric_resample) and by template vertex (smoothing, :49; variance
        // normalisation, :76): only meaningful when the data mesh already has the template's vertices
        if (mask && V != icotmp.nvertices()) throw std::runtime_error("featurespace: exclusion masks need data meshes on the template's vertices");
        std::vector<double> tmp = metric_resample_excl(m.mesh, m.data, m.D, icotmp, mask);
        if (opt.sigma[i] > 0.0) tmp = smooth_data(icotmp.V, tmp, m.D, icotmp.V, opt.sigma[i], mask);
        out_data[i] = tmp;
    }
    const int Vt = icotmp.nvertices();
    if (opt.intensitynorm)
        for (size_t i = 1; i < IN.size(); i++)
            histogram_normalization(out_data[i], Vt, out_data[0], Vt, IN[i].D, out_excl[i].empty() ? nullptr : out_excl[i].data(), 1,
                                    out_excl[0].empty() ? nullptr : out_excl[0].data(), 1);
    if (opt.varnorm)
        for (size_t i = 0; i < IN.size(); i++) variance_normalise(out_data[i], IN[i].D, Vt, out_excl[i].empty() ? nullptr : out_excl[i].data());
    return icotmp;
}

}  // namespace orc
