// featurespace.h — mirror of the reference's featurespace (src/featurespace.h, src/featurespace.cc).
//
// Accessor contract (src/featurespace.h:30-38): C x V feature matrices with 1-based (dim, vertex) access.
// Pre-processing (src/featurespace.cc:18-108): initialize(ico, IN, exclude) resamples every mesh's data onto the level's ico
// template ([EXT] A8 metric_resample), smooths it ([EXT] A9 smooth_data), optionally histogram-matches the input to the
// reference (src/reg_tools.cpp:750-810, [EXT] A10) and variance-normalises (src/featurespace.cc:67-108) — here every one of
// those steps runs on the GPU through the C ABI (include/msm_b200.h; csrc/features.cuh), there is no CPU fallback.
// What stays with the caller is file I/O: the reference's constructors take file names and set_data() loads them
// (src/reg_tools.cpp:812-830); this mirror takes the loaded matrices.
#pragma once
#include <memory>
#include <vector>

#include "mesh_warp.h"
#include "reg_tools.h"

#ifndef RAD
#define RAD 100.0   // src/featurespace.h:7
#endif

namespace newmeshreg {

class featurespace {
public:
    featurespace() = default;
    // channel-major arrays: input[d*Vin + v], ref[d*Vref + v] — DATA[0] / DATA[1] as set_data leaves them.  Without a call of
    // initialize() they are used as they are (callers that hand over already pre-processed data).
    featurespace(const double* input, int Vin, const double* ref, int Vref, int dim)
        : in_(input, input + (size_t)dim * Vin), ref_(ref, ref + (size_t)dim * Vref), Vin_(Vin), Vref_(Vref), dim_(dim) {}

    //---INITIALISE--- src/featurespace.cc:18-65 (IN = {input mesh, reference mesh}); returns the template the data now lives on
    newresampler::Mesh initialize(int ico, std::vector<newresampler::Mesh>& IN, bool exclude) {
        if (IN.size() != 2) throw MeshregException("featurespace::Initialize do not have the same number of datasets and surface meshes");
        if (IN[0].nvertices() != Vin_ || IN[1].nvertices() != Vref_) throw MeshregException("data does not match mesh dimensions");   // set_data, src/reg_tools.cpp:822-828
        newresampler::Mesh icotmp;
        if (ico > 0) {
            icotmp = newresampler::make_mesh_from_icosa(ico);
            newresampler::recentre(icotmp);
            newresampler::true_rescale(icotmp, RAD);
        }
        std::lock_guard<std::mutex> lock(feature_mutex());
        msm_ctx* ctx = feature_context();
        excl_.assign(2, nullptr);
        std::vector<double>* data[2] = {&in_, &ref_};
        int* nv[2] = {&Vin_, &Vref_};
        for (int i = 0; i < 2; i++) {
            const int V = IN[i].nvertices();
            if (ico == 0) icotmp = IN[i];
            if (exclude || _cut) {
                auto m = std::make_shared<std::vector<double>>(V);
                check(ctx, msm_create_exclusion(ctx, data[i]->data(), dim_, V, _fthreshold[0], _fthreshold[1], m->data()));
                // the reference indexes the same mask by data-mesh vertex (:46) and by template vertex (:49,76): only meaningful
                // when the data mesh already has the template's vertices
                if (V != icotmp.nvertices()) throw MeshregException("featurespace: exclusion masks need data meshes on the template's vertices");
                excl_[i] = m;
            }
            const double* mask = excl_[i] ? excl_[i]->data() : nullptr;
            std::vector<double> tmp = newresampler::metric_resample_excl(IN[i], data[i]->data(), dim_, icotmp, mask);
            const int Vt = icotmp.nvertices();
            if (_sigma_in[i] > 0.0) {
                std::vector<double> sm((size_t)dim_ * Vt);
                check(ctx, msm_smooth_data(ctx, icotmp.coord_data(), Vt, tmp.data(), dim_, mask, icotmp.coord_data(), Vt, _sigma_in[i], sm.data()));
                tmp.swap(sm);
            }
            data[i]->swap(tmp);
            *nv[i] = Vt;
        }
        if (_intensitynorm)   // :55-57: DATA[1] (the reference's features) is matched, in place, to the distributions of DATA[0]
            check(ctx, msm_histogram_match(ctx, ref_.data(), Vref_, in_.data(), Vin_, dim_, excl_[1] ? excl_[1]->data() : nullptr, 1,
                                           excl_[0] ? excl_[0]->data() : nullptr, 1, 256));
        if (_varnorm) {
            check(ctx, msm_variance_normalise(ctx, in_.data(), dim_, Vin_, excl_[0] ? excl_[0]->data() : nullptr));
            check(ctx, msm_variance_normalise(ctx, ref_.data(), dim_, Vref_, excl_[1] ? excl_[1]->data() : nullptr));
        }
        return icotmp;
    }

    //---SET--- src/featurespace.h:23-28
    void set_smoothing_parameters(const std::vector<double>& s) { _sigma_in = s; _sigma_in.resize(2, 5.0); }
    void set_cutthreshold(const std::vector<float>& thr) { _fthreshold = thr; _fthreshold.resize(2, 0.0f); }
    void varnorm(bool norm) { _varnorm = norm; }
    void is_sparse(bool sp) { if (sp) throw MeshregException("featurespace: sparse data matrices are file I/O and stay with the caller"); }
    void intensitynormalize(bool norm, bool exclcut) { _intensitynorm = norm; _cut = exclcut; }
    void set_nthreads(int) {}

    //---GET--- src/featurespace.h:30-38
    int get_dim() const { return dim_; }
    double get_input_val(int dim, int vert) const { return in_[(size_t)(dim - 1) * Vin_ + (vert - 1)]; }
    double get_ref_val(int dim, int vert) const { return ref_[(size_t)(dim - 1) * Vref_ + (vert - 1)]; }
    std::shared_ptr<std::vector<double>> get_input_excl() const { return excl_.empty() ? nullptr : excl_[0]; }
    std::shared_ptr<std::vector<double>> get_reference_excl() const { return excl_.empty() ? nullptr : excl_[1]; }
    const double* input_data() const { return in_.data(); }
    const double* ref_data() const { return ref_.data(); }
    int input_vertices() const { return Vin_; }
    int ref_vertices() const { return Vref_; }
    // group registration: one data matrix per subject on its own data mesh (src/DiscreteGroupModel.cpp:79)
    void add_subject(const double* data, int V) { subj_.emplace_back(data, data + V); }
    const std::vector<double>& get_data_matrix(int subject) const { return subj_[subject]; }
    int subjects() const { return (int)subj_.size(); }
private:
    static std::mutex& feature_mutex() { static std::mutex m; return m; }
    static msm_ctx* feature_context() {   // a context of its own for these stateless calls (device, stream, error string)
        static msm_ctx* ctx = nullptr;
        if (!ctx) {
            int dev = 0;
            if (const char* e = std::getenv("MSM_DEVICE")) dev = std::atoi(e);
            if (msm_create(dev, &ctx) != 0) { ctx = nullptr; throw MeshregException("featurespace: no CUDA device (no CPU fallback)"); }
        }
        return ctx;
    }
    static void check(msm_ctx* ctx, int rc) {
        if (rc != 0) { static thread_local std::string keep; keep = msm_last_error(ctx); throw MeshregException(keep.c_str()); }
    }
    std::vector<double> in_, ref_;
    std::vector<std::vector<double>> subj_;
    std::vector<std::shared_ptr<std::vector<double>>> excl_;
    int Vin_ = 0, Vref_ = 0, dim_ = 0;
    std::vector<double> _sigma_in = {5.0, 5.0};      // src/featurespace.cc:6
    std::vector<float> _fthreshold = {0.0f, 0.0f};   // :7
    bool _intensitynorm = false, _varnorm = false, _cut = false;
};

}  // namespace newmeshreg
