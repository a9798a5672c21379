// featurespace.h — the accessor contract of the reference's featurespace (src/featurespace.h:30-32): C x V feature
// matrices with 1-based (dim, vertex) access.  Loading, resampling to the ico template, smoothing and normalisation
// (src/featurespace.cc:18-108) are pre-processing outside the hot path: callers hand over the
// already resampled matrices.
#pragma once
#include <memory>
#include <vector>

#include "reg_tools.h"

namespace newmeshreg {

class featurespace {
public:
    featurespace() = default;
    // channel-major arrays: input[d*Vin + v], ref[d*Vref + v]
    featurespace(const double* input, int Vin, const double* ref, int Vref, int dim)
        : in_(input, input + (size_t)dim * Vin), ref_(ref, ref + (size_t)dim * Vref), Vin_(Vin), Vref_(Vref), dim_(dim) {}
    int get_dim() const { return dim_; }
    double get_input_val(int dim, int vert) const { return in_[(size_t)(dim - 1) * Vin_ + (vert - 1)]; }
    double get_ref_val(int dim, int vert) const { return ref_[(size_t)(dim - 1) * Vref_ + (vert - 1)]; }
    const double* input_data() const { return in_.data(); }
    const double* ref_data() const { return ref_.data(); }
    int input_vertices() const { return Vin_; }
    int ref_vertices() const { return Vref_; }
    // group registration: one data matrix per subject on its own data mesh (src/DiscreteGroupModel.cpp:79)
    void add_subject(const double* data, int V) { subj_.emplace_back(data, data + V); }
    const std::vector<double>& get_data_matrix(int subject) const { return subj_[subject]; }
    int subjects() const { return (int)subj_.size(); }
private:
    std::vector<double> in_, ref_;
    std::vector<std::vector<double>> subj_;
    int Vin_ = 0, Vref_ = 0, dim_ = 0;
};

}  // namespace newmeshreg
