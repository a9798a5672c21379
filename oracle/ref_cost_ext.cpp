// ref_cost_ext.cpp — TEST INFRASTRUCTURE.  The out-of-line pieces of the [EXT] stand-ins (oracle/ext/fsl_ext.h) that
// the reference's own sources link against when they are compiled where they lie into oracle/_ref/libref_cost.so
// (oracle/Makefile `refcost`).  Nothing here evaluates a cost: the cost arithmetic in that library is the reference's.
#include "featurespace.h"  // /root/reference/src/featurespace.h (declaration only; featurespace.cc needs FSL file I/O)

#include "oracle.h"

namespace refcost {
std::map<std::string, std::shared_ptr<MISCMATHS::BFMatrix>>& feature_registry() {
    static std::map<std::string, std::shared_ptr<MISCMATHS::BFMatrix>> reg;
    return reg;
}
}  // namespace refcost

namespace newmeshreg {

// src/featurespace.h:14-15 declares file-name constructors only.  Here the "file names" are keys of the in-memory
// registry filled by the harness (host/harness.cpp make_features): DATA[0] = input (source) features, DATA[1] =
// reference (target) features, rows = feature dimensions, columns = vertices — as src/featurespace.cc:5-20 leaves them.
featurespace::featurespace(const std::string& datain, const std::string& dataref) {
    auto& reg = refcost::feature_registry();
    if (!reg.count(datain) || !reg.count(dataref)) throw MeshregException("featurespace: unknown in-memory data key");
    DATA = {reg[datain], reg[dataref]};
    EXCL = {nullptr, nullptr};
    CMfile_in = {datain, dataref};
}
featurespace::featurespace(const std::vector<std::string>& datalist) {
    auto& reg = refcost::feature_registry();
    for (const auto& k : datalist) {
        if (!reg.count(k)) throw MeshregException("featurespace: unknown in-memory data key");
        DATA.push_back(reg[k]);
        EXCL.push_back(nullptr);
        CMfile_in.push_back(k);
    }
}

}  // namespace newmeshreg

namespace newresampler {

// A6 (oracle/oracle.cpp barycentric_mesh_interpolation): every vertex of `up` is located on START and carried to the
// same barycentric position on END, then put back on the sphere of its original radius.
void barycentric_mesh_interpolation(Mesh& up, const Mesh& START, const Mesh& END, int) {
    std::vector<orc::Pt> pts(up.nvertices()), fin(END.nvertices());
    for (int i = 0; i < up.nvertices(); i++) pts[i] = up.get_coord(i);
    for (int i = 0; i < END.nvertices(); i++) fin[i] = END.get_coord(i);
    orc::barycentric_mesh_interpolation(pts, START.synced(), fin);
    for (int i = 0; i < up.nvertices(); i++) up.set_coord(i, Point(pts[i]));
}

// A8: adaptive barycentric resampling (oracle/geom.h metric_resample): per-vertex values of `in` at the vertices of `ref`
Mesh metric_resample(const Mesh& in, const Mesh& ref, int) {
    const NEWMAT::Matrix pv = in.get_pvalues();
    const bool column = pv.Ncols() == 1 && pv.Nrows() == in.nvertices();  // set_pvalues(ColumnVector), src/DiscreteCostFunction.cpp:381
    const int dims = column ? 1 : pv.Nrows();
    if (!column && pv.Ncols() != in.nvertices()) throw MeshException("metric_resample: values do not match the mesh");
    const int Vin = in.nvertices(), Vref = ref.nvertices();
    std::vector<double> vals((size_t)dims * Vin);
    for (int d = 0; d < dims; d++)
        for (int v = 0; v < Vin; v++) vals[(size_t)d * Vin + v] = column ? pv(v + 1, 1) : pv(d + 1, v + 1);
    const std::vector<double> res = orc::metric_resample(in.synced(), vals, dims, ref.synced());
    NEWMAT::Matrix out(column ? Vref : dims, column ? 1 : Vref);
    for (int d = 0; d < dims; d++)
        for (int i = 0; i < Vref; i++) {
            if (column) out(i + 1, 1) = res[i]; else out(d + 1, i + 1) = res[(size_t)d * Vref + i];
        }
    Mesh o = ref;
    o.set_pvalues(out);
    return o;
}

}  // namespace newresampler
