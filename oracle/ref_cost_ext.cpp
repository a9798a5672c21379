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

#ifndef REF_FEAT_BUILD   // libref_feat.so links the reference's own src/featurespace.cc instead (its constructors + initialize)
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
#endif

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

static std::vector<double> mask_of(const std::shared_ptr<Mesh>& EXCL, int V) {
    std::vector<double> m;
    if (!EXCL) return m;
    if (EXCL->nvertices() != V) throw MeshException("exclusion mask does not match the mesh");
    m.resize(V);
    for (int i = 0; i < V; i++) m[i] = EXCL->get_pvalue(i);
    return m;
}
static void values_of(const Mesh& in, std::vector<double>& vals, int& dims, bool& column) {
    const NEWMAT::Matrix pv = in.get_pvalues();
    column = pv.Ncols() == 1 && pv.Nrows() == in.nvertices();  // set_pvalues(ColumnVector), src/DiscreteCostFunction.cpp:381
    dims = column ? 1 : pv.Nrows();
    if (!column && pv.Ncols() != in.nvertices()) throw MeshException("per-vertex values do not match the mesh");
    const int Vin = in.nvertices();
    vals.resize((size_t)dims * Vin);
    for (int d = 0; d < dims; d++)
        for (int v = 0; v < Vin; v++) vals[(size_t)d * Vin + v] = column ? pv(v + 1, 1) : pv(d + 1, v + 1);
}
static Mesh with_values(const Mesh& ref, const std::vector<double>& res, int dims, bool column) {
    const int Vref = ref.nvertices();
    NEWMAT::Matrix out(column ? Vref : dims, column ? 1 : Vref);
    for (int d = 0; d < dims; d++)
        for (int i = 0; i < Vref; i++) {
            if (column) out(i + 1, 1) = res[i]; else out(d + 1, i + 1) = res[(size_t)d * Vref + i];
        }
    Mesh o = ref;
    o.set_pvalues(out);
    return o;
}

// A8: adaptive barycentric resampling (oracle/geom.h metric_resample): per-vertex values of `in` at the vertices of `ref`
Mesh metric_resample(const Mesh& in, const Mesh& ref, int, std::shared_ptr<Mesh> EXCL) {
    std::vector<double> vals; int dims; bool column;
    values_of(in, vals, dims, column);
    const std::vector<double> mask = mask_of(EXCL, in.nvertices());
    return with_values(ref, orc::metric_resample_excl(in.synced(), vals, dims, ref.synced(), mask.empty() ? nullptr : mask.data()), dims, column);
}

// A9: Gaussian smoothing (oracle/feat.h smooth_data)
Mesh smooth_data(const Mesh& in, const Mesh& ref, double sigma, int, std::shared_ptr<Mesh> EXCL) {
    std::vector<double> vals; int dims; bool column;
    values_of(in, vals, dims, column);
    const std::vector<double> mask = mask_of(EXCL, in.nvertices());
    const orc::Mesh a = in.synced(), b = ref.synced();
    return with_values(ref, orc::smooth_data(a.V, vals, dims, b.V, sigma, mask.empty() ? nullptr : mask.data()), dims, column);
}

// A11: exclusion mask from a value band (oracle/feat.h create_exclusion)
Mesh create_exclusion(const Mesh& data_mesh, float lower, float upper) {
    std::vector<double> vals; int dims; bool column;
    values_of(data_mesh, vals, dims, column);
    return with_values(data_mesh, orc::create_exclusion(vals, dims, data_mesh.nvertices(), lower, upper), 1, false);
}

}  // namespace newresampler
