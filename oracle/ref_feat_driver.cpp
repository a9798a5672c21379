// ref_feat_driver.cpp — TEST INFRASTRUCTURE.  Flat entry point over the reference's OWN featurespace
// (/root/reference/src/featurespace.cc + src/reg_tools.cpp, compiled where they lie into oracle/_ref/libref_feat.so by
// oracle/Makefile `reffeat`): featurespace(datain, dataref), the setters of src/featurespace.h:23-28, initialize().  The file
// names the reference would read (set_data -> Mesh::load, src/reg_tools.cpp:812-830) are keys of an in-memory registry here.
#include <map>
#include <mutex>

#include "featurespace.h"  // /root/reference/src/featurespace.h

namespace {
std::map<std::string, NEWMAT::Matrix>& registry() { static std::map<std::string, NEWMAT::Matrix> r; return r; }
newresampler::Mesh mesh_from(const double* xyz, int nv, const int* tris, int nf) {
    orc::Mesh M;
    M.V.resize(nv);
    for (int i = 0; i < nv; i++) M.V[i] = orc::Pt(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    M.F.resize(nf);
    for (int i = 0; i < nf; i++) M.F[i] = {tris[3 * i], tris[3 * i + 1], tris[3 * i + 2]};
    if (nf > 0) M.build_adjacency();
    return newresampler::Mesh(M);
}
NEWMAT::Matrix matrix_from(const double* d, int rows, int cols) {
    NEWMAT::Matrix m(rows, cols);
    for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) m(r + 1, c + 1) = d[(size_t)r * cols + c];
    return m;
}
}  // namespace

extern "C" {

// Returns the template's vertex count; out0/out1 = DATA[0]/DATA[1] as [dims][V]; excl0/excl1 (may be null) = EXCL masks.
// -1 on an exception (message on stderr).
int ref_featurespace_initialize(int ico, const double* xyz0, int nv0, const int* tris0, int nf0, const double* data0, const double* xyz1, int nv1,
                                const int* tris1, int nf1, const double* data1, int dims, double sigma0, double sigma1, float thr_lower,
                                float thr_upper, int intensitynorm, int varnorm, int cut, int exclude, int threads, double* out0, double* out1,
                                double* excl0, double* excl1) {
    static std::mutex mtx;
    std::lock_guard<std::mutex> lock(mtx);
    try {
        registry()["mem:input"] = matrix_from(data0, dims, nv0);
        registry()["mem:reference"] = matrix_from(data1, dims, nv1);
        newresampler::Mesh::loader() = [](newresampler::Mesh& m, const std::string& name) {
            auto it = registry().find(name);
            if (it == registry().end()) throw newresampler::MeshException("unknown in-memory data key");
            m.set_pvalues(it->second);
        };
        std::vector<newresampler::Mesh> IN = {mesh_from(xyz0, nv0, tris0, nf0), mesh_from(xyz1, nv1, tris1, nf1)};
        newmeshreg::featurespace fs("mem:input", "mem:reference");
        fs.set_smoothing_parameters({sigma0, sigma1});
        fs.set_cutthreshold({thr_lower, thr_upper});
        fs.varnorm(varnorm != 0);
        fs.intensitynormalize(intensitynorm != 0, cut != 0);
        fs.set_nthreads(threads);
        newresampler::Mesh t = fs.initialize(ico, IN, exclude != 0);
        for (int n = 0; n < 2; n++) {
            double* out = n == 0 ? out0 : out1;
            const NEWMAT::Matrix M = fs.get_data_matrix(n);
            for (int r = 0; r < M.Nrows(); r++) for (int c = 0; c < M.Ncols(); c++) out[(size_t)r * M.Ncols() + c] = M(r + 1, c + 1);
            double* ex = n == 0 ? excl0 : excl1;
            auto E = n == 0 ? fs.get_input_excl() : fs.get_reference_excl();
            if (ex && E) for (int i = 0; i < E->nvertices(); i++) ex[i] = E->get_pvalue(i);
        }
        return t.nvertices();
    } catch (const std::exception& e) {
        std::cerr << "ref_featurespace_initialize: " << e.what() << std::endl;
        return -1;
    }
}

}  // extern "C"
