// CPU test helper (tests/test_host_cpu.py): the mirror's mesh container (host/src/newresampler_shim.h) behind a tiny C ABI —
// icosphere generation (cached per level), lazily built adjacency, copy-on-write coordinates, per-vertex values.
#include <cstring>

#include "../msm_newmeshreg_b200/host/src/newresampler_shim.h"

using namespace newresampler;

extern "C" {

// counts of the level's icosphere; xyz / tris filled when non-null
void shim_icosphere(int level, int* nv, int* nf, double* xyz, int* tris) {
    const Mesh m = make_mesh_from_icosa(level);
    *nv = m.nvertices(); *nf = m.ntriangles();
    if (xyz) { auto c = m.coords(); std::memcpy(xyz, c.data(), c.size() * sizeof(double)); }
    if (tris) { auto f = m.faces(); std::memcpy(tris, f.data(), f.size() * sizeof(int)); }
}
// neighbours / incident triangles of every vertex in the container's order: CSR
void shim_adjacency(const double* xyz, int nv, const int* tris, int nf, int* nbr_off, int* nbr, int* tri_off, int* tri) {
    Mesh m;
    m.set(xyz, nv, tris, nf);
    const Mesh copy = m;                       // adjacency is shared by copies and built on first use, by whichever copy asks
    nbr_off[0] = tri_off[0] = 0;
    for (int v = 0; v < nv; v++) {
        int n = nbr_off[v];
        for (auto it = copy.nbegin(v); it != copy.nend(v); ++it) nbr[n++] = *it;
        nbr_off[v + 1] = n;
        int t = tri_off[v];
        for (auto it = m.tIDbegin(v); it != m.tIDend(v); ++it) tri[t++] = *it;
        tri_off[v + 1] = t;
    }
}
// 0 when: a cached icosphere handed out twice is two independent meshes (moving a vertex of one leaves the other and the cache
// alone), rescaling works on a copy, pvalues travel with copies until replaced
int shim_semantics(int level) {
    Mesh a = make_mesh_from_icosa(level), b = make_mesh_from_icosa(level);
    const Point before = b.get_coord(0);
    a.set_coord(0, Point(1, 2, 3));
    true_rescale(a, 100.0);
    if (b.get_coord(0).X != before.X || b.get_coord(0).Y != before.Y || b.get_coord(0).Z != before.Z) return 1;
    const Mesh c = make_mesh_from_icosa(level);
    if (c.get_coord(0).X != before.X || std::fabs(c.get_coord(1).norm() - 1.0) > 1e-12) return 2;
    if (std::fabs(a.get_coord(1).norm() - 100.0) > 1e-9) return 3;
    NEWMAT::Matrix pv(2, b.nvertices());
    pv = 1.5;
    b.set_pvalues(pv);
    Mesh d = b;
    if (d.get_dimension() != 2 || d.npvalues() != b.nvertices() || d.get_pvalue(3, 1) != 1.5) return 4;
    NEWMAT::Matrix other(1, b.nvertices());
    other = 2.5;
    d.set_pvalues(other);
    if (b.get_dimension() != 2 || b.get_pvalue(0, 0) != 1.5 || d.get_pvalue(0) != 2.5) return 5;
    if (a.get_total_neighbours(0) != 5 || c.get_total_neighbours(12 < c.nvertices() ? 12 : 0) < 5) return 6;
    return 0;
}

}  // extern "C"
