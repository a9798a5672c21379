/* msm_b200.h — C ABI of the B200-native cost-evaluation path of newMSM (rbesenczi/msm-newmeshreg).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Every entry
 * point cites the reference interface it replaces (paths relative to the reference tree).  The host-side C++
 * mirror of the reference classes (msm_newmeshreg_b200/host/) is a thin layer over these calls.
 *
 * Conventions
 *  - All host arrays are caller-owned; the context owns every device buffer.  A context is NOT thread-safe (one
 *    host thread at a time drives it), table look-ups on the host side are.  DIFFERENT contexts share no mutable
 *    state and every call selects its context's device, so they may be driven concurrently from different
 *    threads — several per GPU (independent subjects, each on its own stream) or one per GPU.
 *  - Coordinates are double[n][3] on the sphere of radius 100 (RAD, src/DiscreteCostFunction.h:9).
 *  - Feature matrices are channel-major double[C][V], i.e. FEAT->get_ref_val(d+1, v+1) == ref[d*V + v]
 *    (src/featurespace.h:30-32).
 *  - Table layouts are the reference's: unary[label*N + node] (src/DiscreteCostFunction.cpp:312),
 *    pair[pair*L*L + labelB*L + labelA] (:303), triplet[((t*L + la)*L + lb)*L + lc].
 *  - Every function returns 0 on success, non-zero on error; msm_last_error() gives the message.
 *    There is NO CPU fallback: without a CUDA device msm_create fails.
 */
#ifndef MSM_B200_H
#define MSM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct msm_ctx msm_ctx;

/* NonLinearSRegDiscreteCostFunction::set_parameters (src/DiscreteCostFunction.cpp:121-136) plus the scalars the
 * model hands over through set_spacings (src/DiscreteCostFunction.h:178) and set_initial_angles (:95-102). */
typedef struct msm_params {
    double lambda;     /* "lambda"          _reglambda */
    double rexp;       /* "exponent"        _rexp      */
    double mu;         /* "shearmodulus"    _mu        */
    double kappa;      /* "bulkmodulus"     _kappa     */
    double k_exp;      /* "kexponent"       _k_exp     */
    double range;      /* "range"           _controlptrange */
    int simmeasure;    /* "simmeasure": sign of the unary term (:444-447, :522-525) */
    int rmode;         /* "regularisermode": 2 angular, 3 strain (:170-201); 1 = pairwise model */
    int dweight;       /* "weight"  (:182) */
    int anorm;         /* "anorm"   (:190) */
    double meanangle;  /* _MEANANGLE (:101) */
    double mvdmax;     /* MVDmax (set_spacings) — pairwise regulariser (:262) */
    int group;         /* 0: NonLinearSRegDiscreteCostFunction; 1: DiscreteGroupCostFunction triplet rules
                          (fold -> 1e7, k_exp forced to 2, src/DiscreteGroupCostFunction.cpp:23,29) */
    int bary_orthogonal; /* A4 switch: 0 radial projection (default), 1 orthogonal — locate/unary */
} msm_params;

void msm_default_params(msm_params* p);

/* lifetime ------------------------------------------------------------------------------------------------ */
int msm_create(int device, msm_ctx** out);
void msm_destroy(msm_ctx* ctx);
const char* msm_last_error(const msm_ctx* ctx); /* ctx may be NULL: last creation error */
/* Use an existing CUDA stream (cudaStream_t as void*); NULL = the context's own stream. */
int msm_set_stream(msm_ctx* ctx, void* cuda_stream);
int msm_synchronize(msm_ctx* ctx);

/* set_parameters (src/DiscreteCostFunction.cpp:121-136) */
int msm_set_params(msm_ctx* ctx, const msm_params* p);

/* set_meshes/_TARGET + set_featurespace/get_ref_val + set_octrees (src/DiscreteCostFunction.h:166-180,
 * src/DiscreteModel.cpp:88): the target must be a regular icosphere (the reference always resamples data onto
 * one: src/featurespace.cc:27-31,46); any vertex/face numbering and orientation is accepted — the icosahedral
 * hierarchy that replaces the Octree is rebuilt from the geometry.  feat may be NULL (locate only). */
int msm_set_target(msm_ctx* ctx, const double* xyz, int V, const int* tris, int F, const double* feat, int C);

/* set_meshes/_SOURCE + FEAT->get_input_val + set_dataaffintyweighting (src/DiscreteCostFunction.h:144,166-168).
 * cfweight: double[wrows][V] with wrows in {0,1,C}; missing rows weigh 1.0 (:404-407,:474-477).
 * multivariate selects the Multivariate cost function (src/DiscreteModel.cpp:22-36). */
int msm_set_source(msm_ctx* ctx, const double* xyz, int V, const double* feat, int C, const double* cfweight, int wrows,
                   int multivariate);
/* reset_source (src/DiscreteCostFunction.h:181): new coordinates of the deforming source mesh. */
int msm_reset_source(msm_ctx* ctx, const double* xyz);

/* set_meshes/_CPgrid,_oCPgrid,_ORIG + set_spacings (src/DiscreteCostFunction.h:166-168,178).
 * orig_xyz = _ORIG coordinates indexed by CP id (nested numbering, :175-177,:195-197); NULL = xyz.
 * abs_weights = AbsoluteWeights per CP (resample_weights :364-383, [EXT] metric_resample); NULL = all 1. */
int msm_set_cpgrid(msm_ctx* ctx, const double* xyz, int N, const int* tris, int T, const double* maxsep,
                   const double* orig_xyz, const double* abs_weights);
/* AbsoluteWeights alone (resample_weights runs in every get_source_data, src/DiscreteCostFunction.cpp:364-383,409,480): double[N]. */
int msm_set_abs_weights(msm_ctx* ctx, const double* abs_weights);
/* reset_CPgrid (src/DiscreteCostFunction.h:182) */
int msm_reset_cpgrid(msm_ctx* ctx, const double* xyz);
/* Multi-GPU: this context evaluates unary rows / patches only for CPs [k_begin,k_end). */
int msm_set_shard(msm_ctx* ctx, int k_begin, int k_end);

/* set_labels(labels, ROT) (src/DiscreteCostFunction.h:170-173).  ROT[k] = rotation(centre -> CP_k) is
 * recomputed on the device from `centre` and the current CP grid (src/DiscreteModel.cpp:284-294). */
int msm_set_labels(msm_ctx* ctx, const double* labels, int L, const double* centre);

/* setPairs / setTriplets (src/DiscreteCostFunction.h:25-26): int[2P] / int[3T], ids sorted ascending. */
int msm_set_pairs(msm_ctx* ctx, const int* pairs, int P);
int msm_set_triplets(msm_ctx* ctx, const int* triplets, int T);
/* Group model (src/DiscreteGroupCostFunction.h:11-19): S subjects x N_subj control points; cp grid passed to
 * msm_set_cpgrid is the concatenation [S*N_subj]; _ORIG is indexed by the per-subject id. */
int msm_set_group(msm_ctx* ctx, int S, int N_subj, int T_subj);

/* get_source_data (src/DiscreteCostFunction.cpp:394-410 / :458-481): builds the per-CP patch lists for the
 * current source and CP coordinates.  Exact-tie pairs are resolved with the reference's own fp64 formula. */
int msm_build_patches(msm_ctx* ctx);
/* The same build without a host wait, for callers that keep the device busy (get_source_data followed by
 * computeUnaryCosts in one outer iteration, src/mesh_registration.cpp:344-351): the kernels are enqueued and the build's
 * one status word (boundary ties seen / a row outgrew its capacity) follows them to the host.  It is read
 *   - by msm_unary_table / msm_get_patches / msm_patch_count on their own (they wait for the device anyway; a flagged
 *     build is repeated through the exact path and the table recomputed before they return), or
 *   - by msm_patches_resolve, which callers of the device-resident entry points (msm_unary_table_dev / _peers) must call
 *     before trusting what they launched: *redone = 1 means the patches were rebuilt and those launches must be repeated.
 * msm_build_patches == msm_build_patches_async + msm_patches_resolve. */
int msm_build_patches_async(msm_ctx* ctx);
int msm_patches_resolve(msm_ctx* ctx, int* redone);
/* _sourceinrange as CSR: offsets int[N+1] (only this shard's rows are non-empty), idx int[total]. */
int msm_patch_count(msm_ctx* ctx, long long* total);
/* Introspection for tests/bench: how many builds of this context took the one-pass path (second and later builds of a level,
 * no boundary ties: csrc/patchfast.cuh) and how many the exact two-pass path.  Both give identical patches and tables. */
int msm_patch_build_stats(const msm_ctx* ctx, long long* one_pass, long long* two_pass);
int msm_get_patches(msm_ctx* ctx, int* offsets, int* idx);

/* computeUnaryCosts (src/DiscreteCostFunction.cpp:306-313): out double[L][N] (rows of other shards untouched). */
int msm_unary_table(msm_ctx* ctx, double* out);
/* Device-resident variant: d_out = float[L][k_end-k_begin] in device memory, launched on the ctx stream. */
int msm_unary_table_dev(msm_ctx* ctx, float* d_out);

/* computeTripletCost (src/DiscreteCostFunction.cpp:138-254; group: src/DiscreteGroupCostFunction.cpp:9-30).
 *  _costs : n arbitrary queries q[n][4] = (triplet, la, lb, lc)
 *  _table : full float[T1-T0][L][L][L] look-up table for triplets [t0,t1) (serves unmodified Fusion::optimize)
 *  _moves : the 8 fusion-move costs of include/Fusion/Fusion.h:187-194 for every triplet, for label alpha
 *           (out double[T][8]) or, with alpha = -1, for all labels (out double[L][T][8]). */
int msm_triplet_costs(msm_ctx* ctx, int n, const int* q, double* out);
int msm_triplet_table(msm_ctx* ctx, int t0, int t1, float* out);
int msm_triplet_table_dev(msm_ctx* ctx, int t0, int t1, float* d_out);
int msm_triplet_moves(msm_ctx* ctx, const int* labeling, int alpha, double* out);
int msm_triplet_moves_f32(msm_ctx* ctx, const int* labeling, int alpha, float* out); /* same table, fp32: half the read-back */
int msm_triplet_moves_dev(msm_ctx* ctx, const int* d_labeling, int alpha, float* d_out);
/* Multi-GPU: the same for the triplet shard [t0,t1): d_out float[L or 1][t1-t0][8]. */
int msm_triplet_moves_range_dev(msm_ctx* ctx, const int* d_labeling, int alpha, int t0, int t1, float* d_out);

/* Fused exchange for control-point / triplet sharding (replaces "kernel, then all-gather, then
 * re-layout"): the kernels store every finished entry straight into the FULL table of EVERY rank.  peer_tables[r]
 * (r < world <= 16) is a device pointer, valid on this context's GPU, to rank r's buffer — CUDA IPC or symmetric-memory
 * mappings over NVLink; peer_tables[own rank] is the local buffer.
 *   msm_unary_table_peers        : float[L][N], rows [k_begin,k_end) of this context's shard (msm_set_shard)
 *   msm_triplet_moves_range_peers: float[L or 1][T][8], triplets [t0,t1)
 * multicast_table (may be NULL): the NVSwitch multicast address of the same buffer (cuMulticast* / symmetric memory's
 * multicast_ptr); when given, each entry is written with ONE multimem.st that the switch replicates to every rank (NVLS)
 * instead of `world` peer stores.
 * Launched on the context stream; the caller makes the ranks wait for each other afterwards (a barrier on the stream). */
int msm_unary_table_peers(msm_ctx* ctx, void* const* peer_tables, int world, void* multicast_table);
int msm_triplet_moves_range_peers(msm_ctx* ctx, const int* d_labeling, int alpha, int t0, int t1, void* const* peer_tables, int world,
                                  void* multicast_table);

/* computePairwiseCosts (src/DiscreteCostFunction.cpp:256-304): out float[P][L][L] in the reference layout. */
int msm_pair_table(msm_ctx* ctx, float* out);
int msm_pair_table_dev(msm_ctx* ctx, float* d_out);

/* Group data term: DiscreteGroupCostFunction::set_patch_data + computePairwiseCost
 * (src/DiscreteGroupCostFunction.h:39-41, .cpp:32-54).  The std::map<int,double> patches arrive as CSR with ascending
 * keys: map m = (subject*N_subj + node)*L + label holds (keys[off[m]..off[m+1]), vals[...]).  Pairs come from
 * msm_set_pairs (global node ids).
 *  _costs : n queries q[n][3] = (pair, labelA, labelB)
 *  _moves : the 4 fusion-move costs of include/Fusion/Fusion.h:169-172 for every pair and label alpha, out[P][4]. */
int msm_set_group_patches(msm_ctx* ctx, long long nmaps, const long long* off, const int* keys, const double* vals);
int msm_group_pair_costs(msm_ctx* ctx, int n, const int* q, double* out);
int msm_group_pair_moves(msm_ctx* ctx, const int* labeling, int alpha, double* out);
/*  _table : the whole float[P][L][L] table, layout pair*L*L + labelB*L + labelA (src/DiscreteCostFunction.cpp:303): serves
 *           computePairwiseCost of the unmodified Fusion::optimize as a look-up */
int msm_group_pair_table(msm_ctx* ctx, float* out);

/* Octree::get_closest_triangle + barycentric weights on the target ([EXT] A3/A4; call sites
 * src/DiscreteCostFunction.cpp:421-433).  tri = index into the `tris` given to msm_set_target; w double[M][3]
 * in that triangle's vertex order. */
int msm_locate(msm_ctx* ctx, const double* pts, int M, int* tri, double* w);

/* newresampler::barycentric_mesh_interpolation(SPH_up, SPH_low_init, SPH_low_final) ([EXT] A8; call sites
 * src/mesh_registration.cpp:390 (warp the source by the control-grid update), :305-306 (level transition),
 * src/DiscreteModel.h:123): every point is located in the START mesh (any closed spherical triangle mesh, deformed or
 * not), and moved to the same barycentric combination of the END mesh's vertices, pushed back to `radius`.
 * tri_out / w_out (optional): located triangle and weights in that triangle's vertex order. */
int msm_mesh_interpolate(msm_ctx* ctx, const double* pts, int M, const double* from_xyz, int V, const int* tris, int F,
                         const double* to_xyz, double radius, double* out, int* tri_out, double* w_out);

/* ---- Feature pre-processing per resolution level (featurespace::initialize, src/featurespace.cc:18-65).  All arrays are
 * caller-owned host doubles, data as [D][V] (dimension-major, the transpose-free layout of BFMatrix rows); masks hold > 0 for
 * "use", <= 0 for "excluded".  The context only supplies the device, the stream and the error string.
 *
 * msm_smooth_data      : newresampler::smooth_data(in, ref, sigma, threads, EXCL) ([EXT] A9, call site src/featurespace.cc:49):
 *                        out[d][k] = sum_j w_kj values[d][j] / sum_j w_kj over the `in` vertices j within 2 asin(sigma / RAD) of ref
 *                        vertex k, w = exp(-geodesic^2 / 2 sigma^2); mask [Vin] or NULL; 0 where nothing is in range.
 * msm_histogram_match  : multivariate_histogram_normalization (src/reg_tools.cpp:750-810) over MISCMATHS::Histogram ([EXT] A10):
 *                        every dimension of `data` [D][V] is matched IN PLACE to the same dimension of data_ref [D][Vref];
 *                        excl / excl_ref: NULL or [rows][V] — row d when rows > d, else row 0 (:771-776); bins = 256 in the reference.
 * msm_variance_normalise: featurespace::variance_normalise (src/featurespace.cc:67-108), in place, mask [V] or NULL.
 * msm_create_exclusion : newresampler::create_exclusion ([EXT] A11, call site src/featurespace.cc:42): mask_out[V]. */
int msm_smooth_data(msm_ctx* ctx, const double* in_xyz, int Vin, const double* values, int D, const double* mask, const double* ref_xyz, int Vref,
                    double sigma, double* out);
int msm_histogram_match(msm_ctx* ctx, double* data, int V, const double* data_ref, int Vref, int D, const double* excl, int excl_rows,
                        const double* excl_ref, int excl_ref_rows, int bins);
int msm_variance_normalise(msm_ctx* ctx, double* data, int D, int V, const double* mask);
int msm_create_exclusion(msm_ctx* ctx, const double* data, int D, int V, float lower, float upper, double* mask_out);

/* evaluateTotalCostSum (src/DiscreteCostFunction.cpp:41-70) from the device tables of the last *_table calls:
 * sums[0..2] = unary, pairwise, triplet; requires msm_unary_table(_dev) to have run. */
int msm_total_cost(msm_ctx* ctx, const int* labeling, double sums[3]);

/* Introspection for bench/tests: number of kernels this library launched so far, and per-kernel timing of the
 * last call of the named kernel family measured with CUDA events on the ctx stream (ms, <0 if none). */
long long msm_launch_count(const msm_ctx* ctx);
/* mode 0 off; 1 synchronous (msm_last_kernel_ms gives the last duration of a family: "patches", "label_rot", "unary"
 * (setup + main), "unary_main", "triplet_table", "triplet_moves", "triplet_queries", "pair_table", "group_pair",
 * "locate"); 2 asynchronous: event pairs are recorded without host syncs and summed by msm_timing_collect. */
int msm_set_timing(msm_ctx* ctx, int mode);
double msm_last_kernel_ms(const msm_ctx* ctx, const char* family);
int msm_timing_collect(msm_ctx* ctx, const char* family, double* total_ms, long long* count);

#ifdef __cplusplus
}
#endif
#endif /* MSM_B200_H */
