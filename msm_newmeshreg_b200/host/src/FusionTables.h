// FusionTables.h — table-aware fusion-move driver.
//
// The reference's Fusion::optimize (include/Fusion/Fusion.h:119-243) asks the model for every binary-move cost through
// per-element virtuals: per proposal label 2N unary + 8T triplet calls, i.e. 10^6-10^7 calls per outer iteration, and to
// serve them from look-ups the GPU-backed cost function has to materialise the whole T*L^3 triplet table (560 MB at an ico5
// control grid).  This driver takes the same decisions from BATCHED tables instead: the unary costs are two rows of the
// [L][N] table, the eight move costs of every triplet come from ONE kernel launch per proposal label
// (computeTripletMoves, same bit order as Fusion.h:187-194).  No T*L^3 table is ever built or copied.
//
// What is NOT restated: the optimisation itself.  The higher-order clique reduction (ELCReduce::PBF, include/ELC/ELC.h),
// the conversion to a pairwise MRF (DiscreteModelDummy, Fusion.h:66-116) and the primal-dual solver (FPD::FastPD) are the
// reference's own vendored code, used unchanged — which is why this header only compiles where those are on the include
// path (in an FSL tree next to Fusion.h; here: oracle/_ref builds).  Sweeps, label order, acceptance rule and the returned
// energy follow Fusion.h:126-243, so for identical tables the labelling is identical (tests/test_gpu_model.py).
#pragma once
#include "DiscreteModel.h"
#include "Fusion/Fusion.h"

namespace newmeshreg {

class FusionTables {
public:
    // `energy` must be a NonLinearSRegDiscreteModel over the GPU-backed cost function, set up for this outer iteration
    static double optimize(const std::shared_ptr<NonLinearSRegDiscreteModel>& energy, bool verbose, int /*numthreads*/ = 1) {
        auto cf = energy->getNonLinearCostFunction();
        const int N = energy->getNumNodes(), L = energy->getNumLabels(), T = energy->getNumTriplets(), P = energy->getNumPairs();
        const int* pairs = energy->getPairs();
        const int* triplets = energy->getTriplets();
        int* labeling = energy->getLabeling();
        const std::vector<double>& U = cf->unaryTable();            // [L][N]
        if (U.size() != (size_t)L * N)
            throw MeshregException("FusionTables::optimize: the model has no [L][N] unary table (group models go through FusionTables::optimize_group)");
        std::vector<float> moves((size_t)std::max(T, 1) * 8);       // [T][8] of the current proposal label
        auto binary = std::make_shared<DiscreteModelDummy>();
        const int NUM_SWEEPS = 2;                                   // Fusion.h:124
        double lastEnergy = cf->totalCostBatched(labeling);
        for (int sweep = 0; sweep < NUM_SWEEPS; ++sweep)
            for (int label = 0; label < L; ++label) {
                bool any = false;                                   // Fusion.h:150,153: nothing to fuse if all nodes carry it
                for (int node = 0; node < N && !any; ++node) any = labeling[node] != label;
                if (!any) continue;
                ELCReduce::PBF<double> pbf;
                for (int node = 0; node < N; ++node)
                    pbf.AddUnaryTerm(node, U[(size_t)labeling[node] * N + node], U[(size_t)label * N + node]);
                for (int p = 0; p < P; ++p) {                       // pairwise cliques (rare with HOCR): the small [P][L][L] table
                    const int a = pairs[2 * p], b = pairs[2 * p + 1];
                    pbf.AddPairwiseTerm(a, b, cf->computePairwiseCost(p, labeling[a], labeling[b]), cf->computePairwiseCost(p, labeling[a], label),
                                        cf->computePairwiseCost(p, label, labeling[b]), cf->computePairwiseCost(p, label, label));
                }
                if (T > 0) {
                    cf->computeTripletMoves(labeling, label, moves.data());   // one launch, T*8 costs (fp32 = the look-up tables' precision)
                    for (int t = 0; t < T; ++t) {
                        int ids[3] = {triplets[3 * t], triplets[3 * t + 1], triplets[3 * t + 2]};
                        double buf[8];
                        for (int m = 0; m < 8; m++) buf[m] = (double)moves[(size_t)t * 8 + m];
                        pbf.AddHigherTerm(3, ids, buf);
                    }
                }
                binary->reset();
                ELCReduce::PBF<double> qpbf;
                pbf.toQuadratic(qpbf, pbf.maxID() + 1);             // HOCR reduction (include/ELC/ELC.h)
                qpbf.convert(*binary, qpbf.maxID() + 1);
                pbf.clear();
                qpbf.clear();
                binary->initialise();
                int* take = binary->getLabeling();
                FPD::FastPD opt(binary, 100);
                const double newEnergy = opt.run();
                opt.getLabeling(take);
                int changed = 0;
                for (int node = 0; node < N; ++node)
                    if (labeling[node] != label && take[node] == 1) { labeling[node] = label; changed++; }
                if (verbose) {
                    std::cout << "  LAB " << label << ":\t" << lastEnergy << " -> " << newEnergy << " / " << 100.0 * changed / N << "% CHN" << std::endl;
                    lastEnergy = newEnergy;
                }
            }
        return cf->totalCostBatched(labeling);
    }
};

}  // namespace newmeshreg
