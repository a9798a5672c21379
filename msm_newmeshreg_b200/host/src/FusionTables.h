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
#include "DiscreteGroupModel.h"
#include "DiscreteModel.h"
#include "Fusion/Fusion.h"

namespace newmeshreg {

// The binary (2-label) MRF that ELCReduce::PBF::convert fills and FPD::FastPD solves — the role of DiscreteModelDummy /
// DummyCostFunction (include/Fusion/Fusion.h:15-116), over flat arrays instead of std::map<int, std::vector<double>>.
// FastPD asks for a pairwise cost through the PAIR macro ~40 times per edge and sweep (include/FastPD/FastPD.h:40); with
// the reference's dummy every one of those is a red-black-tree look-up, which is most of FastPD::run at an ico5 control
// grid (tests/fusion_phase_timing.py).  Semantics kept as they are there, including the two quirks that matter for
// identical labellings: only the FIRST unary term given for a node counts (setUnaryCost appends, convertenergies reads
// entries 0 and 1 — so the constant that convert() adds to node 0 last is dropped), and a pairwise term is the tuple
// (E00, E01, E10, E11) in the order it was added.
class FlatBinaryCostFunction : public DiscreteCostFunction {
public:
    FlatBinaryCostFunction() { m_num_labels = 2; }
    void clear() { u.clear(); seen.clear(); pe.clear(); }
    void setNodes(int n) { u.assign(2 * (size_t)n, 0.0); seen.assign(n, 0); }
    void setUnaryCost(int node, double c0, double c1) {
        if (node >= (int)seen.size()) { u.resize(2 * (size_t)(node + 1), 0.0); seen.resize(node + 1, 0); }
        if (seen[node]) return;
        seen[node] = 1; u[2 * (size_t)node] = c0; u[2 * (size_t)node + 1] = c1;
    }
    void addPairwiseCost(double E00, double E01, double E10, double E11) { pe.push_back(E00); pe.push_back(E01); pe.push_back(E10); pe.push_back(E11); }
    double computePairwiseCost(int pair, int labelA, int labelB) override { return pe[4 * (size_t)pair + 2 * labelA + labelB]; }
    double computeUnaryCost(int node, int label) override { return u[2 * (size_t)node + label]; }
    void convertenergies(int numNodes, int numPairs) {   // label-major unary table, the layout FastPD reads (and overwrites)
        m_num_nodes = numNodes; m_num_pairs = numPairs; m_num_labels = 2;
        delete[] unarycosts;
        unarycosts = new double[2 * (size_t)numNodes];
        for (int j = 0; j < numNodes; j++) {
            const bool have = j < (int)seen.size();
            unarycosts[j] = have ? u[2 * (size_t)j] : 0.0;
            unarycosts[(size_t)numNodes + j] = have ? u[2 * (size_t)j + 1] : 0.0;
        }
    }
private:
    std::vector<double> u, pe;
    std::vector<char> seen;
};

class FlatBinaryModel : public DiscreteModel {
public:
    FlatBinaryModel() : costfct(std::make_shared<FlatBinaryCostFunction>()) { m_num_labels = 2; }
    std::shared_ptr<DiscreteCostFunction> getCostFunction() override { return costfct; }
    // the OPTIMIZER interface of ELCReduce::PBF::convert (include/ELC/ELC.h:322-340)
    void AddNode(int num) { m_num_nodes = num; costfct->setNodes(num); }
    void AddUnaryTerm(int node, double E0, double E1) { costfct->setUnaryCost(node, E0, E1); }
    void AddPairwiseTerm(int node1, int node2, double E00, double E01, double E10, double E11) {
        pair_ids.push_back(node1); pair_ids.push_back(node2);
        costfct->addPairwiseCost(E00, E01, E10, E11);
        m_num_pairs++;
    }
    void initialise() {
        initLabeling();
        costfct->convertenergies(m_num_nodes, m_num_pairs);
        delete[] pairs;
        pairs = new int[2 * (size_t)std::max(m_num_pairs, 1)];
        std::copy(pair_ids.begin(), pair_ids.end(), pairs);
    }
    void reset() { pair_ids.clear(); m_num_pairs = 0; m_num_nodes = 0; m_num_labels = 2; costfct->clear(); }
private:
    std::vector<int> pair_ids;
    std::shared_ptr<FlatBinaryCostFunction> costfct;
};

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
        const bool flat = std::getenv("MSM_FUSION_DUMMY") == nullptr;   // MSM_FUSION_DUMMY=1: the reference's std::map-backed dummy model
        auto binary = std::make_shared<FlatBinaryModel>();
        auto dummy = std::make_shared<DiscreteModelDummy>();
        // MSM_FUSION_TIMING=1: where the optimiser phase goes (seconds summed over the 2 L proposal steps)
        const bool timing = std::getenv("MSM_FUSION_TIMING") != nullptr;
        double tsec[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        auto tick = [&](int slot, double& t0) { if (timing) { const double t1 = omp_get_wtime(); tsec[slot] += t1 - t0; t0 = t1; } };
        const int NUM_SWEEPS = 2;                                   // Fusion.h:124
        double lastEnergy = cf->totalCostBatched(labeling);
        for (int sweep = 0; sweep < NUM_SWEEPS; ++sweep)
            for (int label = 0; label < L; ++label) {
                bool any = false;                                   // Fusion.h:150,153: nothing to fuse if all nodes carry it
                for (int node = 0; node < N && !any; ++node) any = labeling[node] != label;
                if (!any) continue;
                double t0 = timing ? omp_get_wtime() : 0.0;
                ELCReduce::PBF<double> pbf;
                for (int node = 0; node < N; ++node)
                    pbf.AddUnaryTerm(node, U[(size_t)labeling[node] * N + node], U[(size_t)label * N + node]);
                tick(0, t0);
                for (int p = 0; p < P; ++p) {                       // pairwise cliques (rare with HOCR): the small [P][L][L] table
                    const int a = pairs[2 * p], b = pairs[2 * p + 1];
                    pbf.AddPairwiseTerm(a, b, cf->computePairwiseCost(p, labeling[a], labeling[b]), cf->computePairwiseCost(p, labeling[a], label),
                                        cf->computePairwiseCost(p, label, labeling[b]), cf->computePairwiseCost(p, label, label));
                }
                if (T > 0) {
                    cf->computeTripletMoves(labeling, label, moves.data());   // one launch, T*8 costs (fp32 = the look-up tables' precision)
                    tick(1, t0);
                    for (int t = 0; t < T; ++t) {
                        int ids[3] = {triplets[3 * t], triplets[3 * t + 1], triplets[3 * t + 2]};
                        double buf[8];
                        for (int m = 0; m < 8; m++) buf[m] = (double)moves[(size_t)t * 8 + m];
                        pbf.AddHigherTerm(3, ids, buf);
                    }
                    tick(2, t0);
                }
                binary->reset(); dummy->reset();
                ELCReduce::PBF<double> qpbf;
                pbf.toQuadratic(qpbf, pbf.maxID() + 1);             // HOCR reduction (include/ELC/ELC.h)
                tick(3, t0);
                if (flat) qpbf.convert(*binary, qpbf.maxID() + 1);
                else qpbf.convert(*dummy, qpbf.maxID() + 1);
                pbf.clear();
                qpbf.clear();
                if (flat) binary->initialise();
                else dummy->initialise();
                tick(4, t0);
                std::shared_ptr<DiscreteModel> bm = flat ? std::static_pointer_cast<DiscreteModel>(binary) : std::static_pointer_cast<DiscreteModel>(dummy);
                int* take = bm->getLabeling();
                FPD::FastPD opt(bm, 100);
                tick(5, t0);
                const double newEnergy = opt.run();
                opt.getLabeling(take);
                tick(6, t0);
                int changed = 0;
                for (int node = 0; node < N; ++node)
                    if (labeling[node] != label && take[node] == 1) { labeling[node] = label; changed++; }
                if (verbose) {
                    std::cout << "  LAB " << label << ":\t" << lastEnergy << " -> " << newEnergy << " / " << 100.0 * changed / N << "% CHN" << std::endl;
                    lastEnergy = newEnergy;
                }
            }
        if (timing)
            std::cerr << "FusionTables timing [s]: unary terms " << tsec[0] << ", move tables (GPU + read-back) " << tsec[1] << ", AddHigherTerm " << tsec[2]
                      << ", toQuadratic " << tsec[3] << ", convert + DiscreteModelDummy::initialise " << tsec[4] << ", FastPD ctor " << tsec[5]
                      << ", FastPD::run " << tsec[6] << std::endl;
        return cf->totalCostBatched(labeling);
    }

    // Groupwise registration (src/group_mesh_registration.cpp:74: Fusion::optimize on a DiscreteGroupModel): no unary term;
    // per proposal label ONE batched launch for the 4 costs of every inter-subject pair (msm_group_pair_moves) and one for the
    // 8 costs of every triplet — instead of 4 P + 8 T per-element virtuals.  Sweeps, label order, term order (pairs before
    // triplets) and the acceptance rule follow include/Fusion/Fusion.h:126-243.
    static double optimize_group(const std::shared_ptr<DiscreteGroupModel>& energy, bool verbose, int /*numthreads*/ = 1) {
        auto cf = energy->getGroupCostFunction();
        const int N = energy->getNumNodes(), L = energy->getNumLabels(), T = energy->getNumTriplets(), P = energy->getNumPairs();
        const int* pairs = energy->getPairs();
        const int* triplets = energy->getTriplets();
        int* labeling = energy->getLabeling();
        std::vector<double> pmoves((size_t)std::max(P, 1) * 4);
        std::vector<float> moves((size_t)std::max(T, 1) * 8);
        auto binary = std::make_shared<FlatBinaryModel>();
        for (int sweep = 0; sweep < 2; ++sweep)
            for (int label = 0; label < L; ++label) {
                bool any = false;
                for (int node = 0; node < N && !any; ++node) any = labeling[node] != label;
                if (!any) continue;
                ELCReduce::PBF<double> pbf;
                for (int node = 0; node < N; ++node) pbf.AddUnaryTerm(node, 0.0, 0.0);       // base computeUnaryCost returns 0 (src/DiscreteCostFunction.h:34)
                if (P > 0) {
                    cf->computePairwiseMoves(labeling, label, pmoves.data());
                    for (int p = 0; p < P; ++p)
                        pbf.AddPairwiseTerm(pairs[2 * p], pairs[2 * p + 1], pmoves[4 * (size_t)p], pmoves[4 * (size_t)p + 1], pmoves[4 * (size_t)p + 2],
                                            pmoves[4 * (size_t)p + 3]);
                }
                if (T > 0) {
                    cf->computeTripletMovesGroup(labeling, label, moves.data());
                    for (int t = 0; t < T; ++t) {
                        int ids[3] = {triplets[3 * t], triplets[3 * t + 1], triplets[3 * t + 2]};
                        double buf[8];
                        for (int m = 0; m < 8; m++) buf[m] = (double)moves[(size_t)t * 8 + m];
                        pbf.AddHigherTerm(3, ids, buf);
                    }
                }
                binary->reset();
                ELCReduce::PBF<double> qpbf;
                pbf.toQuadratic(qpbf, pbf.maxID() + 1);
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
                if (verbose) std::cout << "  LAB " << label << ":\t" << newEnergy << " / " << 100.0 * changed / N << "% CHN" << std::endl;
            }
        return cf->totalCostBatchedGroup(labeling);
    }
};

}  // namespace newmeshreg
