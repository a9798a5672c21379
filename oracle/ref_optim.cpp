// ORACLE / TEST INFRASTRUCTURE — oracle/_ref/libref_optim.so
// The reference's OWN optimisers (include/FastPD/FastPD.h, include/ELC/ELC.h, include/Fusion/Fusion.h), compiled
// UNMODIFIED from where they lie under /root/reference/include (nothing is copied into this repository; licences:
// licenses/FastPD_license.txt, licenses/ELC_permission.txt), against the FSL-free mirror headers of
// msm_newmeshreg_b200/host/src.  Used by the tests to check that GPU-produced and oracle-produced cost tables lead to
// identical MRF labellings.  Their `#include "../src/DiscreteModel.h"` resolves through -Ihost/inc (inc/../src).
#define WITH_REF_OPTIM
#include "../msm_newmeshreg_b200/host/harness.cpp"
