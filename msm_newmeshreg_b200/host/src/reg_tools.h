// reg_tools.h — the two boundary types of the reference's src/reg_tools.h that the cost-function / model interfaces
// use: the per-level parameter map `myparam` (src/reg_tools.h:12-13; std::variant replaces boost::variant, with a
// `boost::get` shim so code written against the reference compiles) and MeshregException (src/reg_tools.h:15-23).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <exception>
#include <iostream>
#include <map>
#include <string>
#include <variant>

#include "newresampler_shim.h"

namespace boost {  // spelling used throughout the reference: boost::get<T>(it->second)
template <typename T, typename... Ts>
const T& get(const std::variant<Ts...>& v) { return std::get<T>(v); }
}  // namespace boost

namespace newmeshreg {

// Diagnostics: with MSM_HOST_TIMING=1 in the environment every HostTimer scope prints its wall time to stderr.
struct HostTimer {
    const char* name; double t0 = 0; bool on;
    static bool enabled() { static const bool e = [] { const char* v = std::getenv("MSM_HOST_TIMING"); return v && std::atoi(v) != 0; }(); return e; }
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    explicit HostTimer(const char* n) : name(n), on(enabled()) { if (on) t0 = now(); }
    ~HostTimer() { if (on) std::fprintf(stderr, "[msm host] %-28s %8.3f ms\n", name, now() - t0); }
};

typedef std::map<std::string, std::variant<int, std::string, double, float, bool>> myparam;
typedef std::pair<std::string, std::variant<int, std::string, double, float, bool>> parameterPair;

class MeshregException : public std::exception {
public:
    const char* errmesg;
    explicit MeshregException(const char* msg) : errmesg(msg) {}
    const char* what() const noexcept override {  // the reference's what() also prints (src/reg_tools.cpp:5-8)
        std::cout << errmesg << std::endl;
        return errmesg;
    }
};

// Default parameter set of one resolution level (src/mesh_registration.cpp:491-556,689-715,815-847), FSL-free.
inline myparam default_parameters() {
    myparam P;
    P["lambda"] = 0.1f;  P["exponent"] = 2.0f;  P["weight"] = false;  P["anorm"] = false;
    P["shearmodulus"] = 0.4f;  P["bulkmodulus"] = 1.6f;  P["kexponent"] = 2.0f;  P["range"] = 1.0f;
    P["simmeasure"] = 2;  P["verbosity"] = false;  P["regularisermode"] = 1;  P["numthreads"] = 1;
    P["dOPT"] = std::string("FastPD");  P["SGres"] = 4;  P["CPres"] = 2;  P["multivariate"] = false;
    P["outdir"] = std::string("");  P["TriLikelihood"] = false;  P["rescalelabels"] = false;  P["labeldist"] = 0.5f;
    P["iters"] = 3;
    return P;
}

}  // namespace newmeshreg
