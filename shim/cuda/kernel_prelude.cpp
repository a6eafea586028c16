// shim/cuda/kernel_prelude.cpp — see CudaKernelPrelude.h.  TEST / BUILD INFRASTRUCTURE.
//
// Restates what OpenMM 7.3/7.4's CudaContext (the un-vendored dependency the reference builds against; its CMake pins no
// version, README.md:41-44 names 7.3/7.4) does in its constructor (compilationDefines) and in createModule.  Worth knowing
// when reading the reference's kernels: outside double precision SQRT is sqrtf and RECIP(x) is 1.0f/(x) even where the
// operand is `mixed` (= double in mixed mode), so applyHardWallConstraints' bond length (drudeTGNH.cu:488) has float
// precision in mixed mode while RECIP(velocity.w) is an exact double division; and the TGNH sources pass "" as
// optimization flags (CudaDrudeTGNHKernels.cpp:269), which switches OpenMM's default --use_fast_math off.
#include <nvrtc.h>

#include <sstream>

#include "CudaKernelPrelude.h"

namespace OpenMM {

std::map<std::string, std::string> shimCompilationDefines(bool useDoublePrecision, bool useMixedPrecision) {
    std::map<std::string, std::string> c;
    if (useDoublePrecision) {
        c["USE_DOUBLE_PRECISION"] = "1";
        c["make_real2"] = "make_double2"; c["make_real3"] = "make_double3"; c["make_real4"] = "make_double4";
        c["make_mixed2"] = "make_double2"; c["make_mixed3"] = "make_double3"; c["make_mixed4"] = "make_double4";
    } else if (useMixedPrecision) {
        c["USE_MIXED_PRECISION"] = "1";
        c["make_real2"] = "make_float2"; c["make_real3"] = "make_float3"; c["make_real4"] = "make_float4";
        c["make_mixed2"] = "make_double2"; c["make_mixed3"] = "make_double3"; c["make_mixed4"] = "make_double4";
    } else {
        c["make_real2"] = "make_float2"; c["make_real3"] = "make_float3"; c["make_real4"] = "make_float4";
        c["make_mixed2"] = "make_float2"; c["make_mixed3"] = "make_float3"; c["make_mixed4"] = "make_float4";
    }
    const bool d = useDoublePrecision;
    c["SQRT"] = d ? "sqrt" : "sqrtf";
    c["RSQRT"] = d ? "rsqrt" : "rsqrtf";
    c["RECIP"] = d ? "1.0/" : "1.0f/";
    c["EXP"] = d ? "exp" : "expf";
    c["LOG"] = d ? "log" : "logf";
    c["POW"] = d ? "pow" : "powf";
    c["COS"] = d ? "cos" : "cosf";
    c["SIN"] = d ? "sin" : "sinf";
    c["TAN"] = d ? "tan" : "tanf";
    c["ACOS"] = d ? "acos" : "acosf";
    c["ASIN"] = d ? "asin" : "asinf";
    c["ATAN"] = d ? "atan" : "atanf";
    c["ERF"] = d ? "erf" : "erff";
    c["ERFC"] = d ? "erfc" : "erfcf";
    c["SYNC_WARPS"] = "__syncwarp();";
    c["SHFL(var, srcLane)"] = "__shfl_sync(0xffffffff, var, srcLane);";
    c["BALLOT(var)"] = "__ballot_sync(0xffffffff, var);";
    return c;
}

std::string shimBuildKernelSource(bool useDoublePrecision, bool useMixedPrecision, const std::map<std::string, std::string>& compilationDefines,
                                  const std::string& source, const std::map<std::string, std::string>& defines, const std::string& options) {
    std::stringstream src;
    if (!options.empty()) src << "// Compilation Options: " << options << "\n\n";
    for (std::map<std::string, std::string>::const_iterator it = compilationDefines.begin(); it != compilationDefines.end(); ++it) {
        if (defines.find(it->first) == defines.end()) {
            src << "#define " << it->first;
            if (!it->second.empty()) src << " " << it->second;
            src << "\n";
        }
    }
    src << "\n";
    const char* r = useDoublePrecision ? "double" : "float";
    const char* m = (useDoublePrecision || useMixedPrecision) ? "double" : "float";
    src << "typedef " << r << " real;\ntypedef " << r << "2 real2;\ntypedef " << r << "3 real3;\ntypedef " << r << "4 real4;\n";
    src << "typedef " << m << " mixed;\ntypedef " << m << "2 mixed2;\ntypedef " << m << "3 mixed3;\ntypedef " << m << "4 mixed4;\n";
    src << "typedef unsigned int tileflags;\n";
    for (std::map<std::string, std::string>::const_iterator it = defines.begin(); it != defines.end(); ++it) {
        src << "#define " << it->first;
        if (!it->second.empty()) src << " " << it->second;
        src << "\n";
    }
    if (!defines.empty()) src << "\n";
    src << source << "\n";
    return src.str();
}

bool shimNvrtcCompile(const std::string& source, const std::string& options, std::vector<char>& cubin, std::string& log) {
    nvrtcProgram prog;
    if (nvrtcCreateProgram(&prog, source.c_str(), "openmm_kernel.cu", 0, NULL, NULL) != NVRTC_SUCCESS) {
        log = "nvrtcCreateProgram failed";
        return false;
    }
    std::vector<const char*> opts;
    opts.push_back("--gpu-architecture=sm_100a");
    if (options.find("--use_fast_math") != std::string::npos) opts.push_back("--use_fast_math");
    const nvrtcResult res = nvrtcCompileProgram(prog, (int)opts.size(), &opts[0]);
    size_t logSize = 0;
    nvrtcGetProgramLogSize(prog, &logSize);
    log.assign(logSize, ' ');
    if (logSize) nvrtcGetProgramLog(prog, &log[0]);
    if (res != NVRTC_SUCCESS) {
        nvrtcDestroyProgram(&prog);
        return false;
    }
    size_t cubinSize = 0;
    nvrtcGetCUBINSize(prog, &cubinSize);
    cubin.resize(cubinSize);
    nvrtcGetCUBIN(prog, &cubin[0]);
    nvrtcDestroyProgram(&prog);
    return true;
}

}  // namespace OpenMM
