// TEST INFRASTRUCTURE — compiles the reference's kernel strings (vectorOps + drudeTGNH, exactly what
// CudaDrudeTGNHKernels.cpp:269 hands to cu.createModule) behind OpenMM's prelude for sm_100a with NVRTC, in the three precision
// modes.  Needs no GPU: the CPU test suite runs it to know that oracle/_refcuda/librefcuda.so will be able to JIT its kernels.
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "CudaDrudeTGNHKernelSources.h"
#include "CudaKernelPrelude.h"

using namespace OpenMM;

int main() {
    std::map<std::string, std::string> defines;          // CudaDrudeTGNHKernels.cpp:258-265, sizes of a small system
    defines["NUM_ATOMS"] = "2500"; defines["PADDED_NUM_ATOMS"] = "2528"; defines["NUM_NORMAL_PARTICLES"] = "1476"; defines["NUM_RESIDUES"] = "512";
    defines["NUM_TEMP_GROUPS"] = "2"; defines["NUM_PAIRS"] = "512"; defines["WORK_GROUP_SIZE"] = "64";
    const char* names[3] = {"single", "mixed", "double"};
    for (int mode = 0; mode < 3; mode++) {
        const bool dbl = mode == 2, mixed = mode == 1;
        const std::string src = shimBuildKernelSource(dbl, mixed, shimCompilationDefines(dbl, mixed),
                                                      CudaDrudeTGNHKernelSources::vectorOps + CudaDrudeTGNHKernelSources::drudeTGNH, defines, "");
        std::vector<char> cubin;
        std::string log;
        if (!shimNvrtcCompile(src, "", cubin, log)) {
            printf("%s: FAILED\n%s\n", names[mode], log.c_str());
            return 1;
        }
        printf("%s: ok, cubin %zu bytes\n", names[mode], cubin.size());
    }
    printf("Done\n");
    return 0;
}
