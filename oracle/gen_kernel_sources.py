#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Does what the reference's platforms/cuda/EncodeCUDAFiles.cmake does at configure time: turns every
platforms/cuda/src/kernels/*.cu of the reference into a `static const std::string` member of CudaDrudeTGNHKernelSources
(the .h.in / .cpp.in templates next to the reference's sources), so that the reference's CudaDrudeTGNHKernels.cpp compiles
unmodified.  The reference's files are read where they lie; the two generated files go to the output directory
(oracle/_refcuda/, git-ignored) and nowhere else.

    python gen_kernel_sources.py <reference root> <output dir>
"""
import glob
import os
import re
import sys

ref, out = sys.argv[1], sys.argv[2]
src = os.path.join(ref, "platforms", "cuda", "src")
decls, defs = [], []
for path in sorted(glob.glob(os.path.join(src, "kernels", "*.cu"))):
    name = os.path.basename(path)[:-3]
    text = open(path).read()
    lines = text.split("\n")
    body = "\n".join('"' + ln.replace("\\", "\\\\").replace('"', '\\"') + '\\n"' for ln in lines)
    decls.append(f"static const std::string {name};")
    defs.append(f"const string CudaDrudeTGNHKernelSources::{name} = {body};")


def strip_license(t):
    return re.sub(r"/\*.*?\*/", "", t, count=1, flags=re.S)


h = strip_license(open(os.path.join(src, "CudaDrudeTGNHKernelSources.h.in")).read()).replace("@CUDA_FILE_DECLARATIONS@", "\n".join(decls))
c = strip_license(open(os.path.join(src, "CudaDrudeTGNHKernelSources.cpp.in")).read()) + "\n" + "\n".join(defs) + "\n"
os.makedirs(out, exist_ok=True)
open(os.path.join(out, "CudaDrudeTGNHKernelSources.h"), "w").write(h)
open(os.path.join(out, "CudaDrudeTGNHKernelSources.cpp"), "w").write(c)
