// shim/cuda/shim_cuda.cpp — implementation of the CUDA-platform stand-in (CudaContext, CudaArray, CudaIntegrationUtilities,
// CudaPlatform).  TEST / BUILD INFRASTRUCTURE, see CudaContext.h.  CUDA driver API + NVRTC only (no nvcc-compiled code), so
// it builds with g++ here and runs wherever libcuda is present.
//
// The kernel prelude written by createModule restates what OpenMM 7.3/7.4's CudaContext (the un-vendored dependency the
// reference builds against; its CMake pins no version, README.md:41-44 names 7.3/7.4) puts in front of every kernel source:
// the `compilationDefines` of the precision mode, the real/mixed typedefs, `tileflags`, then the caller's defines.
#include <cuda.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <sstream>

#include "CudaContext.h"
#include "CudaKernelPrelude.h"
#include "openmm/System.h"

namespace OpenMM {

static void check(CUresult r, const char* what) {
    if (r == CUDA_SUCCESS) return;
    const char* msg = NULL;
    cuGetErrorString(r, &msg);
    throw OpenMMException(std::string("CUDA shim: ") + what + " failed: " + (msg ? msg : "unknown error"));
}

// ------------------------------------------------------------------------------------------------ CudaArray
CudaArray::CudaArray(CudaContext& context, int size, int elementSize, const std::string& name)
    : context(&context), pointer(0), size(size), elementSize(elementSize), name(name) {
    context.setAsCurrent();
    const size_t bytes = (size_t)size * elementSize;
    check(cuMemAlloc(&pointer, bytes ? bytes : 16), ("cuMemAlloc of " + name).c_str());
    // OpenMM leaves new arrays uninitialised.  The reference's computeNormalizedKineticEnergies clears only the rows of
    // the threads it launches while sumNormalizedKineticEnergies adds up the whole buffer (drudeTGNH.cu:144-145, 211-214),
    // i.e. it relies on fresh device memory being zero; the shim makes that explicit.
    check(cuMemsetD8(pointer, 0, bytes ? bytes : 16), "cuMemsetD8");
}
CudaArray::~CudaArray() {
    if (pointer) cuMemFree(pointer);
}
void CudaArray::upload(const void* data, bool) {
    context->setAsCurrent();
    check(cuMemcpyHtoD(pointer, data, (size_t)size * elementSize), ("upload of " + name).c_str());
}
void CudaArray::download(void* data, bool) const {
    context->setAsCurrent();
    check(cuMemcpyDtoH(data, pointer, (size_t)size * elementSize), ("download of " + name).c_str());
}

// ------------------------------------------------------------------------------------------------ CudaIntegrationUtilities
CudaIntegrationUtilities::CudaIntegrationUtilities(CudaContext& context, const System& system)
    : constraintCalls(0), velocityConstraintCalls(0), virtualSiteCalls(0), context(context), posDelta(NULL), stepSize(NULL) {
    const bool dbl = context.getUseDoublePrecision() || context.getUseMixedPrecision();
    posDelta = new CudaArray(context, context.getPaddedNumAtoms(), dbl ? 32 : 16, "posDelta");
    stepSize = new CudaArray(context, 1, dbl ? 16 : 8, "stepSize");
}
CudaIntegrationUtilities::~CudaIntegrationUtilities() {
    delete posDelta;
    delete stepSize;
}
double CudaIntegrationUtilities::computeKineticEnergy(double) {
    const int n = context.getNumAtoms(), padded = context.getPaddedNumAtoms();
    double ke = 0.0;
    if (context.getUseDoublePrecision() || context.getUseMixedPrecision()) {
        std::vector<double> v((size_t)4 * padded);
        context.getVelm().download(&v[0]);
        for (int i = 0; i < n; i++)
            if (v[4 * i + 3] != 0) ke += (v[4 * i] * v[4 * i] + v[4 * i + 1] * v[4 * i + 1] + v[4 * i + 2] * v[4 * i + 2]) / v[4 * i + 3];
    } else {
        std::vector<float> v((size_t)4 * padded);
        context.getVelm().download(&v[0]);
        for (int i = 0; i < n; i++)
            if (v[4 * i + 3] != 0) ke += ((double)v[4 * i] * v[4 * i] + (double)v[4 * i + 1] * v[4 * i + 1] + (double)v[4 * i + 2] * v[4 * i + 2]) / v[4 * i + 3];
    }
    return 0.5 * ke;
}

// ------------------------------------------------------------------------------------------------ CudaContext
CudaContext::CudaContext(const System& system, const std::string& precision, CudaPlatform::PlatformData& platformData)
    : shimReorderInterval(0), shimReorderCount(0), shimReorderAttempts(0), shimKernelLaunches(0), system(system), platformData(platformData), cuContext(NULL), device(0),
      deviceIndex(0), numAtoms(system.getNumParticles()), paddedNumAtoms(0), numThreadBlocks(0), stepCount(0), useDoublePrecision(precision == "double"),
      useMixedPrecision(precision == "mixed"), atomsWereReordered(false), time(0.0), posq(NULL), posqCorrection(NULL), velm(NULL), force(NULL),
      atomIndexDevice(NULL), integration(NULL) {
    if (precision != "single" && precision != "mixed" && precision != "double")
        throw OpenMMException("Illegal value for Precision: " + precision);
    check(cuInit(0), "cuInit");
    int count = 0;
    check(cuDeviceGetCount(&count), "cuDeviceGetCount");
    if (count == 0) throw OpenMMException("No compatible CUDA device is available");
    // the primary context of the runtime's current device, so that runtime-API libraries (libtgnh) share it
    int current = 0;
    if (CUcontext cur = NULL; cuCtxGetCurrent(&cur) == CUDA_SUCCESS && cur != NULL) {
        CUdevice d;
        if (cuCtxGetDevice(&d) == CUDA_SUCCESS) current = (int)d;
    }
    deviceIndex = current;
    check(cuDeviceGet(&device, deviceIndex), "cuDeviceGet");
    check(cuDevicePrimaryCtxRetain(&cuContext, device), "cuDevicePrimaryCtxRetain");
    check(cuCtxSetCurrent(cuContext), "cuCtxSetCurrent");
    int multiprocessors = 0;
    check(cuDeviceGetAttribute(&multiprocessors, CU_DEVICE_ATTRIBUTE_MULTIPROCESSOR_COUNT, device), "cuDeviceGetAttribute");
    numThreadBlocks = 4 * multiprocessors;                         // OpenMM 7.x: numThreadBlocksPerComputeUnit = 4 on CC >= 6
    paddedNumAtoms = TileSize * ((numAtoms + TileSize - 1) / TileSize);
    const bool dblVel = useDoublePrecision || useMixedPrecision;
    posq = new CudaArray(*this, paddedNumAtoms, useDoublePrecision ? 32 : 16, "posq");
    posqCorrection = new CudaArray(*this, paddedNumAtoms, 16, "posqCorrection");      // OpenMM allocates it in mixed mode only; unused otherwise
    velm = new CudaArray(*this, paddedNumAtoms, dblVel ? 32 : 16, "velm");
    force = new CudaArray(*this, 3 * paddedNumAtoms, 8, "force");
    atomIndexDevice = new CudaArray(*this, paddedNumAtoms, 4, "atomIndex");
    atomIndex.resize(paddedNumAtoms);
    for (int i = 0; i < paddedNumAtoms; i++) atomIndex[i] = i;
    atomIndexDevice->upload(&atomIndex[0]);
    integration = new CudaIntegrationUtilities(*this, system);

    compilationDefines = shimCompilationDefines(useDoublePrecision, useMixedPrecision);
}

CudaContext::~CudaContext() {
    setAsCurrent();
    for (size_t i = 0; i < forces.size(); i++) delete forces[i];
    for (size_t i = 0; i < modules.size(); i++) cuModuleUnload(modules[i]);
    delete integration;
    delete posq; delete posqCorrection; delete velm; delete force; delete atomIndexDevice;
    cuDevicePrimaryCtxRelease(device);
}

void CudaContext::setAsCurrent() {
    if (cuContext != NULL) cuCtxSetCurrent(cuContext);
}

CUmodule CudaContext::createModule(const std::string source, const char* optimizationFlags) {
    return createModule(source, std::map<std::string, std::string>(), optimizationFlags);
}

CUmodule CudaContext::createModule(const std::string source, const std::map<std::string, std::string>& defines, const char* optimizationFlags) {
    const std::string options = optimizationFlags == NULL ? "--use_fast_math" : std::string(optimizationFlags);
    shimLastSource = shimBuildKernelSource(useDoublePrecision, useMixedPrecision, compilationDefines, source, defines, options);
    std::vector<char> cubin;
    std::string log;
    if (!shimNvrtcCompile(shimLastSource, options, cubin, log)) throw OpenMMException("Error compiling kernel: " + log);
    setAsCurrent();
    CUmodule module;
    check(cuModuleLoadData(&module, &cubin[0]), "cuModuleLoadData");
    modules.push_back(module);
    return module;
}

CUfunction CudaContext::getKernel(CUmodule& module, const std::string& name) {
    CUfunction function;
    check(cuModuleGetFunction(&function, module, name.c_str()), ("cuModuleGetFunction(" + name + ")").c_str());
    return function;
}

void CudaContext::executeKernel(CUfunction kernel, void** arguments, int threads, int blockSize, unsigned int sharedSize) {
    if (blockSize == -1) blockSize = ThreadBlockSize;
    int gridSize = (threads + blockSize - 1) / blockSize;
    if (gridSize > numThreadBlocks) gridSize = numThreadBlocks;
    if (gridSize < 1) gridSize = 1;
    check(cuLaunchKernel(kernel, gridSize, 1, 1, blockSize, 1, 1, sharedSize, getCurrentStream(), arguments, NULL), "cuLaunchKernel");
    shimKernelLaunches++;
}

void CudaContext::clearBuffer(CudaArray& array) {
    setAsCurrent();
    check(cuMemsetD8(array.getDevicePointer(), 0, (size_t)array.getSize() * array.getElementSize()), "cuMemsetD8");
}

std::string CudaContext::intToString(int value) const {
    std::stringstream s;
    s << value;
    return s.str();
}
std::string CudaContext::doubleToString(double value) const {
    std::stringstream s;
    s.precision(useDoublePrecision ? 16 : 8);
    s << std::scientific << value;
    if (!useDoublePrecision) s << "f";
    return s.str();
}

// Swap neighbouring interchangeable molecules (same particle count, every pair of corresponding particles identical for every
// registered CudaForceInfo and of equal mass).  posq, posqCorrection, velm and atomIndex move; the force buffer does NOT (OpenMM's
// reorderAtoms leaves forces stale too, which is why the reference recomputes them, CudaDrudeTGNHKernels.cpp:344-347).
void CudaContext::reorderAtoms() {
    atomsWereReordered = false;
    if (shimReorderInterval <= 0 || stepCount % shimReorderInterval != 0) return;
    const std::vector<std::vector<int> > molecules = platformData.context->getMolecules();
    // slot ranges of the molecules in the current order: molecules are contiguous index ranges and only whole, equal-sized,
    // identical molecules ever trade places, so slot ranges == original index ranges
    std::vector<std::pair<int, int> > swaps;
    for (size_t k = (size_t)(shimReorderAttempts & 1); k + 1 < molecules.size(); k += 2) {
        const std::vector<int>&a = molecules[k], &b = molecules[k + 1];
        if (a.size() != b.size()) continue;
        bool same = true;
        for (size_t j = 0; j < a.size() && same; j++) {
            const int pa = atomIndex[a[j]], pb = atomIndex[b[j]];       // the particles that currently sit in these slots
            if (a[j] != a[0] + (int)j || b[j] != b[0] + (int)j) same = false;
            if (system.getParticleMass(pa) != system.getParticleMass(pb)) same = false;
            for (size_t f = 0; f < forces.size() && same; f++) same = forces[f]->areParticlesIdentical(pa, pb);
        }
        if (same) swaps.push_back(std::make_pair(a[0], b[0] | ((int)a.size() << 24)));
    }
    shimReorderAttempts++;
    if (swaps.empty()) return;
    shimReorderCount++;
    auto permute = [&](CudaArray& arr) {
        const int es = arr.getElementSize();
        std::vector<char> h((size_t)arr.getSize() * es), t(es);
        arr.download(&h[0]);
        for (size_t s = 0; s < swaps.size(); s++) {
            const int a0 = swaps[s].first, b0 = swaps[s].second & 0xffffff, len = swaps[s].second >> 24;
            for (int j = 0; j < len; j++) {
                memcpy(&t[0], &h[(size_t)(a0 + j) * es], es);
                memcpy(&h[(size_t)(a0 + j) * es], &h[(size_t)(b0 + j) * es], es);
                memcpy(&h[(size_t)(b0 + j) * es], &t[0], es);
            }
        }
        arr.upload(&h[0]);
    };
    permute(*posq); permute(*posqCorrection); permute(*velm);
    for (size_t s = 0; s < swaps.size(); s++) {
        const int a0 = swaps[s].first, b0 = swaps[s].second & 0xffffff, len = swaps[s].second >> 24;
        for (int j = 0; j < len; j++) std::swap(atomIndex[a0 + j], atomIndex[b0 + j]);
    }
    atomIndexDevice->upload(&atomIndex[0]);
    atomsWereReordered = true;
}

// ------------------------------------------------------------------------------------------------ CudaPlatform
namespace {
// host copies <-> device arrays, in slot order (atomIndex) and OpenMM's layouts
void uploadState(ContextImpl& ctx, CudaContext& cu) {
    const int n = cu.getNumAtoms(), padded = cu.getPaddedNumAtoms();
    const bool dblVel = cu.getUseDoublePrecision() || cu.getUseMixedPrecision(), dblPos = cu.getUseDoublePrecision();
    std::vector<double> vd, xd;
    std::vector<float> vf, xf, cf((size_t)4 * padded, 0.f);
    if (dblVel) vd.assign((size_t)4 * padded, 0.0); else vf.assign((size_t)4 * padded, 0.f);
    if (dblPos) { xd.assign((size_t)4 * padded, 0.0); cu.getPosq().download(&xd[0]); } else { xf.assign((size_t)4 * padded, 0.f); cu.getPosq().download(&xf[0]); }   // keeps the charges
    for (int s = 0; s < n; s++) {
        const int p = cu.getAtomIndex()[s];
        const double mass = ctx.getSystem().getParticleMass(p);
        for (int c = 0; c < 3; c++) {
            const double x = ctx.shimPositions()[p][c], v = ctx.shimVelocities()[p][c];
            if (dblVel) vd[4 * s + c] = v; else vf[4 * s + c] = (float)v;
            if (dblPos) xd[4 * s + c] = x;
            else { xf[4 * s + c] = (float)x; cf[4 * s + c] = (float)(x - (double)(float)x); }
        }
        if (dblVel) vd[4 * s + 3] = mass == 0.0 ? 0.0 : 1.0 / mass; else vf[4 * s + 3] = mass == 0.0 ? 0.f : (float)(1.0 / mass);
    }
    if (dblVel) cu.getVelm().upload(&vd[0]); else cu.getVelm().upload(&vf[0]);
    if (dblPos) cu.getPosq().upload(&xd[0]); else cu.getPosq().upload(&xf[0]);
    if (cu.getUseMixedPrecision()) cu.getPosqCorrection().upload(&cf[0]);
}

void downloadState(ContextImpl& ctx, CudaContext& cu) {
    const int n = cu.getNumAtoms(), padded = cu.getPaddedNumAtoms();
    const bool dblVel = cu.getUseDoublePrecision() || cu.getUseMixedPrecision(), dblPos = cu.getUseDoublePrecision(), mixed = cu.getUseMixedPrecision();
    std::vector<double> vd, xd;
    std::vector<float> vf, xf, cf;
    std::vector<long long> f((size_t)3 * padded);
    if (dblVel) { vd.resize((size_t)4 * padded); cu.getVelm().download(&vd[0]); } else { vf.resize((size_t)4 * padded); cu.getVelm().download(&vf[0]); }
    if (dblPos) { xd.resize((size_t)4 * padded); cu.getPosq().download(&xd[0]); } else { xf.resize((size_t)4 * padded); cu.getPosq().download(&xf[0]); }
    if (mixed) { cf.resize((size_t)4 * padded); cu.getPosqCorrection().download(&cf[0]); }
    cu.getForce().download(&f[0]);
    for (int s = 0; s < n; s++) {
        const int p = cu.getAtomIndex()[s];
        for (int c = 0; c < 3; c++) {
            ctx.shimVelocities()[p][c] = dblVel ? vd[4 * s + c] : (double)vf[4 * s + c];
            ctx.shimPositions()[p][c] = dblPos ? xd[4 * s + c] : mixed ? (double)xf[4 * s + c] + (double)cf[4 * s + c] : (double)xf[4 * s + c];
            ctx.shimForces()[p][c] = (double)f[(size_t)c * padded + s] / 4294967296.0;
        }
    }
    ctx.time = cu.getTime();
}
}  // namespace

CudaPlatform::PlatformData::PlatformData(ContextImpl* context, const System& system, const std::string& precision)
    : context(context), contextsInitialized(false), initializeCalls(0) {
    contexts.push_back(new CudaContext(system, precision, *this));
}
CudaPlatform::PlatformData::~PlatformData() {
    for (size_t i = 0; i < contexts.size(); i++) delete contexts[i];
}
void CudaPlatform::PlatformData::initializeContexts(const System& system) {
    initializeCalls++;
    if (contextsInitialized) return;
    contextsInitialized = true;
}

void CudaPlatform::contextCreated(ContextImpl& context, const std::map<std::string, std::string>& properties) const {
    std::string precision = "single";
    std::map<std::string, std::string>::const_iterator it = properties.find("Precision");
    if (it == properties.end()) it = properties.find("CudaPrecision");
    if (it != properties.end()) precision = it->second;
    PlatformData* data = new PlatformData(&context, context.getSystem(), precision);
    context.setPlatformData(data);
    ContextImpl* ctx = &context;
    context.shimUpload = [ctx, data]() { uploadState(*ctx, *data->contexts[0]); };
    context.shimDownload = [ctx, data]() { downloadState(*ctx, *data->contexts[0]); };
}

void CudaPlatform::contextDestroyed(ContextImpl& context) const {
    delete static_cast<PlatformData*>(context.getPlatformData());
    context.setPlatformData(NULL);
}

/** Host force model routed through the device buffers: positions down (slot order resolved), model, fixed-point forces up.
 *  With model == nullptr the force buffer keeps whatever was installed (fixed forces), but still follows the atom order. */
void shimCudaInstallForceModel(ContextImpl& context, ShimForceModel model, const std::vector<Vec3>* fixedForces) {
    CudaPlatform::PlatformData* data = static_cast<CudaPlatform::PlatformData*>(context.getPlatformData());
    ContextImpl* ctx = &context;
    std::vector<Vec3> fixed = fixedForces ? *fixedForces : std::vector<Vec3>();
    context.shimForcesOnDevice = [ctx, data, model, fixed]() {
        CudaContext& cu = *data->contexts[0];
        const int n = cu.getNumAtoms(), padded = cu.getPaddedNumAtoms();
        if (model) {
            downloadState(*ctx, cu);
            model(ctx->shimPositions(), ctx->shimForces());
        } else if (!fixed.empty())
            ctx->shimForces() = fixed;
        else
            return;
        std::vector<long long> f((size_t)3 * padded, 0);
        for (int s = 0; s < n; s++) {
            const int p = cu.getAtomIndex()[s];
            for (int c = 0; c < 3; c++) f[(size_t)c * padded + s] = (long long)llrint(ctx->shimForces()[p][c] * 4294967296.0);
        }
        cu.getForce().upload(&f[0]);
    };
}

}  // namespace OpenMM
