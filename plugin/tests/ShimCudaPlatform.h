// TEST INFRASTRUCTURE — a "CUDA" platform for the OpenMM API shim: owns device arrays in OpenMM's CUDA layouts
// (single: float4 velm / posq; mixed: double4 velm, float4 posq + float4 posqCorrection; double: double4 velm / posq; SoA
// forces), selected with the
// "Precision" context property like OpenMM's CUDA platform; moves state between them and the Context's host copies, and evaluates the shim's
// host force model into the device force buffer.  It lets the real plugin stack (DrudeTGNHIntegrator -> KernelImpl ->
// C-ABI -> sm_100a kernels) run end to end without OpenMM.
#ifndef TGNH_SHIM_CUDA_PLATFORM_H_
#define TGNH_SHIM_CUDA_PLATFORM_H_
#include <cuda_runtime.h>

#include <cmath>
#include <vector>

#include "../src/B200DrudeTGNHKernels.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"

namespace OpenMM {

class ShimCudaPlatform : public Platform {
public:
    class Data : public TgnhDeviceAccess {
    public:
        Data(ContextImpl& c, int forceFormat, int precision) : constraintCalls(0), ctx(c), n(c.getSystem().getNumParticles()), padded(((n + 31) / 32) * 32), forceFormat(forceFormat),
              precision(precision), mixed(precision != TGNH_PRECISION_SINGLE), dbl(precision == TGNH_PRECISION_DOUBLE), velm(NULL), posq(NULL), corr(NULL), force(NULL), posDelta(NULL), time(0.0), steps(0) {
            int count = 0;
            if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) throw OpenMMException("ShimCudaPlatform: no CUDA device");
            const size_t fbytes = (size_t)3 * padded * (forceFormat == TGNH_FORCE_I64_SOA ? 8 : 4);
            const size_t vbytes = (size_t)padded * (mixed ? 32 : 16);
            const size_t xbytes = (size_t)padded * (dbl ? 32 : 16);
            if (cudaMalloc(&velm, vbytes) || cudaMalloc(&posq, xbytes) || cudaMalloc(&corr, (size_t)padded * 16) || cudaMalloc(&force, fbytes) ||
                cudaMalloc(&posDelta, vbytes))
                throw OpenMMException("ShimCudaPlatform: cudaMalloc failed");
            cudaMemset(velm, 0, vbytes); cudaMemset(posq, 0, xbytes); cudaMemset(corr, 0, (size_t)padded * 16); cudaMemset(force, 0, fbytes);
            hv.assign((size_t)padded * 4, 0.f); hx.assign((size_t)padded * 4, 0.f); hc.assign((size_t)padded * 4, 0.f); hvd.assign((size_t)padded * 4, 0.0); hxd.assign((size_t)padded * 4, 0.0);
        }
        ~Data() { cudaFree(velm); cudaFree(posq); cudaFree(corr); cudaFree(force); cudaFree(posDelta); }
        TgnhDeviceView view() {
            TgnhDeviceView v = {velm, posq, force, posDelta, padded, forceFormat, NULL, 0, precision, precision == TGNH_PRECISION_MIXED ? corr : NULL};
            return v;
        }
        void advanceTime(double dt) { time += dt; steps++; ctx.time = time; }
        // the shim has no constraint solver: the calls are counted so that tests can see the split sequence ran
        void applyConstraints(double) { constraintCalls++; }
        void applyVelocityConstraints(double) { constraintCalls++; }
        int constraintCalls;
        void upload() {
            for (int i = 0; i < n; i++) {
                const double m = ctx.getSystem().getParticleMass(i);
                for (int c = 0; c < 3; c++) {
                    const double x = ctx.shimPositions()[i][c], v = ctx.shimVelocities()[i][c];
                    hv[4 * i + c] = (float)v; hvd[4 * i + c] = v;
                    hx[4 * i + c] = (float)x; hc[4 * i + c] = (float)(x - (double)(float)x); hxd[4 * i + c] = x;
                }
                hv[4 * i + 3] = m == 0.0 ? 0.f : (float)(1.0 / m);
                hvd[4 * i + 3] = m == 0.0 ? 0.0 : 1.0 / m;
            }
            if (mixed) {
                cudaMemcpy(velm, hvd.data(), hvd.size() * 8, cudaMemcpyHostToDevice);
                cudaMemcpy(corr, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice);
            } else
                cudaMemcpy(velm, hv.data(), hv.size() * 4, cudaMemcpyHostToDevice);
            if (dbl) cudaMemcpy(posq, hxd.data(), hxd.size() * 8, cudaMemcpyHostToDevice);
            else cudaMemcpy(posq, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
        }
        void download() {
            if (dbl) cudaMemcpy(hxd.data(), posq, hxd.size() * 8, cudaMemcpyDeviceToHost);
            else cudaMemcpy(hx.data(), posq, hx.size() * 4, cudaMemcpyDeviceToHost);
            if (mixed) {
                cudaMemcpy(hvd.data(), velm, hvd.size() * 8, cudaMemcpyDeviceToHost);
                cudaMemcpy(hc.data(), corr, hc.size() * 4, cudaMemcpyDeviceToHost);
            } else
                cudaMemcpy(hv.data(), velm, hv.size() * 4, cudaMemcpyDeviceToHost);
            for (int i = 0; i < n; i++)
                for (int c = 0; c < 3; c++) {
                    ctx.shimVelocities()[i][c] = mixed ? hvd[4 * i + c] : (double)hv[4 * i + c];
                    ctx.shimPositions()[i][c] = dbl ? hxd[4 * i + c] : mixed ? (double)hx[4 * i + c] + (double)hc[4 * i + c] : (double)hx[4 * i + c];
                }
        }
        /** forces: host model on the downloaded positions, written into the device buffer in the platform's format */
        void forcesOnDevice(const ShimForceModel& model) {
            download();
            model(ctx.shimPositions(), ctx.shimForces());
            if (forceFormat == TGNH_FORCE_I64_SOA) {
                std::vector<long long> f((size_t)3 * padded, 0);
                for (int i = 0; i < n; i++) for (int c = 0; c < 3; c++) f[(size_t)c * padded + i] = (long long)llrint(ctx.shimForces()[i][c] * 4294967296.0);
                cudaMemcpy(force, f.data(), f.size() * 8, cudaMemcpyHostToDevice);
            } else {
                std::vector<float> f((size_t)3 * padded, 0.f);
                for (int i = 0; i < n; i++) for (int c = 0; c < 3; c++) f[(size_t)c * padded + i] = (float)ctx.shimForces()[i][c];
                cudaMemcpy(force, f.data(), f.size() * 4, cudaMemcpyHostToDevice);
            }
        }
    private:
        ContextImpl& ctx;
        int n, padded, forceFormat;
        int precision;
        bool mixed, dbl;      // mixed: double velm (mixed and double modes); dbl: double posq too
        void *velm, *posq, *corr, *force, *posDelta;
        double time;
        int steps;
        std::vector<float> hv, hx, hc;
        std::vector<double> hvd, hxd;
    };
    explicit ShimCudaPlatform(int forceFormat = TGNH_FORCE_I64_SOA) : forceFormat(forceFormat) {}
    const std::string& getName() const { static const std::string n = "CUDA"; return n; }
    void contextCreated(ContextImpl& context, const std::map<std::string, std::string>& properties) const {
        std::map<std::string, std::string>::const_iterator it = properties.find("Precision");
        const std::string precision = it == properties.end() ? "single" : it->second;
        if (precision != "single" && precision != "mixed" && precision != "double") throw OpenMMException("ShimCudaPlatform: Precision must be single, mixed or double");
        Data* d = new Data(context, forceFormat, precision == "double" ? TGNH_PRECISION_DOUBLE : precision == "mixed" ? TGNH_PRECISION_MIXED : TGNH_PRECISION_SINGLE);
        context.setPlatformData(static_cast<TgnhDeviceAccess*>(d));
        context.shimUpload = [d]() { d->upload(); };
        context.shimDownload = [d]() { d->download(); };
    }
    void contextDestroyed(ContextImpl& context) const { delete static_cast<Data*>(static_cast<TgnhDeviceAccess*>(context.getPlatformData())); }
    /** route the context's force model through the device buffers */
    static void installForceModel(ContextImpl& context, ShimForceModel model) {
        Data* d = static_cast<Data*>(static_cast<TgnhDeviceAccess*>(context.getPlatformData()));
        context.shimForcesOnDevice = [d, model]() { d->forcesOnDevice(model); };
    }
private:
    int forceFormat;
};

}  // namespace OpenMM
#endif
