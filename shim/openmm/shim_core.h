// shim/openmm/shim_core.h — the subset of the OpenMM 7.x API that the TGNH plugin sources touch, header-only.
// Build/test infrastructure (see shim/README.md); the class and member names are OpenMM's so that the
// reference's own sources and this repo's plugin/ compile against it unchanged.
#ifndef SHIM_OPENMM_CORE_H_
#define SHIM_OPENMM_CORE_H_

#include <algorithm>
#include <functional>
#include <map>
#include <set>
#include <string>
#include <typeinfo>
#include <utility>
#include <vector>

#include "openmm/OpenMMException.h"
#include "openmm/Vec3.h"
#include "openmm/internal/windowsExport.h"

namespace OpenMM {

class Context;
class ContextImpl;
class Platform;
class System;

// ---------------------------------------------------------------------------------------------- forces
class Force {
public:
    Force() : forceGroup(0) {}
    virtual ~Force() {}
    int getForceGroup() const { return forceGroup; }
    void setForceGroup(int g) { forceGroup = g; }
    /** shim: pairs of particles this force bonds (ContextImpl::getMolecules groups by them) */
    virtual std::vector<std::pair<int, int> > shimGetBondedParticles() const { return std::vector<std::pair<int, int> >(); }
private:
    int forceGroup;
};

class CMMotionRemover : public Force {
public:
    explicit CMMotionRemover(int frequency = 1) : frequency(frequency) {}
    int getFrequency() const { return frequency; }
private:
    int frequency;
};

/** shim-only: declares bonds so that molecules larger than a Drude pair exist without a force field */
class ShimBondForce : public Force {
public:
    void addBond(int a, int b) { bonds.push_back(std::make_pair(a, b)); }
    std::vector<std::pair<int, int> > shimGetBondedParticles() const { return bonds; }
private:
    std::vector<std::pair<int, int> > bonds;
};

class VirtualSite {
public:
    virtual ~VirtualSite() {}
};

// ---------------------------------------------------------------------------------------------- System
class System {
public:
    System() {}
    ~System() {
        for (size_t i = 0; i < forces.size(); i++) delete forces[i];
        for (std::map<int, VirtualSite*>::iterator it = vsites.begin(); it != vsites.end(); ++it) delete it->second;
    }
    int getNumParticles() const { return (int)masses.size(); }
    int addParticle(double mass) { masses.push_back(mass); return (int)masses.size() - 1; }
    double getParticleMass(int index) const { check(index, masses.size()); return masses[index]; }
    void setParticleMass(int index, double mass) { check(index, masses.size()); masses[index] = mass; }
    int getNumConstraints() const { return (int)constraints.size(); }
    int addConstraint(int p1, int p2, double distance) { Constraint c = {p1, p2, distance}; constraints.push_back(c); return (int)constraints.size() - 1; }
    void getConstraintParameters(int index, int& p1, int& p2, double& distance) const {
        check(index, constraints.size());
        p1 = constraints[index].p1; p2 = constraints[index].p2; distance = constraints[index].d;
    }
    int addForce(Force* force) { forces.push_back(force); return (int)forces.size() - 1; }
    int getNumForces() const { return (int)forces.size(); }
    const Force& getForce(int index) const { check(index, forces.size()); return *forces[index]; }
    Force& getForce(int index) { check(index, forces.size()); return *forces[index]; }
    void setVirtualSite(int index, VirtualSite* site) { vsites[index] = site; }
    bool isVirtualSite(int index) const { return vsites.count(index) != 0; }
    void setDefaultPeriodicBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c) { box[0] = a; box[1] = b; box[2] = c; }
private:
    struct Constraint { int p1, p2; double d; };
    static void check(int index, size_t n) { if (index < 0 || index >= (int)n) throw OpenMMException("Index out of range"); }
    std::vector<double> masses;
    std::vector<Constraint> constraints;
    std::vector<Force*> forces;
    std::map<int, VirtualSite*> vsites;
    Vec3 box[3];
    System(const System&);
    System& operator=(const System&);
};

// ---------------------------------------------------------------------------------------------- kernels
class KernelImpl {
public:
    KernelImpl(std::string name, const Platform& platform) : name(name), platform(&platform), referenceCount(1) {}
    virtual ~KernelImpl() {}
    std::string getName() const { return name; }
    const Platform& getPlatform() { return *platform; }
private:
    friend class Kernel;
    std::string name;
    const Platform* platform;
    int referenceCount;
};

/** ref-counted handle, as in OpenMM */
class Kernel {
public:
    Kernel() : impl(NULL) {}
    Kernel(KernelImpl* impl) : impl(impl) {}
    Kernel(const Kernel& copy) : impl(copy.impl) { if (impl) impl->referenceCount++; }
    ~Kernel() { release(); }
    Kernel& operator=(const Kernel& copy) {
        if (copy.impl) copy.impl->referenceCount++;
        release();
        impl = copy.impl;
        return *this;
    }
    std::string getName() const { return impl->getName(); }
    const KernelImpl& getImpl() const { return *impl; }
    KernelImpl& getImpl() { return *impl; }
    template <class T> const T& getAs() const { return dynamic_cast<const T&>(*impl); }
    template <class T> T& getAs() { return dynamic_cast<T&>(*impl); }
private:
    void release() { if (impl && --impl->referenceCount == 0) delete impl; impl = NULL; }
    KernelImpl* impl;
};

class KernelFactory {
public:
    virtual KernelImpl* createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const = 0;
    virtual ~KernelFactory() {}
};

// ---------------------------------------------------------------------------------------------- Platform
class Platform {
public:
    virtual ~Platform() {}
    virtual const std::string& getName() const = 0;
    virtual double getSpeed() const { return 1.0; }
    virtual bool supportsDoublePrecision() const { return true; }
    /** shim: every platform owns opaque per-context data, created when the ContextImpl is built */
    virtual void contextCreated(ContextImpl& context, const std::map<std::string, std::string>& properties) const {}
    virtual void contextDestroyed(ContextImpl& context) const {}
    void registerKernelFactory(const std::string& name, KernelFactory* factory) { kernelFactories[name] = factory; }
    bool supportsKernels(const std::vector<std::string>& names) const {
        for (size_t i = 0; i < names.size(); i++) if (!kernelFactories.count(names[i])) return false;
        return true;
    }
    Kernel createKernel(const std::string& name, ContextImpl& context) const {
        std::map<std::string, KernelFactory*>::const_iterator it = kernelFactories.find(name);
        if (it == kernelFactories.end()) throw OpenMMException("Called createKernel() on a Platform which does not support the requested kernel");
        return Kernel(it->second->createKernelImpl(name, *this, context));
    }
    void setPropertyDefaultValue(const std::string& property, const std::string& value) { defaults[property] = value; }
    const std::string& getPropertyDefaultValue(const std::string& property) const {
        static const std::string empty;
        std::map<std::string, std::string>::const_iterator it = defaults.find(property);
        return it == defaults.end() ? empty : it->second;
    }
    static void registerPlatform(Platform* platform) { platforms().push_back(platform); }
    static int getNumPlatforms() { return (int)platforms().size(); }
    static Platform& getPlatform(int index) { return *platforms().at(index); }
    static Platform& getPlatformByName(const std::string& name) {
        for (size_t i = 0; i < platforms().size(); i++) if (platforms()[i]->getName() == name) return *platforms()[i];
        throw OpenMMException("There is no registered Platform called \"" + name + "\"");
    }
private:
    static std::vector<Platform*>& platforms() { static std::vector<Platform*> p; return p; }
    std::map<std::string, KernelFactory*> kernelFactories;
    std::map<std::string, std::string> defaults;
};

// ---------------------------------------------------------------------------------------------- State
class State {
public:
    enum DataType { Positions = 1, Velocities = 2, Forces = 4, Energy = 8, Parameters = 16 };
    State() : time(0), ke(0), pe(0) {}
    double getTime() const { return time; }
    const std::vector<Vec3>& getPositions() const { return positions; }
    const std::vector<Vec3>& getVelocities() const { return velocities; }
    const std::vector<Vec3>& getForces() const { return forces; }
    double getKineticEnergy() const { return ke; }
    double getPotentialEnergy() const { return pe; }
private:
    friend class Context;
    double time, ke, pe;
    std::vector<Vec3> positions, velocities, forces;
};

// ---------------------------------------------------------------------------------------------- Integrator
class Integrator {
public:
    Integrator() : owner(NULL), context(NULL), stepSize(0), constraintTol(1e-5) {}
    virtual ~Integrator() {}
    virtual double getStepSize() const { return stepSize; }
    virtual void setStepSize(double size) { stepSize = size; }
    virtual double getConstraintTolerance() const { return constraintTol; }
    virtual void setConstraintTolerance(double tol) { constraintTol = tol; }
    virtual void step(int steps) = 0;
protected:
    friend class Context;
    friend class ContextImpl;
    Context* owner;
    ContextImpl* context;
    virtual void initialize(ContextImpl& context) = 0;
    virtual void cleanup() {}
    virtual std::vector<std::string> getKernelNames() = 0;
    virtual void stateChanged(State::DataType changed) {}
    virtual double computeKineticEnergy() = 0;
private:
    double stepSize, constraintTol;
};

// ---------------------------------------------------------------------------------------------- ContextImpl / Context
/** Force model of the shim: fills `forces` for the given positions (the real ContextImpl runs OpenMM's force kernels). */
typedef std::function<void(const std::vector<Vec3>& positions, std::vector<Vec3>& forces)> ShimForceModel;

class ContextImpl {
public:
    ContextImpl(Context& owner, const System& system, Integrator& integrator, Platform* platform)
        : time(0), owner(owner), system(system), integrator(integrator), platform(platform), platformData(NULL), lastForceGroups(-1),
          forceCalls(0) {
        int n = system.getNumParticles();
        positions.assign(n, Vec3()); velocities.assign(n, Vec3()); forces.assign(n, Vec3());
    }
    Context& getOwner() { return owner; }
    const System& getSystem() const { return system; }
    Integrator& getIntegrator() { return integrator; }
    Platform& getPlatform() { return *platform; }
    void* getPlatformData() { return platformData; }
    const void* getPlatformData() const { return platformData; }
    void setPlatformData(void* data) { platformData = data; }
    /** connected components of the bond graph (DrudeForce pairs, constraints, declared bonds), ordered by first particle */
    std::vector<std::vector<int> > getMolecules() const;
    bool updateContextState() { return false; }
    int getLastForceGroups() const { return lastForceGroups; }
    double calcForcesAndEnergy(bool includeForces, bool includeEnergy, int groups = 0xFFFFFFFF) {
        lastForceGroups = groups;
        forceCalls++;
        if (shimForcesOnDevice) shimForcesOnDevice();
        else if (forceModel) forceModel(shimPositions(), shimForces());
        return 0.0;
    }
    // ---- shim state: host copies used by host platforms (the shim "Reference" platform points its PlatformData here)
    std::vector<Vec3>& shimPositions() { return positions; }
    std::vector<Vec3>& shimVelocities() { return velocities; }
    std::vector<Vec3>& shimForces() { return forces; }
    void shimSetForceModel(ShimForceModel model) { forceModel = model; }
    /** device platforms install a hook that evaluates forces on device buffers instead */
    std::function<void()> shimForcesOnDevice;
    /** device platforms install hooks that move state between the host copies above and their device arrays */
    std::function<void()> shimUpload, shimDownload;
    int shimForceCalls() const { return forceCalls; }
    double time;
private:
    Context& owner;
    const System& system;
    Integrator& integrator;
    Platform* platform;
    void* platformData;
    int lastForceGroups;
    int forceCalls;
    std::vector<Vec3> positions, velocities, forces;
    ShimForceModel forceModel;
};

class Context {
public:
    Context(const System& system, Integrator& integrator, Platform& platform, const std::map<std::string, std::string>& properties = std::map<std::string, std::string>())
        : system(system), integrator(integrator) {
        impl = new ContextImpl(*this, system, integrator, &platform);
        platform.contextCreated(*impl, properties);
        integrator.initialize(*impl);
    }
    ~Context() {
        integrator.cleanup();
        impl->getPlatform().contextDestroyed(*impl);
        delete impl;
    }
    const System& getSystem() const { return system; }
    Integrator& getIntegrator() { return integrator; }
    Platform& getPlatform() { return impl->getPlatform(); }
    void setPositions(const std::vector<Vec3>& p) { check(p); impl->shimPositions() = p; upload(); integrator.stateChanged(State::Positions); }
    void setVelocities(const std::vector<Vec3>& v) { check(v); impl->shimVelocities() = v; upload(); integrator.stateChanged(State::Velocities); }
    State getState(int types) {
        if (impl->shimDownload) impl->shimDownload();
        State s;
        s.time = impl->time;
        if (types & State::Positions) s.positions = impl->shimPositions();
        if (types & State::Velocities) s.velocities = impl->shimVelocities();
        if (types & State::Forces) s.forces = impl->shimForces();
        if (types & State::Energy) s.ke = integrator.computeKineticEnergy();
        return s;
    }
    void applyConstraints(double) {}
    ContextImpl& getImpl() { return *impl; }   // shim: tests reach the force model through it
private:
    void upload() { if (impl->shimUpload) impl->shimUpload(); }
    void check(const std::vector<Vec3>& v) const { if ((int)v.size() != system.getNumParticles()) throw OpenMMException("Called setPositions()/setVelocities() on a Context with the wrong number of values"); }
    const System& system;
    Integrator& integrator;
    ContextImpl* impl;
};

}  // namespace OpenMM

#include "openmm/DrudeForce.h"

namespace OpenMM {
inline std::vector<std::vector<int> > ContextImpl::getMolecules() const {
    const int n = system.getNumParticles();
    std::vector<int> parent(n);
    for (int i = 0; i < n; i++) parent[i] = i;
    std::function<int(int)> find = [&](int x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
    auto unite = [&](int a, int b) { a = find(a); b = find(b); if (a != b) parent[std::max(a, b)] = std::min(a, b); };
    for (int i = 0; i < system.getNumConstraints(); i++) { int a, b; double d; system.getConstraintParameters(i, a, b, d); unite(a, b); }
    for (int f = 0; f < system.getNumForces(); f++) {
        std::vector<std::pair<int, int> > bonds = system.getForce(f).shimGetBondedParticles();
        for (size_t i = 0; i < bonds.size(); i++) unite(bonds[i].first, bonds[i].second);
    }
    std::map<int, int> index;
    std::vector<std::vector<int> > molecules;
    for (int i = 0; i < n; i++) {
        int r = find(i);
        if (!index.count(r)) { index[r] = (int)molecules.size(); molecules.push_back(std::vector<int>()); }
        molecules[index[r]].push_back(i);
    }
    return molecules;
}
}  // namespace OpenMM

#endif
