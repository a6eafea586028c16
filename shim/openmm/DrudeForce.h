#ifndef SHIM_OPENMM_DRUDEFORCE_H_
#define SHIM_OPENMM_DRUDEFORCE_H_
#include "openmm/shim_core.h"
namespace OpenMM {
/** OpenMM's Drude plugin force: only the particle table the integrator reads (the force itself stays in OpenMM). */
class DrudeForce : public Force {
public:
    int getNumParticles() const { return (int)particles.size(); }
    int addParticle(int particle, int particle1, int particle2, int particle3, int particle4, double charge, double polarizability, double aniso12, double aniso34) {
        ParticleInfo p = {particle, particle1, particle2, particle3, particle4, charge, polarizability, aniso12, aniso34};
        particles.push_back(p);
        return (int)particles.size() - 1;
    }
    void getParticleParameters(int index, int& particle, int& particle1, int& particle2, int& particle3, int& particle4, double& charge,
                               double& polarizability, double& aniso12, double& aniso34) const {
        if (index < 0 || index >= (int)particles.size()) throw OpenMMException("Index out of range");
        const ParticleInfo& p = particles[index];
        particle = p.p; particle1 = p.p1; particle2 = p.p2; particle3 = p.p3; particle4 = p.p4;
        charge = p.charge; polarizability = p.polarizability; aniso12 = p.aniso12; aniso34 = p.aniso34;
    }
    std::vector<std::pair<int, int> > shimGetBondedParticles() const {
        std::vector<std::pair<int, int> > b;
        for (size_t i = 0; i < particles.size(); i++) b.push_back(std::make_pair(particles[i].p, particles[i].p1));
        return b;
    }
private:
    struct ParticleInfo { int p, p1, p2, p3, p4; double charge, polarizability, aniso12, aniso34; };
    std::vector<ParticleInfo> particles;
};
}  // namespace OpenMM
#endif
