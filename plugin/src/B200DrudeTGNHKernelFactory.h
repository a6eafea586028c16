#ifndef TGNH_B200_KERNEL_FACTORY_H_
#define TGNH_B200_KERNEL_FACTORY_H_
#include "openmm/KernelFactory.h"
namespace OpenMM {
/** Creates the B200 TGNH kernel for OpenMM's "CUDA" platform (replaces CudaDrudeTGNHKernelFactory). */
class B200DrudeTGNHKernelFactory : public KernelFactory {
public:
    KernelImpl* createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const;
};
}  // namespace OpenMM
#endif
