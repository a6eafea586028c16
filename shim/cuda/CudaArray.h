// shim/cuda/CudaArray.h — stand-in for OpenMM 7.x's CudaArray (platforms/cuda/include/CudaArray.h), written from the
// members the TGNH CUDA sources call.  TEST / BUILD INFRASTRUCTURE (shim/README.md): real device memory through the CUDA
// driver API, so that the reference's own platforms/cuda sources and this repo's OpenMM-facing glue run on a GPU without OpenMM.
#ifndef SHIM_CUDA_ARRAY_H_
#define SHIM_CUDA_ARRAY_H_
#include <cuda.h>

#include <string>
#include <vector>

#include "openmm/OpenMMException.h"

namespace OpenMM {

class CudaContext;

class CudaArray {
public:
    /** OpenMM: CudaArray::create<T>(context, size, name) */
    template <class T>
    static CudaArray* create(CudaContext& context, int size, const std::string& name) { return new CudaArray(context, size, sizeof(T), name); }
    CudaArray(CudaContext& context, int size, int elementSize, const std::string& name);
    ~CudaArray();
    int getSize() const { return size; }
    int getElementSize() const { return elementSize; }
    const std::string& getName() const { return name; }
    CudaContext& getContext() { return *context; }
    /** a reference: the reference's sources take its address to build kernel argument lists */
    CUdeviceptr& getDevicePointer() { return pointer; }
    template <class T>
    void upload(const std::vector<T>& data, bool convert = false) {
        if (convert && (int)data.size() == size && (int)sizeof(T) != elementSize) {
            // double <-> float, as OpenMM does for precision-dependent arrays
            if (sizeof(T) == 2 * (size_t)elementSize) {
                std::vector<float> v(elementSize / 4 * size);
                const double* d = reinterpret_cast<const double*>(&data[0]);
                for (size_t i = 0; i < v.size(); i++) v[i] = (float)d[i];
                upload(&v[0], true);
                return;
            }
            if (2 * sizeof(T) == (size_t)elementSize) {
                std::vector<double> v(elementSize / 8 * size);
                const float* d = reinterpret_cast<const float*>(&data[0]);
                for (size_t i = 0; i < v.size(); i++) v[i] = (double)d[i];
                upload(&v[0], true);
                return;
            }
        }
        if ((int)sizeof(T) != elementSize || (int)data.size() != size)
            throw OpenMMException("Error uploading array " + name + ": The specified vector does not match the size of the array");
        upload(&data[0], true);
    }
    template <class T>
    void download(std::vector<T>& data) const {
        if ((int)sizeof(T) != elementSize) throw OpenMMException("Error downloading array " + name + ": The specified vector has the wrong element size");
        if ((int)data.size() != size) data.resize(size);
        download(&data[0], true);
    }
    void upload(const void* data, bool blocking = true);
    void download(void* data, bool blocking = true) const;
private:
    CudaContext* context;
    CUdeviceptr pointer;
    int size, elementSize;
    std::string name;
    CudaArray(const CudaArray&);
    CudaArray& operator=(const CudaArray&);
};

}  // namespace OpenMM
#endif
