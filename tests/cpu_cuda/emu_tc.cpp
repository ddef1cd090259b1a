// Host side of the tcgen05 / TMA emulation: the "driver" tensor-map encoder records its arguments (stub_tc/cuda.h).
// TEST INFRASTRUCTURE (tests/cpu_cuda).
#include <string.h>
#include <cuda.h>
#include <cuda_runtime.h>

static CUresult emu_encode_tiled(CUtensorMap* m, CUtensorMapDataType, cuuint32_t rank, void* base, const cuuint64_t* dims,
                                 const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr,
                                 CUtensorMapInterleave, CUtensorMapSwizzle swz, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
    memset(m, 0, sizeof(*m));
    m->base = (uint64_t)(uintptr_t)base;
    m->rank = rank;
    m->swizzle = (uint32_t)swz;
    for (cuuint32_t i = 0; i < rank; ++i) {
        m->dims[i] = (uint32_t)dims[i];
        m->box[i] = box[i];
        m->estr[i] = estr[i];
        if (i + 1 < rank) m->strides[i] = strides[i];
    }
    return CUDA_SUCCESS;
}

cudaError_t cudaGetDriverEntryPoint(const char* name, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* res) {
    *fn = strcmp(name, "cuTensorMapEncodeTiled") == 0 ? (void*)emu_encode_tiled : nullptr;
    if (res) *res = cudaDriverEntryPointSuccess;
    return cudaSuccess;
}
