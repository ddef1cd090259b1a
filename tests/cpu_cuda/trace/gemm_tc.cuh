// gemm_tc.cuh STAND-IN for the call-trace build of embed.cu (tests/cpu_cuda/trace): embed.cu only needs the tensor-map
// type and the stem-window probe from it.
#pragma once
#include <string.h>
#include "common.cuh"
struct CUtensorMap { char opaque[128]; };
namespace ssg {
int make_tmap_stem_windows(CUtensorMap* map, const void* base, uint64_t images);
}
