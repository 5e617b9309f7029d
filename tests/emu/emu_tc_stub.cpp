// emu_tc_stub.cpp -- the tcgen05 coarse quantizer (dfx_tc.cu) cannot be emulated (tensor-core and
// TMA instructions); in the CPU emulator build of the library it reports "not supported", so the
// search and build drivers take their FFMA paths, exactly as on a shape dfx_tc.cu does not cover.
#define SIMT_IMPLEMENTATION
#include "dfx_internal.h"

bool dfx_tc_supported(int) { return false; }
void dfx_tc_prepare_centroids(dfx_index*, cudaStream_t) {}
void dfx_tc_coarse_search(dfx_index*, const float*, int64_t, int, int32_t*, cudaStream_t) {
    throw DfxError{"tensor-core coarse quantizer is not part of the CPU emulator build"};
}
void dfx_tc_assign(dfx_index*, int, const float*, const float*, int64_t, int, int64_t, const float*, int32_t*,
                   cudaStream_t) {
    throw DfxError{"tensor-core coarse quantizer is not part of the CPU emulator build"};
}
