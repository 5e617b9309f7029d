// the one translation unit of libdfx_emu_full.so that defines the context-switch primitive of simt.h
#define SIMT_IMPLEMENTATION
#include <cuda_runtime.h>
