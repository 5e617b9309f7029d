// stand-in for <cuda_fp16.h> (CPU emulator build only; no half arithmetic is emulated)
#pragma once
struct __half { unsigned short x; };
