// Host-side helpers shared by the translation units of libcaro_b200.so.
#pragma once
#include <cuda_runtime.h>

int caro_fail(int code, const char* msg);          // records the thread-local message, returns code
int caro_check_launch(const char* what);           // cudaGetLastError() -> CARO_OK / CARO_E_CUDA
