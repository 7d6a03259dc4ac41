// oracle/ref_shim: empty stand-in (sampling_gpu.h includes it but the .cu uses nothing from it)
#pragma once
#include <cuda_runtime_api.h>
