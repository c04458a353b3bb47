// Host-callable launchers of the broad-phase kernels (implemented in broad.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
