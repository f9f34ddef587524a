// TEST INFRASTRUCTURE: the driver types the engine names live in the runtime stand-in (see cuda_runtime.h)
#pragma once
#include <cuda_runtime.h>
