// api.cu -- version / error strings of librv3d.
#include "common.cuh"

extern "C" int rv3d_version(void) { return RV3D_VERSION; }

extern "C" const char *rv3d_strerror(int status) {
  switch (status) {
    case RV3D_OK: return "ok";
    case RV3D_ERR_ARG: return "invalid argument (shape, null pointer or enum value)";
    case RV3D_ERR_ALIGN: return "pointer alignment requirement not met";
    case RV3D_ERR_SCRATCH: return "scratch buffer too small";
    case RV3D_ERR_CUDA: return "CUDA runtime error";
    case RV3D_ERR_KEYBITS: return "sort key does not fit 64 bits: split the batch";
    case RV3D_ERR_CAPACITY: return "output capacity exceeded";
    default: return "unknown rv3d status";
  }
}
