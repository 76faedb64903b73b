// v3d_mlp.cu -- learned-projection PE (PositionEmbeddingMLP) on tcgen05/TMEM.  Placeholder until
// the tensor-core kernel lands: the entry point exists (include/v3d.h) and fails loudly.
#include "v3d_common.cuh"

extern "C" int v3d_pe_mlp(const float*, int, int, const float*, const float*, const float*, const float*, float, const void*,
                          const float*, const void*, void*, v3d_stream_t) {
  v3d::set_error("v3d_pe_mlp: tcgen05 kernel not built into this library yet");
  return V3D_E_UNSUPPORTED;
}
