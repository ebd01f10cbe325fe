// Host-side launch helper shared by every .cu of the library: cudaLaunchKernelEx with an optional thread-block cluster
// and the programmatic-dependent-launch attribute.  With PDL on, a kernel may start (prologue: barrier init, TMEM
// allocation, descriptor prefetch) while its predecessor on the stream is still draining; every kernel launched this way
// executes griddepcontrol.wait (vb::griddep_wait) before it touches memory that an earlier kernel wrote or still reads.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

namespace vb {

inline bool pdl_enabled() {
  static const int on = getenv("VB_PDL") ? atoi(getenv("VB_PDL")) : 1;
  return on != 0;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                             bool pdl, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  unsigned n = 0;
  if (cluster_x > 1) {
    attrs[n].id = cudaLaunchAttributeClusterDimension;
    attrs[n].val.clusterDim.x = static_cast<unsigned>(cluster_x);
    attrs[n].val.clusterDim.y = 1;
    attrs[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl && pdl_enabled()) {
    attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace vb
