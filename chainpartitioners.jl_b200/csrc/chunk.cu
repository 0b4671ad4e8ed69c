// chunk.cu -- pack_stripe: the VBR chunkers (DynamicChunker.jl, ConvexTotalChunker.jl,
// OverlapChunker.jl, StrictChunker.jl, EquiPartitioner.jl).
#include <algorithm>
#include "engine.cuh"
#include "primitives.cuh"

namespace cpb {

void solve_pack(Matrix& A, Oracle* f, int method, const cpb_constraint* con, double rho, i64 w_max, int64_t* h_spl_out,
                int64_t* K_out, int64_t* n_nets_out) {
  const i64 n = A.n;
  switch (method) {
    case CPB_PACK_EQUI: {  // EquiPartitioner.jl:15-21: [1:w:n; n+1]
      CPB_REQUIRE(w_max >= 1, "EquiChunker width must be >= 1");
      i64 K = 0;
      for (i64 j = 1; j <= n; j += w_max) h_spl_out[K++] = j;
      h_spl_out[K] = n + 1;
      *K_out = K;
      return;
    }
    default: throw Error(CPB_ERR_UNSUPPORTED, "pack_stripe method not built yet");
  }
}

}  // namespace cpb
