// Gobang plugin instantiations: (Main.N, Main.Nvict) are compile-time constants in the reference too
// (mainGobang.jl:24-26).  Add a size by adding a line to AG_GOBANG_SIZES.
#include "engine.cuh"
#define AG_GOBANG_SIZES(X) X(3, 3) X(5, 4) X(9, 5)
namespace ag {
EngineBase* make_engine_gobang(int n, int nvict) {
#define X(N, NV) if (n == N && nvict == NV) return new EngineT<Gobang<N, NV>>();
  AG_GOBANG_SIZES(X)
#undef X
  return nullptr;
}
}
