// Hex plugin instantiations (Main.N is a source constant in the reference, mainHex.jl:23-24).
#include "engine.cuh"
#define AG_HEX_SIZES(X) X(5) X(7)
namespace ag {
EngineBase* make_engine_hex(int n) {
#define X(N) if (n == N) return new EngineT<Hex<N>>();
  AG_HEX_SIZES(X)
#undef X
  return nullptr;
}
}
