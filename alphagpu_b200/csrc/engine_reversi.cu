#include "engine.cuh"
namespace ag {
EngineBase* make_engine_reversi(int n) {
  if (n == 8) return new EngineT<Reversi<8>>();
  if (n == 6) return new EngineT<Reversi<6>>();
  return nullptr;
}
}
