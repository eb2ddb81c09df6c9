#include "engine.cuh"
namespace ag {
EngineBase* make_engine_connect4() { return new EngineT<Connect4>(); }
}
