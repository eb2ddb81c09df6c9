# fmcts_b200.jl — drop-in `module FMCTS` for the interactive drivers (testHex.jl, testgobang.jl, testrev6.jl, testrev8.jl).
#
# The reference's `FMCTS.MctsContext(c, nn, prealloc)(pos, readout) -> (p, v)` is a CPU tree search (fast_mcts.jl:262-302); each
# driver carries, commented out, the GPU call it would be replaced with (testrev8.jl:24, testgobang.jl:25, testHex.jl:39):
#     _,p = mcts_gpu.mcts_single(actor, readout, 256, vnodes, vnodesStats, leaf, newindex, 1, training=false, cpuct=1.5, …)
# This module is that replacement: one game, training=false, through libalphagpu.so (mcts_gpu_b200.jl must be included first).
# `nn` is the actor the drivers already hold (`convert_back(actor)`; the `_cpu` copy is no longer needed), `prealloc` is ignored.
# NOT EXECUTED here (no Julia in the build container or on the GPU box); the Python twin alphagpu_b200/fast_mcts.py is what
# tests/test_fast_mcts.py runs against the oracle.
module FMCTS

export MctsContext

using ..Game
using ..mcts_gpu

const MAX_READOUT = 255     # node ids are 8 bit in the tree layout (alphagpu.h: agpu_config.rollouts)

mutable struct MctsContext
    c::Float32
    nn
    prealloc
    ctx::Union{Nothing,mcts_gpu.Context}
    calls::UInt32
end
MctsContext(c, nn, prealloc=nothing) = MctsContext(Float32(c), nn, prealloc, nothing, 0)

# decode_cpu(pos) (fast_mcts.jl:105-124): the 0/1 encoding the drivers feed to the actor for their "α, β" printout
function decode_cpu(pos, fstate=nothing)
    fstate === nothing && (fstate = zeros(Int8, 2 * VectorizedState))
    for j in 1:VectorizedState
        fstate[j] = pos.bplayer[j] ? 1 : 0
        fstate[j + VectorizedState] = pos.bopponent[j] ? 1 : 0
    end
    return fstate
end

# the call operator of fast_mcts.jl:270-289; `komi` is accepted and ignored as `evaluate` ignores it (fast_mcts.jl:126-142)
function (m::MctsContext)(pos::Position, readout, komi=0)
    1 <= readout <= MAX_READOUT || error("readout must be 1..$MAX_READOUT")
    if m.ctx === nothing || m.ctx.visits < readout
        m.ctx = mcts_gpu.init(1, readout, m.nn)
    end
    positions = [pos]
    uid = UInt32[m.calls]                     # a fresh random stream per call
    GC.@preserve positions uid mcts_gpu.check(m.ctx.h, ccall((:agpu_reinit, mcts_gpu.LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{UInt32}),
                                                              m.ctx.h, pointer(positions), 1, pointer(uid)))
    m.calls += 1
    p, _ = mcts_gpu.mcts_single(m.ctx, readout, 1, training=false, cpuct=m.c)
    # extractRoot (fast_mcts.jl:293-302): v = sum(w)/N with w = q.*n of the root and N = readout
    R = m.ctx.visits
    q = Array{Float32}(undef, maxActions, R); n = Array{Float32}(undef, maxActions, R)
    dump = Ptr{Cvoid}[C_NULL for _ in 1:11]   # agpu_tree_dump: nnodes parent action child order nchild expanded prior q visits states
    dump[9] = pointer(q); dump[10] = pointer(n)
    GC.@preserve q n dump mcts_gpu.check(m.ctx.h, ccall((:agpu_get_tree, mcts_gpu.LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Ptr{Cvoid}}), m.ctx.h, 1, dump))
    v = sum(Float64.(q[:, 1]) .* Float64.(n[:, 1])) / readout
    return vec(p), v
end

end # module
