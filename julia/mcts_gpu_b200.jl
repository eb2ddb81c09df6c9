# mcts_gpu_b200.jl — drop-in `module mcts_gpu` over libalphagpu.so (C ABI: include/alphagpu.h).
#
# Replaces `include("mcts_gpu.jl")` in main4IARow.jl / mainGobang.jl / mainHex.jl / mainReversi8x8.jl
# (e.g. main4IARow.jl:82).  Same exported names and call signatures as the reference module:
#   mcts(actor, visits, ngames, buffer; cpuct, noise)        self-play      (mcts_gpu.jl:477)
#   mcts(actor1, actor2, visits, ngames; cpuct, noise, conv) duel           (mcts_gpu.jl:581)
#   duelnetwork(actor1, actor2, visits, ngames, conv)                       (mcts_gpu.jl:653)
#   init / re_init / mcts_single                                            (mcts_gpu.jl:342-373, 376)
# NOT EXECUTED in this repository's CI: Julia is not installed in the build container or on the GPU box.
# The struct layouts and every entry point below are exercised byte for byte by the Python ctypes twin
# (alphagpu_b200/_lib.py, tests/test_gpu_parity.py).
module mcts_gpu

export mcts, duelnetwork

using ..Game      # Position, canPlay, play, isOver, VectorizedState, FeatureSize, maxActions, maxLengthGame, PoolSample …

const LIB = get(ENV, "ALPHAGPU_LIB", joinpath(@__DIR__, "..", "alphagpu_b200", "libalphagpu.so"))

# agpu_game ids (alphagpu.h).  The including main*.jl sets these three constants for its plugin:
#   main4IARow.jl: (0,0,0)   mainGobang.jl: (1,Main.N,Main.Nvict)   mainHex.jl: (2,Main.N,0)   mainReversi8x8.jl: (3,0,0)
const GAME_ID = Int32(get(ENV, "ALPHAGPU_GAME", "0") |> x -> parse(Int, x))
const GAME_N = Int32(isdefined(Main, :N) ? Main.N : 0)
const GAME_NVICT = Int32(isdefined(Main, :Nvict) ? Main.Nvict : 0)

struct AgpuConfig                 # agpu_config, 40 bytes
    game::Int32; n::Int32; nvict::Int32; rollouts::Int32
    max_games::Int64
    width::Int32; blocks::Int32; device::Int32; nn_mode::Int32
end

mutable struct AgpuSamples        # agpu_samples
    capacity::Int64; count::Int64
    state::Ptr{Int8}; policy::Ptr{Float32}; player::Ptr{Int8}; value::Ptr{Float32}; fstate::Ptr{Int8}
    game::Ptr{Int32}; ply::Ptr{Int32}
end

mutable struct AgpuRunStats       # agpu_run_stats
    sims::Int64; positions::Int64; plies::Int64; total_length::Int64; faults::Int64; kernel_launches::Int64
    device_ms::Float64; search_ms::Float64
    AgpuRunStats() = new(0, 0, 0, 0, 0, 0, 0.0, 0.0)
end

mutable struct Context
    h::Ptr{Cvoid}
    visits::Int
    ngames::Int
end

function check(ctx::Ptr{Cvoid}, rc::Cint)
    rc == 0 && return
    msg = unsafe_string(ccall((:agpu_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx))
    error("libalphagpu error $rc: $msg")
end

# `actor` is convert_back(net)::snetwork2 (DenseNet.jl:279-286): raw column-major Float32 arrays (host or CuArray -> Array)
function set_weights!(ctx::Context, actor, slot::Integer=0)
    base = Array(actor.base); res = [Array(w) for w in actor.res]
    pol = Array(actor.policy); polb = Array(actor.policy_bias); val = Array(actor.value); valb = Array(actor.value_bias)
    resptr = [pointer(w) for w in res]
    GC.@preserve base res pol polb val valb resptr begin
        check(ctx.h, ccall((:agpu_set_weights, LIB), Cint,
                           (Ptr{Cvoid}, Int32, Ptr{Float32}, Ptr{Ptr{Float32}}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
                           ctx.h, slot, base, resptr, pol, polb, val, valb))
    end
end

# init(positions, visits) (mcts_gpu.jl:350-357): allocates the tree arrays for length(positions) games on the GPU
function init(ngames::Integer, visits::Integer, actor; device=0, nn_mode=2)
    cfg = Ref(AgpuConfig(GAME_ID, GAME_N, GAME_NVICT, visits, ngames, size(actor.base, 1), length(actor.res), device, nn_mode))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:agpu_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Ptr{AgpuConfig}), h, cfg)
    rc == 0 || check(Ptr{Cvoid}(C_NULL), rc)
    ctx = Context(h[], visits, ngames)
    finalizer(c -> ccall((:agpu_destroy, LIB), Cvoid, (Ptr{Cvoid},), c.h), ctx)
    set_weights!(ctx, actor, 0)
    return ctx
end

# re_init(cu(positions), vnodes, L, …) (mcts_gpu.jl:368-373): positions is a Vector{Position} (isbits, 104 or 152 bytes each)
function re_init(ctx::Context, positions::Vector{Position})
    GC.@preserve positions check(ctx.h, ccall((:agpu_reinit, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{UInt32}),
                                              ctx.h, pointer(positions), length(positions), C_NULL))
end

# mcts_single(actor, visits, 256, vnodes, vnodesStats, leaf, newindex, L; training, cpuct, noise) (mcts_gpu.jl:376)
# returns (policy_final A×L, batch 2VS×L) as the reference leaves them in vnodesStats
function mcts_single(ctx::Context, visits, L; training=true, cpuct=2f0, noise=Float32(1 / maxActions), slot=0, seed=UInt64(0), ply=0)
    check(ctx.h, ccall((:agpu_search, LIB), Cint, (Ptr{Cvoid}, Int64, Int32, Int32, Int32, Float32, Float32, Ptr{Float32}, UInt64, UInt32),
                       ctx.h, L, slot, visits, training ? 1 : 0, cpuct, noise, C_NULL, seed, ply))
    policy = Array{Float32}(undef, maxActions, L); batch = Array{Float32}(undef, 2 * VectorizedState, L)
    check(ctx.h, ccall((:agpu_get_roots, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float32}, Ptr{Float32}), ctx.h, L, policy, batch))
    return policy, batch
end

# mcts(actor, visits, ngames, buffer; cpuct, noise) (mcts_gpu.jl:477-579): one generation of self-play, whole ply loop on the GPU;
# samples are pushed into `buffer` with push_buffer/update_buffer semantics (main4IARow.jl:49-75)
# Contexts are kept between calls — one per (ngames, visits, width, blocks, ngpus) — so that a generation does not pay for the allocation
# of the tree arrays again (the reference allocates per call, mcts_gpu.jl:481; its arrays are 1.6 GiB at 32768 games); only the weights
# are uploaded each time.  NGPUS[] > 1 drives that many devices from this one process through agpu_multi_* (the library runs one host
# thread per device and gathers the samples itself).
const NGPUS = Ref(1)
const CONTEXTS = Dict{NTuple{5,Int},Context}()
const MULTI = Dict{NTuple{5,Int},Ptr{Cvoid}}()
function cached_context(ngames, visits, actor)
    key = (Int(ngames), Int(visits), size(actor.base, 1), length(actor.res), 1)
    ctx = get!(CONTEXTS, key) do
        init(ngames, visits, actor)
    end
    set_weights!(ctx, actor, 0)
    return ctx
end
function cached_multi(ngames, visits, actor, ngpus)
    key = (Int(ngames), Int(visits), size(actor.base, 1), length(actor.res), Int(ngpus))
    h = get!(MULTI, key) do
        cfg = Ref(AgpuConfig(GAME_ID, GAME_N, GAME_NVICT, visits, ngames, size(actor.base, 1), length(actor.res), 0, 2))
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:agpu_multi_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Ptr{AgpuConfig}, Int32, Ptr{Int32}), out, cfg, ngpus, C_NULL)
        rc == 0 || error("agpu_multi_create: ", unsafe_string(ccall((:agpu_multi_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        out[]
    end
    return h
end
function multi_set_weights!(h, actor, slot)
    base, res = actor.base, actor.res
    pol, polb, val, valb = actor.policy, actor.policy_bias, actor.value, actor.value_bias
    resptr = [pointer(w) for w in res]
    GC.@preserve base res pol polb val valb resptr begin
        rc = ccall((:agpu_multi_set_weights, LIB), Cint,
                   (Ptr{Cvoid}, Int32, Ptr{Float32}, Ptr{Ptr{Float32}}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}),
                   h, slot, base, resptr, pol, polb, val, valb)
        rc == 0 || error("agpu_multi_set_weights: ", unsafe_string(ccall((:agpu_multi_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))
    end
end

# Page-locked sample arrays (agpu_host_alloc), one set per capacity, kept between generations: with page-locked destinations the library
# streams the rows of every ply out while the next ply searches and the final copies run at link speed (96 MB: 2 ms instead of 25).
const SAMPLE_ARRAYS = Dict{Int,Any}()
function pinned_array(::Type{T}, dims...) where {T}
    p = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:agpu_host_alloc, LIB), Cint, (Ptr{Ptr{Cvoid}}, UInt64), p, UInt64(max(1, prod(dims) * sizeof(T))))
    rc == 0 || error("agpu_host_alloc failed")
    unsafe_wrap(Array, Ptr{T}(p[]), dims; own=false)       # freed with agpu_host_free when the cache is dropped
end
function sample_arrays(cap)
    get!(SAMPLE_ARRAYS, cap) do
        (pinned_array(Int8, 2 * VectorizedState, cap), pinned_array(Float32, maxActions, cap), pinned_array(Int8, cap),
         pinned_array(Float32, cap), pinned_array(Int8, FeatureSize, cap))
    end
end

function mcts(actor, visits, ngames, buffer::Main.PoolSample; θ=1, cpuct=2.0, noise=Float32(1 / maxActions), seed=rand(UInt64), ngpus=NGPUS[])
    multi = ngpus > 1
    ctx = multi ? nothing : cached_context(ngames, visits, actor)
    mh = multi ? cached_multi(ngames, visits, actor, ngpus) : C_NULL
    multi && multi_set_weights!(mh, actor, 0)
    cap = ngames * maxLengthGame
    state, policy, player, value, fstate = sample_arrays(cap)
    smp = AgpuSamples(cap, 0, pointer(state), pointer(policy), pointer(player), pointer(value), pointer(fstate), C_NULL, C_NULL)
    results = zeros(Int64, 3); stats = AgpuRunStats()
    rc = GC.@preserve state policy player value fstate (multi ?
        ccall((:agpu_multi_selfplay, LIB), Cint,
              (Ptr{Cvoid}, Int32, Int32, Int64, UInt32, Float32, Float32, UInt64, Ref{AgpuSamples}, Ptr{Int64}, Ref{AgpuRunStats}),
              mh, 0, visits, ngames, 0, Float32(cpuct), Float32(noise), seed, smp, results, stats) :
        ccall((:agpu_selfplay, LIB), Cint,
              (Ptr{Cvoid}, Int32, Int32, Int64, UInt32, Float32, Float32, UInt64, Ref{AgpuSamples}, Ptr{Int64}, Ref{AgpuRunStats}),
              ctx.h, 0, visits, ngames, 0, Float32(cpuct), Float32(noise), seed, smp, results, stats))
    if rc == -5      # AGPU_ERR_ILLEGAL_MOVE == the reference's "faute" (mcts_gpu.jl:526-529)
        println("faute")
        return (data=[], valid=false)
    end
    multi ? (rc == 0 || error(unsafe_string(ccall((:agpu_multi_last_error, LIB), Cstring, (Ptr{Cvoid},), mh)))) : check(ctx.h, rc)
    for s in 1:smp.count          # push_buffer + update_buffer, ring semantics of main4IARow.jl:49-75
        index = buffer.currentIndex
        buffer.pool[index].state .= @view state[:, s]
        buffer.pool[index].policy .= @view policy[:, s]
        buffer.pool[index].player = player[s]
        buffer.pool[index].value = value[s]
        buffer.pool[index].fstate .= @view fstate[:, s]
        newindex = index == buffer.length ? 1 : index + 1
        newindex == 1 && (buffer.full = true)
        buffer.currentIndex = newindex
    end
    println("victoires,nul,défaites", results)
    println("longueur moyenne: ", stats.total_length / ngames)
    return (data=[], valid=true)
end

# mcts(actor1, actor2, visits, ngames; cpuct, noise, conv) (mcts_gpu.jl:581-651) -> [v, n, d]
function mcts(actor1, actor2, visits, ngames; cpuct=2f0, noise=Float32(1 / maxActions), conv=2, seed=rand(UInt64))
    ctx = cached_context(ngames, visits, actor1)
    set_weights!(ctx, actor2, 1)
    results = zeros(Int64, 3); stats = AgpuRunStats()
    rc = ccall((:agpu_duel, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Int32, Int64, UInt32, Float32, UInt64, Ptr{Int64}, Ref{AgpuRunStats}),
               ctx.h, 0, 1, visits, ngames, 0, Float32(cpuct), seed, results, stats)
    rc == -5 && println("faute")
    rc == -5 || check(ctx.h, rc)
    return results
end

# duelnetwork(actor1, actor2, visits, ngames, conv) (mcts_gpu.jl:653-668)
function duelnetwork(actor1, actor2, visits, ngames, conv=2)
    hngames = div(ngames, 2)
    println("net1 commence:")
    v1, n1, d1 = mcts(actor1, actor2, visits, hngames, conv=conv)
    println("v:$v1 n:$n1 d:$d1")
    println("net2 commence:")
    d2, n2, v2 = mcts(actor2, actor1, visits, hngames, conv=conv)
    println("v:$v2 n:$n2 d:$d2")
    return v1 + v2, n1 + n2, d1 + d2
end

end # module
