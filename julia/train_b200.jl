# train_b200.jl — drop-in for the batch loop of train.jl (traininPipe, train.jl:47-126): the forward/backward/ADAM step of a
# `networkf` runs in libalphagpu.so behind include/alphagpu_train.h.  Written against Julia 1.6 / Flux 0.12.6 like the reference;
# not executed in this repository (no Julia in the build image) — the Python twin alphagpu_b200/train.py binds the same symbols
# and is what the tests run.
module train_b200

using StatsBase: sample

const LIB = get(ENV, "ALPHAGPU_LIB", joinpath(@__DIR__, "..", "alphagpu_b200", "libalphagpu.so"))

struct AgpuTrainConfig           # alphagpu_train.h: agpu_train_config (80 bytes)
    device::Int32; in::Int32; width::Int32; blocks::Int32; actions::Int32; fsize::Int32; max_batch::Int32; reserved::Int32
    lr::Float64; beta1::Float64; beta2::Float64; eps::Float64; weight_decay::Float64
    feature_weight::Float32; reserved2::Float32
end

mutable struct Trainer
    h::Ptr{Cvoid}
    cfg::AgpuTrainConfig
end

function check(h, rc)
    rc == 0 && return
    msg = unsafe_string(ccall((:agpu_trainer_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
    error("libalphagpu trainer error $rc: $msg")
end

# net::networkf (DenseNet.jl:161-198) on the CPU (to_cpu(net)); Dense.weight is out x in, column-major Float32
weights(net) = (net.base.weight, [r.c1.weight for r in net.res], net.policy.weight, net.policy.bias, net.value.weight, net.value.bias,
                net.feature.weight, net.feature.bias)

function Trainer(net, batchsize; lr=0.001, device=0)
    base, res, pw, pb, vw, vb, fw, fb = weights(net)
    cfg = AgpuTrainConfig(device, size(base, 2), size(base, 1), length(res), size(pw, 1), size(fw, 1), batchsize, 0,
                          lr, 0.9, 0.999, 1e-8, 0.0001, 0.001f0, 0f0)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(C_NULL, ccall((:agpu_trainer_create, LIB), Cint, (Ptr{Ptr{Cvoid}}, Ref{AgpuTrainConfig}), h, Ref(cfg)))
    tr = Trainer(h[], cfg)
    finalizer(t -> ccall((:agpu_trainer_destroy, LIB), Cvoid, (Ptr{Cvoid},), t.h), tr)
    set_params!(tr, net)
    return tr
end

const W8 = (Ptr{Float32}, Ptr{Ptr{Float32}}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32}, Ptr{Float32})

function set_params!(tr::Trainer, net; reset_optimizer=true)
    base, res, pw, pb, vw, vb, fw, fb = weights(net)
    ptrs = [pointer(w) for w in res]
    GC.@preserve base res pw pb vw vb fw fb ptrs check(tr.h, ccall((:agpu_trainer_set_params, LIB), Cint, (Ptr{Cvoid}, W8..., Int32),
        tr.h, base, ptrs, pw, pb, vw, vb, fw, fb, reset_optimizer ? 1 : 0))
end

# writes the trained parameters back into the Flux arrays of `net` (then convert_back(net) feeds the search as before)
function get_params!(tr::Trainer, net)
    base, res, pw, pb, vw, vb, fw, fb = weights(net)
    ptrs = [pointer(w) for w in res]
    GC.@preserve base res pw pb vw vb fw fb ptrs check(tr.h, ccall((:agpu_trainer_get_params, LIB), Cint, (Ptr{Cvoid}, W8...),
        tr.h, base, ptrs, pw, pb, vw, vb, fw, fb))
    return net
end

# one custom_train! iteration (train.jl:128-162) on a batch in the sample layout: state Int8 (in x B), policy Float32 (A x B),
# value Float32 (B), fstate Int8 (FS x B) — column-major, i.e. sample-major in memory, as the C side wants
function train_step!(tr::Trainer, state::Matrix{Int8}, policy::Matrix{Float32}, value::Vector{Float32}, fstate::Matrix{Int8})
    loss = zeros(Float32, 4)
    GC.@preserve state policy value fstate loss check(tr.h, ccall((:agpu_trainer_step, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Int8}, Ptr{Float32}, Ptr{Float32}, Ptr{Int8}, Int64, Ptr{Float32}), tr.h, state, policy, value, fstate, size(state, 2), loss))
    return loss[1]
end

function loss_tot(tr::Trainer, state::Matrix{Int8}, policy::Matrix{Float32}, value::Vector{Float32}, fstate::Matrix{Int8})
    loss = zeros(Float32, 4)
    GC.@preserve state policy value fstate loss check(tr.h, ccall((:agpu_trainer_loss, LIB), Cint,
        (Ptr{Cvoid}, Ptr{Int8}, Ptr{Float32}, Ptr{Float32}, Ptr{Int8}, Int64, Ptr{Float32}), tr.h, state, policy, value, fstate, size(state, 2), loss))
    return loss[1]
end

# traininPipe(batchsize, net, p; ...) (train.jl:47-126) with the batch loop handed to the library; `net` is the CPU copy of trainingnet
function traininPipe(batchsize, net, p; in=98, out=7, fsize=1, epoch=1, lr=0.001, length_buffer)
    tr = Trainer(net, batchsize; lr=lr)                      # a fresh optimiser every call, as train.jl:50
    for i in 1:epoch
        println("epoque: ", i)
        L = length_buffer(p)
        q = sample(p.pool[1:L], min(2000000, L))
        L = div(length(q), batchsize)
        println("batch number: ", L)
        totloss = 0.0
        t = time()
        tmpx = zeros(Int8, in, batchsize); tmpy = zeros(Float32, out, batchsize); tmpf = zeros(Int8, fsize, batchsize); tmpr = zeros(Float32, batchsize)
        for (cpt, a) in enumerate(Iterators.partition(q, batchsize))
            cpt >= L && break
            for (k, sp) in enumerate(a)
                tmpx[:, k] .= sp.state; tmpy[:, k] .= sp.policy; tmpr[k] = sp.value; tmpf[:, k] .= sp.fstate
            end
            totloss += train_step!(tr, tmpx, tmpy, tmpr, tmpf)
        end
        println("total loss: ", totloss / (L - 1))
        println("training time :", time() - t)
    end
    return get_params!(tr, net)
end

end # module
