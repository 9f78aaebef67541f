# DeepQLearningB200.jl - reference-side binding of libdqn_b200.so (NOT executed in this repository: Julia is not
# installed in the build image; every behaviour behind these ccalls is exercised from Python ctypes, which binds
# the identical symbols - tests/test_capi_cpu.py, tests/test_gpu_parity.py).
#
# Drop-in scope: the methods below replace, inside JuliaPOMDP/DeepQLearning.jl,
#   PrioritizedReplayBuffer / add_exp! / update_priorities! / sample / get_batch   (src/prioritized_experience_replay.jl)
#   batch_train!(solver, env, policy, optimizer, target_q, replay)                 (src/solver.jl:191-236)
#   the target sync Flux.loadparams!(target_q, params(active_q))                   (src/solver.jl:142-145)
#   the acting forward policy.qnetwork(obatch)                                     (src/policy.jl:38-64)
# dqn_train!, solve, exploration, evaluation, logging and BSON checkpoints stay as they are and call these.
module DeepQLearningB200

using Flux

const LIB = get(ENV, "DQN_B200_LIB", "libdqn_b200.so")
const MAX_LAYERS = 16

struct LayerDesc          # dqn_layer_t
    kind::Int32; act::Int32; in::Int32; out::Int32; kh::Int32; kw::Int32; stride::Int32
end

mutable struct Config     # dqn_config_t (field order and types must match include/dqn_b200.h)
    abi_version::Int32; device::Int32
    obs_c::Int32; obs_h::Int32; obs_w::Int32; obs_dtype::Int32
    n_actions::Int32; n_layers::Int32
    layers::NTuple{MAX_LAYERS,LayerDesc}
    dueling::Int32; double_q::Int32; prioritized_replay::Int32; batch_size::Int32
    buffer_size::Int64
    alpha::Float32; beta::Float32; eps::Float32; learning_rate::Float32; discount::Float32
    adam_beta1::Float64; adam_beta2::Float64; adam_eps::Float64
    seed::UInt64
    math_mode::Int32; use_graph::Int32; rank::Int32; world::Int32
    nccl_id::NTuple{128,UInt8}
    max_act_rows::Int32
    trace_length::Int32; max_episode_length::Int32
    reserved::NTuple{5,Int32}
    Config() = new()
end

struct EngineError <: Exception
    code::Int32; msg::String
end

mutable struct Engine
    h::Ptr{Cvoid}
    batch_size::Int
    n_actions::Int
    nparams::Int
end

check(e::Engine, rc) = rc == 0 ? nothing :
    throw(EngineError(rc, unsafe_string(ccall((:dqn_last_error, LIB), Cstring, (Ptr{Cvoid},), e.h))))

act_code(f) = f === identity ? 0 : f === relu ? 1 : (f === tanh || f === Flux.tanh_fast) ? 2 : (f === σ || f === Flux.sigmoid_fast) ? 3 :
    error("DeepQLearningB200: unsupported activation $f")

function layer_desc(l)
    if l isa Dense
        LayerDesc(0, act_code(l.σ), size(l.weight, 2), size(l.weight, 1), 0, 0, 0)
    elseif l isa Conv
        kw, kh, cin, cout = size(l.weight)
        all(==(0), l.pad) || error("DeepQLearningB200: only pad = 0 convolutions")
        LayerDesc(1, act_code(l.σ), cin, cout, kh, kw, l.stride[1])
    elseif l isa Flux.Recur && l.cell isa Flux.LSTMCell        # Flux.LSTM(in, out): the recurrent engine (src/solver.jl:239-287)
        LayerDesc(3, 0, size(l.cell.Wi, 2), size(l.cell.Wh, 2), 0, 0, 0)
    else
        LayerDesc(2, 0, 0, 0, 0, 0, 0)      # flattenbatch / identity closures
    end
end

"""
    Engine(solver, env, qnetwork; discount, device=0)
Built where `solve` builds the buffer, the dueling split and the Adam optimiser (src/solver.jl:40-66).
`qnetwork` is the Chain handed to the solver (before create_dueling_network).
"""
function Engine(solver, env, qnetwork::Chain; discount, n_actions, obs_size, device=0, obs_u8=false, math_mode=1, seed=0)
    cfg = Config()
    ccall((:dqn_config_default, LIB), Cint, (Ref{Config},), cfg)
    cfg.device = device
    if length(obs_size) == 1
        cfg.obs_c, cfg.obs_h, cfg.obs_w = obs_size[1], 1, 1
    else
        cfg.obs_w, cfg.obs_h, cfg.obs_c = obs_size[1], obs_size[2], length(obs_size) >= 3 ? obs_size[3] : 1
    end
    cfg.obs_dtype = obs_u8 ? 1 : 0
    cfg.n_actions = n_actions
    descs = [layer_desc(l) for l in qnetwork.layers]
    cfg.n_layers = length(descs)
    cfg.layers = ntuple(i -> i <= length(descs) ? descs[i] : LayerDesc(0, 0, 0, 0, 0, 0, 0), MAX_LAYERS)
    cfg.dueling, cfg.double_q, cfg.prioritized_replay = solver.dueling, solver.double_q, solver.prioritized_replay
    cfg.batch_size, cfg.buffer_size = solver.batch_size, solver.buffer_size
    cfg.learning_rate, cfg.discount = solver.learning_rate, Float32(discount)
    cfg.math_mode, cfg.seed = math_mode, seed
    if solver.recurrence
        cfg.trace_length, cfg.max_episode_length = solver.trace_length, solver.max_episode_length
    end
    h = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:dqn_engine_create, LIB), Cint, (Ref{Config}, Ref{Ptr{Cvoid}}), cfg, h)
    rc == 0 || throw(EngineError(rc, unsafe_string(ccall((:dqn_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))))
    e = Engine(h[], solver.batch_size, n_actions, ccall((:dqn_num_params, LIB), Int64, (Ptr{Cvoid},), h[]))
    finalizer(x -> ccall((:dqn_engine_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.h), e)
    return e
end

# Flux.params(active_q) <-> engine: the flat image is the concatenation of the arrays as they lie in memory
flat(ps) = reduce(vcat, [vec(Float32.(p)) for p in ps])
set_params!(e::Engine, active_q; target=false) = (v = flat(Flux.params(active_q));
    check(e, ccall((:dqn_set_params, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float32}, Int64), e.h, target ? 1 : 0, v, length(v))))
function get_params!(active_q, e::Engine; target=false)
    v = Vector{Float32}(undef, e.nparams)
    check(e, ccall((:dqn_get_params, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float32}, Int64), e.h, target ? 1 : 0, v, length(v)))
    o = 0
    for p in Flux.params(active_q)
        copyto!(p, reshape(view(v, o+1:o+length(p)), size(p))); o += length(p)
    end
    active_q
end

# add_exp!(replay, exp, td_err)  src/prioritized_experience_replay.jl:65-74 (callers src/solver.jl:88-95)
function add_exp!(e::Engine, s::AbstractArray, a::Integer, r::Real, sp::AbstractArray, done::Bool, td_err::Real=abs(r))
    check(e, ccall((:dqn_replay_add, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Ptr{Float32}, Ptr{Cvoid}, Ptr{UInt8}, Ptr{Float32}, Int64),
                   e.h, s, Int32[a], Float32[r], sp, UInt8[done], Float32[td_err], 1))
end

# update_priorities!(replay, indices, td)  :76-80   (indices 1-based on the Julia side)
update_priorities!(e::Engine, indices::Vector{Int64}, td::Vector{Float32}) =
    check(e, ccall((:dqn_update_priorities, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Float32}, Int64), e.h, indices .- 1, td, length(td)))

# batch_train!(solver, env, policy, optimizer, target_q, replay) -> (loss_val, grad_norm)   src/solver.jl:191-236
function batch_train!(e::Engine)
    loss = Ref{Float32}(0); gn = Ref{Float32}(0)
    check(e, ccall((:dqn_train_step, LIB), Cint, (Ptr{Cvoid}, Ref{Float32}, Ref{Float32}), e.h, loss, gn))
    return loss[], gn[]
end

# One step ahead: the scalars are only logged (src/solver.jl:147-166), so the host may launch step k+1 before it reads step k's:
#   add_exp!(...); batch_train_async!(e); loss_k, gn_k = step_result(e, 1)
batch_train_async!(e::Engine) = check(e, ccall((:dqn_train_step_async, LIB), Cint, (Ptr{Cvoid},), e.h))
function step_result(e::Engine, back::Integer = 0)
    loss = Ref{Float32}(0); gn = Ref{Float32}(0)
    check(e, ccall((:dqn_step_result, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{Float32}, Ref{Float32}), e.h, back, loss, gn))
    return loss[], gn[]
end

# Flux.loadparams!(target_q, Flux.params(active_q))   src/solver.jl:142-145
sync_target!(e::Engine) = check(e, ccall((:dqn_sync_target, LIB), Cint, (Ptr{Cvoid},), e.h))

# policy.qnetwork(obatch)   src/policy.jl:38-64 ; obatch is (obs_dims..., n); returns (|A|, n)
function q_values(e::Engine, obatch::AbstractArray)
    n = size(obatch)[end]
    q = Matrix{Float32}(undef, e.n_actions, n)
    check(e, ccall((:dqn_q_values, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Int64, Ptr{Float32}), e.h, 0, obatch, n, q))
    return q
end

# ---- recurrent path: EpisodeReplayBuffer + recurrent batch_train!  (src/episode_replay.jl, src/solver.jl:239-287) --------------------
# add_episode!(r, ep) :60-66 - the host keeps collecting the running episode exactly as add_exp! :52-58 does and hands it over when it ends
function add_episode!(e::Engine, ep::Vector)
    s = reduce(hcat, [vec(Float32.(x.s)) for x in ep]); sp = reduce(hcat, [vec(Float32.(x.sp)) for x in ep])
    check(e, ccall((:dqn_episode_add, LIB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Ptr{Int32}, Ptr{Float32}, Ptr{Float32}, Ptr{UInt8}, Int64),
                   e.h, s, Int32[x.a for x in ep], Float32[x.r for x in ep], sp, UInt8[x.done for x in ep], length(ep)))
end
# resetstate!(policy)  src/policy.jl:32-34: the acting hidden state goes back to state0 (q_values carries it between calls)
resetstate!(e::Engine) = check(e, ccall((:dqn_policy_reset, LIB), Cint, (Ptr{Cvoid},), e.h))

# ---- vectorised acting / batched evaluation: argmax + epsilon-greedy on the device for n lanes  (src/solver.jl:83, src/policy.jl:38-46)
function act(e::Engine, obatch::AbstractArray; eps=0f0, call=0)
    n = size(obatch)[end]
    a = Vector{Int32}(undef, n)
    check(e, ccall((:dqn_act, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Float32, UInt64, Ptr{Int32}, Ptr{Float32}), e.h, obatch, n, eps, call, a, C_NULL))
    return a
end

# ---- all GPUs of the box from this one Julia process (dqn_group_*): ndev engines, one NCCL communicator inside -----------------------
mutable struct Group
    g::Ptr{Cvoid}
    ranks::Vector{Engine}
end
function Group(cfg::Config, ndev::Integer, batch_size, n_actions)
    g = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:dqn_group_create, LIB), Cint, (Ref{Config}, Cint, Ptr{Cint}, Ref{Ptr{Cvoid}}), cfg, ndev, C_NULL, g)
    rc == 0 || throw(EngineError(rc, unsafe_string(ccall((:dqn_group_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))))
    ranks = [Engine(ccall((:dqn_group_engine, LIB), Ptr{Cvoid}, (Ptr{Cvoid}, Cint), g[], r), batch_size, n_actions,
                    ccall((:dqn_num_params, LIB), Int64, (Ptr{Cvoid},), ccall((:dqn_group_engine, LIB), Ptr{Cvoid}, (Ptr{Cvoid}, Cint), g[], r))) for r in 0:ndev-1]
    grp = Group(g[], ranks)
    finalizer(x -> ccall((:dqn_group_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.g), grp)
    return grp
end
function batch_train!(grp::Group)          # one data-parallel step: every shard samples its own batch, one gradient all-reduce
    loss = Ref{Float32}(0); gn = Ref{Float32}(0)
    rc = ccall((:dqn_group_train_step, LIB), Cint, (Ptr{Cvoid}, Ref{Float32}, Ref{Float32}), grp.g, loss, gn)
    rc == 0 || throw(EngineError(rc, unsafe_string(ccall((:dqn_group_last_error, LIB), Cstring, (Ptr{Cvoid},), grp.g))))
    return loss[], gn[]
end

end # module
