# mint_fixtures.jl - pins the oracle against the REFERENCE ITSELF (TEST INFRASTRUCTURE; cannot run in the authoring image: no Julia).
#
# Run by a maintainer who has Julia >= 1.9 with DeepQLearning.jl v0.7.1 (+ Flux 0.14, StatsBase, NPZ) installed:
#
#     julia --project=<env with DeepQLearning, Flux, NPZ> oracle/julia/mint_fixtures.jl
#
# For every committed golden case tests/golden/<name>_d?q?.npz it loads the INPUTS minted by tests/golden/make_golden.py (initial online /
# target parameters in Flux.params order and memory layout, the transitions, the sampled indices), rebuilds the reference's own objects
# (Flux Chain -> create_dueling_network, PrioritizedReplayBuffer + add_exp!), and evaluates exactly the statements of the reference's
# batch_train! (src/solver.jl:203-235) on get_batch(replay, idx) - the reference's sampler draws its own indices from MersenneTwister(0),
# which is why the statements are replayed here on the given indices instead of calling batch_train! itself.  Outputs go to
# tests/golden/ref_<name>_d?q?.npz; tests/test_golden_cpu.py::test_oracle_against_reference_fixtures and the GPU suite consume them when
# present, which lifts the "parity unpinned" status recorded in DESIGN.md section 5.
using DeepQLearning, Flux, NPZ, Random, StatsBase
import DeepQLearning: PrioritizedReplayBuffer, DQExperience, add_exp!, get_batch, update_priorities!, create_dueling_network,
                      huber_loss, globalnorm, flattenbatch

const ROOT = normpath(joinpath(@__DIR__, "..", ".."))
const GOLD = joinpath(ROOT, "tests", "golden")

# the networks of tests/util.py SPECS (activation codes: 0 identity, 1 relu, 2 tanh, 3 sigmoid)
act(code) = (identity, relu, tanh, sigmoid)[code + 1]
function chain_of(name)
    if name == "c1_gridworld"        # README.md:38
        return Chain(Dense(2, 32), Dense(32, 4)), (2,), 4, 32, 1000, 5f-3
    elseif name == "testmdp"         # test/runtests.jl:98
        return Chain(x -> flattenbatch(x), Dense(100, 8, tanh), Dense(8, 4)), (5, 5, 4), 4, 32, 500, 5f-3
    elseif name == "conv_small"
        return Chain(Conv((4, 4), 4 => 8, relu; stride=2), Conv((3, 3), 8 => 12, relu; stride=1), x -> flattenbatch(x),
                     Dense(12 * 3 * 3, 20, relu), Dense(20, 5)), (12, 12, 4), 5, 24, 300, 1f-3
    end
    error("unknown golden case $name")
end

# flat vector in Flux.params order and Julia memory layout -> the parameter arrays
function load_flat!(net, flat)
    o = 0
    for p in Flux.params(net)
        n = length(p)
        copyto!(p, reshape(flat[o+1:o+n], size(p)))
        o += n
    end
    @assert o == length(flat)
end
flat_of(ps) = vcat([vec(Array(p)) for p in ps]...)

# minimal environment carrying only what PrioritizedReplayBuffer(env, ...) asks for (observation size)
struct ShapeEnv <: DeepQLearning.CommonRLInterface.AbstractEnv
    o::Array{Float32}
end
DeepQLearning.CommonRLInterface.observe(e::ShapeEnv) = e.o

for path in filter(f -> endswith(f, ".npz") && !startswith(basename(f), "ref_"), readdir(GOLD; join=true))
    base = basename(path)[1:end-4]
    name, flags = rsplit(base, "_"; limit=2)
    dueling, double_q = flags[2] == '1', flags[4] == '1'
    g = npzread(path)
    model, oshape, nA, B, N, lr = chain_of(String(name))
    active_q = dueling ? create_dueling_network(model) : model
    target_q = deepcopy(active_q)
    load_flat!(active_q, g["theta0"]); load_flat!(target_q, g["theta_t"])
    # numpy (n, C, H, W) row-major == Julia (W, H, C, n) column-major: NPZ reverses nothing, so permute the axes back
    to_julia(a) = ndims(a) == 2 ? permutedims(a, (2, 1)) : permutedims(a, (4, 3, 2, 1))
    s_all, sp_all = to_julia(g["s"]), to_julia(g["sp"])
    deq(x) = eltype(x) == UInt8 ? Float32.(x) ./ 255f0 : Float32.(x)          # the u8 store stands for Float32(k)/255f0 (SURVEY F12)
    n = length(g["a"])
    replay = PrioritizedReplayBuffer(ShapeEnv(zeros(Float32, oshape...)), N, B)
    for i in 1:n
        si = deq(collect(selectdim(s_all, ndims(s_all), i))); spi = deq(collect(selectdim(sp_all, ndims(sp_all), i)))
        add_exp!(replay, DQExperience(si, Int32(g["a"][i]), Float32(g["r"][i]), spi, g["done"][i] != 0), abs(Float32(g["r"][i])))   # src/solver.jl:92
    end
    idx = Int.(g["idx"]) .+ 1
    s_batch, a_batch, r_batch, sp_batch, done_batch, indices, importance_weights = get_batch(replay, idx)     # PER.jl:89-104
    # ---- src/solver.jl:203-235, statement by statement ----
    p = Flux.params(active_q)
    loss_val = nothing; td_vals = nothing
    γ = 0.99f0
    if double_q
        qp_values = active_q(sp_batch)
        target_q_values = target_q(sp_batch)
        best_a = [CartesianIndex(argmax(qp_values[:, i]), i) for i = 1:B]
        q_sp_max = target_q_values[best_a]
    else
        best_a = [CartesianIndex(argmax(target_q(sp_batch)[:, i]), i) for i = 1:B]
        q_sp_max = dropdims(maximum(target_q(sp_batch), dims=1), dims=1)
    end
    q_targets = r_batch .+ (1f0 .- done_batch) .* γ .* q_sp_max
    q_all = active_q(s_batch)
    gs = Flux.gradient(p) do
        q_values = active_q(s_batch)
        q_sa = q_values[a_batch]
        td_vals = q_sa .- q_targets
        loss_val = sum(huber_loss, importance_weights .* td_vals)
        loss_val /= B
    end
    grad_norm = globalnorm(p, gs)
    grads = flat_of([gs[x] === nothing ? zero(x) : gs[x] for x in p])
    optimizer = Adam(lr)
    Flux.Optimise.update!(optimizer, p, gs)
    update_priorities!(replay, indices, td_vals)
    npzwrite(joinpath(GOLD, "ref_" * base * ".npz"), Dict(
        "q" => permutedims(q_all, (2, 1)), "y" => q_targets, "td" => td_vals, "w" => importance_weights,
        "best_a" => Int64[b[1] - 1 for b in best_a], "loss" => Float32[loss_val], "grad_norm" => Float32[grad_norm],
        "grads" => grads, "theta1" => flat_of(p), "prio1" => replay._priorities))
    println("minted ref_", base, ".npz  loss ", loss_val, "  grad_norm ", grad_norm)
end
