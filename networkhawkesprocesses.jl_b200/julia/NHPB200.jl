# NHPB200.jl -- the reference-side binding a maintainer adds to NetworkHawkesProcesses.jl so that its
# event-history hot path runs in libnhp.so (include/nhp.h) on a B200.  The package API is unchanged:
# the methods below REDEFINE the hot-path methods of the existing types and forward to `ccall`.
#
# This file cannot be executed in the build image (no Julia toolchain, SURVEY.md section 8c); it is
# mechanically derived from include/nhp.h and kept minimal.  The Python mirror
# (networkhawkesprocesses.jl_b200/nhp_b200) exercises exactly the same entry points in the tests.
#
# Usage:   using NetworkHawkesProcesses; include("NHPB200.jl"); NHPB200.init!("/path/to/libnhp.so")
module NHPB200

using NetworkHawkesProcesses
import NetworkHawkesProcesses: loglikelihood, intensity, resample_parents, resample_adjacency_matrix!, convolve,
    ContinuousHawkesProcess, ContinuousStandardHawkesProcess, ContinuousNetworkHawkesProcess,
    DiscreteHawkesProcess, DiscreteStandardHawkesProcess, DiscreteNetworkHawkesProcess,
    ExponentialImpulseResponse, LogitNormalImpulseResponse, link_probability, ndims, basis

const LIB = Ref{String}("libnhp")
const CTX = Ref{Ptr{Cvoid}}(C_NULL)

function check(rc::Cint)
    rc == 0 && return
    msg = unsafe_string(ccall((:nhp_last_error, LIB[]), Cstring, (Ptr{Cvoid},), CTX[]))
    error("libnhp error $rc: $msg")          # the reference throws `error(...)` / DomainError at the same places
end

function init!(lib::String="libnhp"; device::Integer=0)
    LIB[] = lib
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:nhp_create, LIB[]), Cint, (Cint, Ref{Ptr{Cvoid}}), device, ctx)
    rc == 0 || error("nhp_create failed ($rc): " * unsafe_string(ccall((:nhp_last_error, LIB[]), Cstring, (Ptr{Cvoid},), C_NULL)))
    CTX[] = ctx[]
    atexit(() -> ccall((:nhp_destroy, LIB[]), Cint, (Ptr{Cvoid},), CTX[]))
    return nothing
end

# ---- device-resident data: upload once, reuse across the thousands of mle!/mcmc! evaluations -------------
mutable struct DeviceEvents
    h::Ptr{Cvoid}
    n::Int
    duration::Float64
end
const EVENT_CACHE = IdDict{Any,DeviceEvents}()   # keyed by the `events` vector object of `data`

function device_events(process, data)
    events, nodes, duration = data
    get!(EVENT_CACHE, events) do
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:nhp_events_upload, LIB[]), Cint,
            (Ptr{Cvoid}, Ptr{Float64}, Ptr{Int64}, Int64, Float64, Int64, Int64, Int64, Cint, Ref{Ptr{Cvoid}}),
            CTX[], events, Vector{Int64}(nodes), length(events), Float64(duration), ndims(process), 0, 0, 1, h))
        ev = DeviceEvents(h[], length(events), Float64(duration))
        finalizer(e -> ccall((:nhp_events_free, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), CTX[], e.h), ev)
        ev
    end
end

kind(::ExponentialImpulseResponse) = Cint(0)
kind(::LogitNormalImpulseResponse) = Cint(1)
p1(i::ExponentialImpulseResponse) = Matrix{Float64}(i.θ)
p1(i::LogitNormalImpulseResponse) = Matrix{Float64}(i.μ)
p2(i::ExponentialImpulseResponse) = C_NULL
p2(i::LogitNormalImpulseResponse) = Matrix{Float64}(i.τ)
adjacency(p::ContinuousStandardHawkesProcess) = C_NULL
adjacency(p::ContinuousNetworkHawkesProcess) = Matrix{Float64}(p.adjacency_matrix)   # Bool / Int64 / Float64 -> Float64

function push_params!(process::ContinuousHawkesProcess)
    K = ndims(process)
    check(ccall((:nhp_cont_params_set, LIB[]), Cint,
        (Ptr{Cvoid}, Cint, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64),
        CTX[], kind(process.impulses), K, Vector{Float64}(process.baseline.λ), Matrix{Float64}(process.weights.W),
        adjacency(process), p1(process.impulses), p2(process.impulses), Float64(process.impulses.Δtmax)))
end

# ---- continuous.jl:210 / :360 ---------------------------------------------------------------------------
function loglikelihood(process::ContinuousHawkesProcess, data; recursive=true)
    ev = device_events(process, data)
    push_params!(process)
    ll = Ref{Float64}(0.0)
    check(ccall((:nhp_cont_loglik, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ref{Float64}), CTX[], ev.h, recursive ? 1 : 0, ll))
    return ll[]
end

# ---- extension: objective + analytic gradient for mle! (Optim.only_fg!); layout of params(process), continuous.jl:116-119
function loglikelihood_gradient(process::ContinuousStandardHawkesProcess, data; recursive=true)
    ev = device_events(process, data)
    push_params!(process)
    K = ndims(process)
    ll = Ref{Float64}(0.0)
    dλ = Vector{Float64}(undef, K); dW = Matrix{Float64}(undef, K, K); d1 = Matrix{Float64}(undef, K, K); d2 = Matrix{Float64}(undef, K, K)
    check(ccall((:nhp_cont_loglik_grad, LIB[]), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ref{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        CTX[], ev.h, recursive ? 1 : 0, ll, dλ, dW, d1, d2))
    impulse_grad = process.impulses isa ExponentialImpulseResponse ? vec(d1) : [vec(d1); vec(d2)]
    return ll[], [dλ; impulse_grad; vec(dW)]
end

# ---- continuous.jl:76 / :84 -----------------------------------------------------------------------------
function intensity(process::ContinuousHawkesProcess, data, times::Vector{Float64})
    ev = device_events(process, data)
    push_params!(process)
    out = Matrix{Float64}(undef, length(times), ndims(process))          # column-major [q + nq*k]
    check(ccall((:nhp_cont_intensity, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}),
        CTX[], ev.h, times, length(times), out))
    return out
end
intensity(process::ContinuousHawkesProcess, data, time::Float64) = vec(intensity(process, data, [time]))

# ---- parents.jl:1 ----------------------------------------------------------------------------------------
const SWEEP = Ref{UInt64}(0)
function resample_parents(process::ContinuousHawkesProcess, data)
    ev = device_events(process, data)
    push_params!(process)
    parents = Vector{Int64}(undef, ev.n); parentnodes = Vector{Int64}(undef, ev.n)
    SWEEP[] += 1
    check(ccall((:nhp_cont_resample_parents, LIB[]), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, UInt64, UInt64, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}),
        CTX[], ev.h, rand(UInt64), SWEEP[], C_NULL, parents, parentnodes))
    return parents, parentnodes
end

# ---- counters / impulse statistics (parents.jl:61,70; baselines.jl:87; impulses.jl:84,230,242) -----------
# one fused call replaces node_counts x2, parent_counts, duration_mean / log_duration_sum / log_duration_variation
function gibbs_statistics(process::ContinuousHawkesProcess, data)
    ev = device_events(process, data)
    K = ndims(process)
    M0 = zeros(K); Mn = zeros(K); Mnm = zeros(K, K); S1 = zeros(K, K); S2 = zeros(K, K)
    check(ccall((:nhp_cont_suffstats, LIB[]), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), CTX[], ev.h, M0, Mn, Mnm, S1, S2))
    return (M0=M0, Mn=Mn, Mnm=Mnm, S1=S1, S2=S2)     # Xnm = S1 ./ Mnm; Vnm = S2 (LogitNormal); duration_mean = fillna!(S1 ./ Mnm, 0)
end

# ---- continuous.jl:444 ----------------------------------------------------------------------------------
function resample_adjacency_matrix!(process::ContinuousNetworkHawkesProcess, data)
    ev = device_events(process, data)
    push_params!(process)
    A = Matrix{Float64}(process.adjacency_matrix)
    rho = Matrix{Float64}(link_probability(process.network))
    SWEEP[] += 1
    check(ccall((:nhp_cont_resample_adjacency, LIB[]), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, UInt64, UInt64, Ptr{Float64}, Ptr{Float64}), CTX[], ev.h, rho, rand(UInt64), SWEEP[], C_NULL, A))
    process.adjacency_matrix .= A
    return nothing
end

# ---- optional: the conjugate draws of resample!(process, data) (continuous.jl:202-208) on the device ---------------
# baselines.jl:72-77, weights.jl:59-64, impulses.jl:68-73 / 204-214; the process's arrays are refreshed from the device.
function resample_on_device!(process::ContinuousHawkesProcess, data)
    ev = device_events(process, data)
    push_params!(process)
    SWEEP[] += 1
    seed = rand(UInt64)
    check(ccall((:nhp_cont_resample_parents, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, UInt64, UInt64, Ptr{Float64}, Ptr{Int64}, Ptr{Int64}),
        CTX[], ev.h, seed, SWEEP[], C_NULL, C_NULL, C_NULL))
    b, w, imp = process.baseline, process.weights, process.impulses
    hyper = imp isa ExponentialImpulseResponse ? Float64[b.α0, b.β0, w.κ, w.ν, imp.α, imp.β] :
                                                  Float64[b.α0, b.β0, w.κ, w.ν, imp.μμ, imp.κμ, imp.α0, imp.β0]
    check(ccall((:nhp_cont_resample_params, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, UInt64, UInt64, Float64, Ptr{Float64}, Cint, Cint),
        CTX[], ev.h, seed, SWEEP[], Float64(data[3]), hyper, length(hyper), 1))
    K = ndims(process)
    λ = Vector{Float64}(undef, K); W = Matrix{Float64}(undef, K, K); q1 = Matrix{Float64}(undef, K, K); q2 = Matrix{Float64}(undef, K, K)
    check(ccall((:nhp_cont_params_get, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        CTX[], λ, W, C_NULL, q1, imp isa ExponentialImpulseResponse ? C_NULL : q2))
    b.λ .= λ; w.W .= W
    if imp isa ExponentialImpulseResponse; imp.θ .= q1 else imp.μ .= q1; imp.τ .= q2 end
    return nothing
end

# ---- discrete path (discrete.jl:86-151; parents.jl:82-177) ------------------------------------------------
mutable struct DeviceCounts
    h::Ptr{Cvoid}
end
const COUNT_CACHE = IdDict{Any,DeviceCounts}()
function device_counts(data::Matrix{Int64})
    get!(COUNT_CACHE, data) do
        h = Ref{Ptr{Cvoid}}(C_NULL)
        N, T = size(data)
        check(ccall((:nhp_disc_upload, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Int64, Int64, Ref{Ptr{Cvoid}}), CTX[], data, N, T, 0, h))
        d = DeviceCounts(h[])
        finalizer(x -> ccall((:nhp_disc_free, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), CTX[], x.h), d)
        d
    end
end

function convolve(process::DiscreteHawkesProcess, data)
    d = device_counts(Matrix{Int64}(data))
    ϕ = hcat(basis(process.impulses)...)                     # L x B, column-major phi[l + L*b]
    N, T = size(data); L, B = size(ϕ)
    out = Array{Float64,3}(undef, T, N, B)                   # conv[t + T*(n + N*b)]
    check(ccall((:nhp_disc_convolve, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Ptr{Float64}), CTX[], d.h, ϕ, L, B, out))
    return out
end

adjacency(p::DiscreteStandardHawkesProcess) = C_NULL
adjacency(p::DiscreteNetworkHawkesProcess) = Matrix{Float64}(p.adjacency_matrix)
function push_params!(process::DiscreteHawkesProcess)
    N = ndims(process); B = size(process.impulses.θ, 3)
    check(ccall((:nhp_disc_params_set, LIB[]), Cint,
        (Ptr{Cvoid}, Int64, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64),
        CTX[], N, B, Vector{Float64}(process.baseline.λ), Matrix{Float64}(process.weights.W), adjacency(process),
        Array{Float64,3}(process.impulses.θ), process.dt))
end

# loglikelihood(process, data, convolved): `convolved` is the device-resident result of `convolve` on the same data
function loglikelihood(process::DiscreteHawkesProcess, data, convolved)
    d = device_counts(Matrix{Int64}(data))
    push_params!(process)
    ll = Ref{Float64}(0.0)
    check(ccall((:nhp_disc_loglik, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Float64}), CTX[], d.h, ll))
    return ll[]
end

# reduced form of resample_parents(process, data, convolved): counts[c, k] = sum_t parents[t, c, k]
function resample_parent_counts(process::DiscreteHawkesProcess, data)
    d = device_counts(Matrix{Int64}(data))
    push_params!(process)
    N = ndims(process); B = size(process.impulses.θ, 3)
    counts = zeros(N, 1 + N * B)
    SWEEP[] += 1
    check(ccall((:nhp_disc_gibbs_counts, LIB[]), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, UInt64, UInt64, Ptr{Float64}, Int64, Ptr{Float64}),
        CTX[], d.h, rand(UInt64), SWEEP[], C_NULL, 0, counts))
    return counts
end

end # module
