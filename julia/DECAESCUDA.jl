# DECAESCUDA.jl — reference-side binding for libdecaes_cuda (see INTEGRATION.md).
#
# Loading this file after `using DECAES` adds GPU methods that replace the two worker loops
#   T2mapSEcorr!  : src/T2mapSEcorr.jl:177-193   (voxelwise_T2_distribution! per voxel)
#   T2partSEcorr  : src/T2partSEcorr.jl:58-68    (voxelwise_T2_parts! per voxel)
# with one blocking `ccall` each.  Everything before (options, NaN-filled outputs,
# src/T2mapSEcorr.jl:24-54) and after (Dict / Array conversion, :195) is the reference's own code,
# so the public API, option structs and output maps are unchanged.
#
# NOTE: there is no Julia runtime in the build image, so this file is not exercised by the test
# suite (tests/test_abi.py only checks that it names every field of the C structs).
module DECAESCUDA

using DECAES
using DECAES: T2mapOptions, T2partOptions, T2Maps, T2Distributions, T2Parts, is_B1map_provided

const libdecaes_cuda = get(ENV, "DECAES_CUDA_LIB", "libdecaes_cuda")

# Flat mirrors of include/decaes_cuda.h — field order and types must match exactly.
struct CT2mapOpts
    nx::Int32; ny::Int32; nz::Int32
    nTE::Int32
    nT2::Int32
    nRefAngles::Int32
    nRefAnglesMin::Int32
    reg::Int32
    legacy::Int32
    alpha_provided::Int32
    ngpus::Int32
    reserved::Int32
    TE::Float64
    T2min::Float64; T2max::Float64
    T1::Float64
    Threshold::Float64
    MinRefAngle::Float64
    RefConAngle::Float64
    Chi2Factor::Float64
    NoiseLevel::Float64
    SetFlipAngle::Float64
end

struct CT2partOpts
    nx::Int32; ny::Int32; nz::Int32
    nT2::Int32
    T2min::Float64; T2max::Float64
    SPWin_lo::Float64; SPWin_hi::Float64
    MPWin_lo::Float64; MPWin_hi::Float64
    Sigmoid::Float64
end

struct CT2mapOut
    gdn::Ptr{Float64}; ggm::Ptr{Float64}; gva::Ptr{Float64}; fnr::Ptr{Float64}; snr::Ptr{Float64}; alpha::Ptr{Float64}
    dist::Ptr{Float64}
    resnorm::Ptr{Float64}
    decaycurve::Ptr{Float64}
    mu::Ptr{Float64}; chi2factor::Ptr{Float64}
    decaybasis::Ptr{Float64}
    sfr::Ptr{Float64}; sgm::Ptr{Float64}; mfr::Ptr{Float64}; mgm::Ptr{Float64}
end

const REG_CODES = Dict("none" => 0, "lcurve" => 1, "gcv" => 2, "chi2" => 3, "mdp" => 4)
nan_if_nothing(x) = x === nothing ? NaN : Float64(x)
ptr_or_null(x::Nothing) = Ptr{Float64}(C_NULL)
ptr_or_null(x::Array{Float64}) = pointer(x)

function CT2mapOpts(o::T2mapOptions{Float64}; alpha_provided::Bool, ngpus::Int = 0)
    return CT2mapOpts(
        o.MatrixSize..., o.nTE, o.nT2, o.nRefAngles, o.nRefAnglesMin, REG_CODES[o.Reg], o.legacy, alpha_provided,
        ngpus, 0, o.TE, o.T2Range..., o.T1, o.Threshold, o.MinRefAngle, o.RefConAngle,
        nan_if_nothing(o.Chi2Factor), nan_if_nothing(o.NoiseLevel), nan_if_nothing(o.SetFlipAngle),
    )
end

function CT2partOpts(o::T2partOptions{Float64})
    return CT2partOpts(o.MatrixSize..., o.nT2, o.T2Range..., o.SPWin..., o.MPWin..., nan_if_nothing(o.Sigmoid))
end

# decaes_run_stats (ABI v2) — timings and the voxels the north star wants "counted and reported"
struct CRunStats
    voxels_total::Int64; voxels_processed::Int64
    ngpus_used::Int32; kernel_launches::Int32
    setup_ms::Float64; pipeline_ms::Float64; h2d_ms::Float64; d2h_ms::Float64; total_ms::Float64
    early_returns::Int64      # chi2 / MDP early-return voxels (reference-undefined: src/lsqnonneg.jl:510-515, 708-718)
    lcurve_overflow::Int64    # L-curve searches that outgrew the per-voxel caches (0 expected)
    nnls_itercap::Int64       # NNLS solves stopped by the 3n iteration cap
    pinned_staging::Int64     # 1: pageable arrays were staged through the library's pinned ring
end

"""
    last_stats() -> CRunStats

Statistics of the last library call made by this thread (`decaes_get_stats`).
"""
function last_stats()
    st = Ref{CRunStats}()
    ccall((:decaes_get_stats, libdecaes_cuda), Cvoid, (Ref{CRunStats},), st)
    return st[]
end

const DECAES_ECUDA = Cint(-2)        # no usable device / CUDA error
const DECAES_EUNSUPPORTED = Cint(-3) # outside the accelerated path: nT2 > 64, nRefAngles > 64, nTE > 96

last_error() = unsafe_string(ccall((:decaes_last_error, libdecaes_cuda), Cstring, ()))

# `true`: done on the GPU; `false`: the library declined (no device, unsupported size) and the caller should run
# DECAES.jl's own CPU method; anything else is an error.
function check_status(status::Cint)
    status == 0 && return true
    (status == DECAES_EUNSUPPORTED || status == DECAES_ECUDA) && return false
    return error("libdecaes_cuda failed with status $status: $(last_error())")
end

gpu_available() = ccall((:decaes_device_count, libdecaes_cuda), Cint, ()) > 0

"""
    t2map_gpu!(maps, dist, image, opts; ngpus = 0)

Drop-in body for the worker loop of `DECAES.T2mapSEcorr!` (src/T2mapSEcorr.jl:177-193).
"""
function t2map_gpu!(maps::T2Maps{Float64}, dist::T2Distributions{Float64}, image::Array{Float64, 4}, opts::T2mapOptions{Float64}; ngpus::Int = 0)
    @assert size(image) == (opts.MatrixSize..., opts.nTE)
    copts = Ref(CT2mapOpts(opts; alpha_provided = is_B1map_provided(maps), ngpus))
    decaybasis = opts.SetFlipAngle === nothing ? maps.decaybasis : nothing # shared basis is not per-voxel (:581-587)
    GC.@preserve image maps dist begin
        out = Ref(CT2mapOut(
            pointer(maps.gdn), pointer(maps.ggm), pointer(maps.gva), pointer(maps.fnr), pointer(maps.snr), pointer(maps.alpha),
            pointer(dist.distributions),
            ptr_or_null(maps.resnorm), ptr_or_null(maps.decaycurve), ptr_or_null(maps.mu), ptr_or_null(maps.chi2factor),
            ptr_or_null(decaybasis),
            Ptr{Float64}(C_NULL), Ptr{Float64}(C_NULL), Ptr{Float64}(C_NULL), Ptr{Float64}(C_NULL),
        ))
        status = ccall((:decaes_t2map, libdecaes_cuda), Cint,
            (Ptr{Float64}, Ref{CT2mapOpts}, Ptr{Cvoid}, Ref{CT2mapOut}),
            image, copts, C_NULL, out)
        if !check_status(status)
            # not accelerated (e.g. nT2 = 120, a 96-echo train, no GPU in this machine): the reference's CPU worker loop
            @warn "libdecaes_cuda declined ($(last_error())); running DECAES.jl's CPU path" maxlog = 1
            return cpu_t2map!(maps, dist, image, opts)
        end
    end
    st = last_stats()
    st.lcurve_overflow == 0 || @warn "L-curve cache overflow in $(st.lcurve_overflow) voxels (results approximate there)"
    opts.Silent || st.early_returns == 0 || @info "chi2/MDP early-return voxels (reference-undefined, src/lsqnonneg.jl:510-515, 708-718): $(st.early_returns)"
    return convert(Dict{String, Any}, maps), convert(Array{Float64, 4}, dist)
end

"""
    t2map_gpu(image::Array{Float32,4}, opts::T2mapOptions{Float64})

Float32 volume in (what NIfTI files hold), Float64 arithmetic and outputs: replaces `load_image`'s
`copyto!(Array{Float64,4}(undef, sz), data)` (src/main.jl:612-617) + `T2mapSEcorr` with half the host-to-device traffic
(`decaes_t2map_f32`; the conversion is exact and happens on the device).
"""
function t2map_gpu(image::Array{Float32, 4}, opts::T2mapOptions{Float64}; ngpus::Int = 0)
    @assert size(image) == (opts.MatrixSize..., opts.nTE)
    maps, dist = T2Maps(opts), T2Distributions(opts)
    copts = Ref(CT2mapOpts(opts; alpha_provided = false, ngpus))
    decaybasis = opts.SetFlipAngle === nothing ? maps.decaybasis : nothing
    GC.@preserve image maps dist begin
        out = Ref(CT2mapOut(
            pointer(maps.gdn), pointer(maps.ggm), pointer(maps.gva), pointer(maps.fnr), pointer(maps.snr), pointer(maps.alpha),
            pointer(dist.distributions),
            ptr_or_null(maps.resnorm), ptr_or_null(maps.decaycurve), ptr_or_null(maps.mu), ptr_or_null(maps.chi2factor),
            ptr_or_null(decaybasis),
            Ptr{Float64}(C_NULL), Ptr{Float64}(C_NULL), Ptr{Float64}(C_NULL), Ptr{Float64}(C_NULL),
        ))
        status = ccall((:decaes_t2map_f32, libdecaes_cuda), Cint,
            (Ptr{Float32}, Ref{CT2mapOpts}, Ptr{Cvoid}, Ref{CT2mapOut}), image, copts, C_NULL, out)
        check_status(status) || return DECAES.T2mapSEcorr(convert(Array{Float64, 4}, image), opts)
    end
    return convert(Dict{String, Any}, maps), convert(Array{Float64, 4}, dist)
end

# The reference's own generic method, reached past the GPU method below (no method piracy games: `invoke` names the
# signature of src/T2mapSEcorr.jl:151-156 explicitly).
cpu_t2map!(maps, dist, image, opts) =
    invoke(DECAES.T2mapSEcorr!, Tuple{T2Maps{T}, T2Distributions{T}, Array{T, 4}, T2mapOptions{T}} where {T}, maps, dist, image, opts)
cpu_t2part(T2distributions, opts) =
    invoke(DECAES.T2partSEcorr, Tuple{Array{T, 4}, T2partOptions{T}} where {T}, T2distributions, opts)

"""
    pinned_array(Float64, dims...)

An `Array` backed by page-locked memory from `decaes_host_alloc`: buffers allocated this way are copied to and from
the GPUs directly; ordinary Arrays work too (the library stages them through a pinned ring, about 1 % slower on one
GPU).  Free with `free_pinned(A)` once no Array aliases it.
"""
function pinned_array(::Type{T}, dims::Integer...) where {T}
    p = ccall((:decaes_host_alloc, libdecaes_cuda), Ptr{Cvoid}, (Csize_t,), prod(dims) * sizeof(T))
    p == C_NULL && error("decaes_host_alloc failed: $(last_error())")
    return unsafe_wrap(Array, Ptr{T}(p), dims; own = false)
end
free_pinned(A::Array) = ccall((:decaes_host_free, libdecaes_cuda), Cvoid, (Ptr{Cvoid},), pointer(A))

"""
    t2part_gpu(T2distributions, opts)

Drop-in body for the worker loop of `DECAES.T2partSEcorr` (src/T2partSEcorr.jl:58-68).
"""
function t2part_gpu(T2distributions::Array{Float64, 4}, opts::T2partOptions{Float64})
    @assert size(T2distributions) == (opts.MatrixSize..., opts.nT2)
    maps = T2Parts(opts)
    copts = Ref(CT2partOpts(opts))
    GC.@preserve T2distributions maps begin
        status = ccall((:decaes_t2part, libdecaes_cuda), Cint,
            (Ptr{Float64}, Ref{CT2partOpts}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
            T2distributions, copts, maps.sfr, maps.sgm, maps.mfr, maps.mgm)
        check_status(status) || return cpu_t2part(T2distributions, opts)
    end
    return convert(Dict{String, Any}, maps)
end

"""
    release()

Give back the device workspaces the library keeps between calls (`decaes_release`): basis tables, per-warp
scratch and the device copy of each GPU's voxel slab.
"""
release() = ccall((:decaes_release, libdecaes_cuda), Cvoid, ())

# Routing the public Float64 entry points through the GPU library is OPT-IN: `DECAESCUDA.enable!()` adds the two
# more specific methods below (inside DECAES itself the same thing is the two-line patch of INTEGRATION.md; adding
# methods to another package's functions at load time would break precompilation on Julia >= 1.10).  Calls the library
# does not accelerate - sizes beyond nT2 = 64 / nRefAngles = 64 / nTE = 96, a machine without a GPU - fall back to
# DECAES.jl's CPU worker loop through `invoke`, so nothing the reference handles starts to throw.
# `legacy = true` (sampled FITPACK spline for the flip angle and for the chi2 root, src/splines.jl:419-446,
# src/lsqnonneg.jl:595-636) runs on the GPU as well; an all-Float32 `T2mapSEcorr(image::Array{Float32,4})` keeps
# DECAES.jl's generic CPU method (Float32 arithmetic), `t2map_gpu(image32, opts64)` is the Float32-volume entry point.
function enable!()
    @eval DECAES.T2mapSEcorr!(maps::T2Maps{Float64}, dist::T2Distributions{Float64}, image::Array{Float64, 4}, opts::T2mapOptions{Float64}) =
        DECAESCUDA.t2map_gpu!(maps, dist, image, opts)
    @eval DECAES.T2partSEcorr(T2distributions::Array{Float64, 4}, opts::T2partOptions{Float64}) =
        DECAESCUDA.t2part_gpu(T2distributions, opts)
    return nothing
end

end # module
