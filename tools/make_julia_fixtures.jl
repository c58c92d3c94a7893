#!/usr/bin/env julia
# Reference fixtures from the REAL DECAES.jl for the parity tests of libdecaes_cuda (tests/test_julia_fixtures.py).
#
# The build image of this repository has no Julia runtime, so its parity evidence is GPU-vs-CPU-oracle (a C restatement
# of the reference).  This script closes the loop on any machine that has Julia >= 1.9:
#
#     julia --project=/path/to/DECAES.jl tools/make_julia_fixtures.jl        # or: ] add DECAES@0.6.1 first
#
# It reads tests/golden/julia/cases.toml and the raw images next to it (written by
# tools/export_julia_fixture_inputs.py: little-endian Float64, [echo][voxel] = Array{Float64,4} of size (nvox,1,1,nTE)),
# runs T2mapSEcorr + T2partSEcorr exactly as test/cli.jl:309-348 does for its Julia-vs-CLI comparison, and writes
#     tests/golden/julia/<case>.out.f64     the maps listed under `keys`, concatenated in that order
#                                           (dist is nvox*nT2 values, [bin][voxel]; every other key nvox values)
#     tests/golden/julia/versions.toml      Julia / DECAES versions, CPU, thread count
# Commit the .out.f64 files: tests/test_julia_fixtures.py then checks the CPU oracle (always) and the CUDA library
# (-m gpu) against them with the north_star tolerances; without them those tests skip.
using DECAES, TOML, Dates

const ROOT = normpath(joinpath(@__DIR__, ".."))
const DIR = joinpath(ROOT, "tests", "golden", "julia")

function run_case(name::String, c::Dict, keys::Vector{String})
    nvox, nTE, nT2 = c["nvox"], c["nTE"], c["nT2"]
    image = Array{Float64, 4}(undef, nvox, 1, 1, nTE)
    read!(joinpath(DIR, name * ".image.f64"), image)
    kw = Dict{Symbol, Any}(
        :TE => c["TE"], :nT2 => nT2, :T2Range => (c["T2Range"][1], c["T2Range"][2]), :Reg => c["Reg"],
        :SaveRegParam => true, :SaveResidualNorm => true, :Silent => true, :Threaded => false,
    )
    for k in ("Chi2Factor", "NoiseLevel", "RefConAngle", "SetFlipAngle", "legacy", "nRefAngles", "nRefAnglesMin", "MinRefAngle", "T1", "Threshold")
        haskey(c, k) && (kw[Symbol(k)] = c[k])
    end
    maps, dist = T2mapSEcorr(image; kw...)
    part = T2partSEcorr(dist; T2Range = kw[:T2Range], SPWin = (c["SPWin"][1], c["SPWin"][2]),
                        MPWin = (c["MPWin"][1], c["MPWin"][2]), Silent = true, Threaded = false)
    open(joinpath(DIR, name * ".out.f64"), "w") do io
        for k in keys
            v = k == "dist" ? dist : haskey(maps, k) ? maps[k] : part[k]
            write(io, htol.(vec(Float64.(v))))
        end
    end
    return nothing
end

function main()
    cases = TOML.parsefile(joinpath(DIR, "cases.toml"))
    keys = Vector{String}(cases["keys"])
    for (name, c) in sort(collect(cases); by = first)
        c isa Dict || continue
        @info "case" name
        run_case(name, c, keys)
    end
    open(joinpath(DIR, "versions.toml"), "w") do io
        TOML.print(io, Dict(
            "julia" => string(VERSION), "decaes" => string(pkgversion(DECAES)), "cpu" => Sys.CPU_NAME,
            "threads" => Threads.nthreads(), "date" => string(Dates.now()),
        ))
    end
end

main()
