# RayTracingWeekendB200.jl -- the thin `ccall` shim a RayTracingWeekend.jl user loads to run the package's hot path
# (render -> ray_color -> hit/scatter) on B200 GPUs through librtw_b200.so (C-ABI: include/rtw_b200.h).
#
# Everything host-side stays the reference's own Julia code: Camera/default_camera (src/camera.jl), Sphere, the
# materials, HittableList (src/structs.jl, src/material.jl) and the scene builders (src/scenes.jl).  This module
# adds one method, `render_b200`, with the positional signature of the reference's
#     render(scene::HittableList, cam::Camera{T}, image_width=400, n_samples=1)          src/render.jl:8-9
# and `RayTracingWeekendB200.install!()` makes it THE method: it defines `RayTracingWeekend.render` for
# Camera{Float32} / Camera{Float64}, so existing scripts (`render(scene_random_spheres(; elem_type=Float32), cam, 1920,
# 1000)`) run on the GPUs unchanged.  Opt-in, because it replaces a method of another package.
#
# NOTE: Julia is not installed in the build image, so this file is exercised only where `julia` exists; the
# Python package next to it (api.py) makes the identical sequence of C-ABI calls and is what the tests drive.
module RayTracingWeekendB200

using RayTracingWeekend
using RayTracingWeekend: Camera, HittableList, Sphere, Lambertian, Metal, Dielectric
using Images: RGB
using StaticArrays: SA

const librtw = get(ENV, "RTW_B200_LIB", joinpath(@__DIR__, "..", "csrc", "librtw_b200.so"))

const RTW_LAMBERTIAN = UInt32(0)
const RTW_METAL      = UInt32(1)
const RTW_DIELECTRIC = UInt32(2)

"mirror of `rtw_stats` (include/rtw_b200.h)"
struct RtwStats
    paths::UInt64
    ray_segments::UInt64
    sphere_tests::UInt64
    n_spheres::UInt32
    image_width::Int32
    image_height::Int32
    rows_rendered::Int32
    kernel_launches::Int32
    ms_total::Float32
    ms_trace::Float32
    ms_resolve::Float32
    ms_h2d::Float32
    ms_d2h::Float32
    n_devices::Int32            # ABI v3
    reserved0::Int32
    grid_fallback_rays::UInt64
    grid_loose_cells::UInt64
    grid_cells::UInt64
    grid_tests::UInt64
end

const RTW_ABI_VERSION = 3

# The ABI passes Julia's own structs: check the layouts once, when the module loads (include/rtw_b200.h:
# rtw_camera = 22 x f32, rtw_camera_f64 = 22 x f64, rtw_stats = 104 bytes) and that the library speaks this ABI.
function __init__()
    @assert isbitstype(Camera{Float32}) && sizeof(Camera{Float32}) == 88 "Camera{Float32} is not the 22 x Float32 the ABI expects"
    @assert isbitstype(Camera{Float64}) && sizeof(Camera{Float64}) == 176 "Camera{Float64} is not the 22 x Float64 the ABI expects"
    @assert fieldnames(Camera{Float32}) == (:origin, :lower_left_corner, :horizontal, :vertical, :u, :v, :w, :lens_radius)
    @assert sizeof(RtwStats) == 104 "RtwStats does not match rtw_stats"
    @assert sizeof(RGB{Float32}) == 12 && sizeof(RGB{Float64}) == 24
    v = ccall((:rtw_abi_version, librtw), Cint, ())
    v == RTW_ABI_VERSION || error("librtw_b200.so has ABI version $v, this shim was written for $RTW_ABI_VERSION")
end

mutable struct Context
    ptr::Ptr{Cvoid}
end

function check(ctx::Ptr{Cvoid}, status::Cint)
    status == 0 && return
    msg = ctx == C_NULL ? "status $status" :
          unsafe_string(ccall((:rtw_last_error, librtw), Cstring, (Ptr{Cvoid},), ctx))
    error("rtw_b200 error $status: $msg")   # no CPU fallback on the hot path
end

"""
    Context(devices = [0])

Owns streams and device buffers on the given CUDA devices (rtw_create); rows are split over all of them.
"""
function Context(devices::Vector{<:Integer} = [0])
    ids = Cint.(devices)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(C_NULL, ccall((:rtw_create, librtw), Cint, (Ptr{Cint}, Cint, Ptr{Ptr{Cvoid}}), ids, length(ids), out))
    ctx = Context(out[])
    finalizer(c -> (c.ptr != C_NULL && ccall((:rtw_destroy, librtw), Cint, (Ptr{Cvoid},), c.ptr); c.ptr = C_NULL), ctx)
    ctx
end

# rtw_set_option: e.g. `set_option!(ctx, RTW_OPT_MODE, RTW_MODE_GRID)` renders the same image bits through a uniform grid
# (2.3x faster on scene_random_spheres, ~140x on 100k spheres); the default is the reference's linear sweep
const RTW_OPT_MODE            = Cint(1)   # RTW_MODE_*
const RTW_OPT_BLOCKS_PER_SM   = Cint(3)
const RTW_OPT_COLLECT_TIMING  = Cint(4)
const RTW_OPT_RAYS_PER_LANE   = Cint(5)   # options 5-9 select kernel variants: need a library built with RTW_BUILD_VARIANTS=1
const RTW_OPT_SWEEP           = Cint(6)
const RTW_OPT_COOP            = Cint(7)
const RTW_OPT_TAIL            = Cint(8)
const RTW_OPT_WALK            = Cint(9)
const RTW_OPT_GATHER          = Cint(10)  # RTW_GATHER_*: framebuffer gather of a multi-device context
const RTW_OPT_SMALL_RENDER    = Cint(11)  # 1 (default): small renders take the single-launch latency path
const RTW_MODE_FUSED          = 0         # the reference's linear sweep (default, the benchmarked path)
const RTW_MODE_WAVEFRONT      = 1         # separate raygen / intersect / shade / accumulate kernels
const RTW_MODE_CTA_WAVEFRONT  = 2         # (RTW_BUILD_VARIANTS=1)
const RTW_MODE_GRID           = 3         # uniform-grid traversal, same image bits for every list size
const RTW_GATHER_PEER         = 0
const RTW_GATHER_NCCL         = 1
set_option!(ctx::Context, option::Integer, value::Integer) =
    check(ctx.ptr, ccall((:rtw_set_option, librtw), Cint, (Ptr{Cvoid}, Cint, Int64), ctx.ptr, option, value))

const DEFAULT_CTX = Ref{Union{Nothing,Context}}(nothing)
default_context() = (DEFAULT_CTX[] === nothing && (DEFAULT_CTX[] = Context()); DEFAULT_CTX[])

# ---- the one piece of glue: Sphere.mat is an abstract field (src/structs.jl:34), so a HittableList is not isbits
#      and has to be flattened into the SoA arrays of rtw_set_scene.  List order is kept (tie-break, src/hit.jl:44-46).
matrow(m::Lambertian{Float32}) = (RTW_LAMBERTIAN, (m.albedo[1], m.albedo[2], m.albedo[3], 0f0))
matrow(m::Metal{Float32})      = (RTW_METAL,      (m.albedo[1], m.albedo[2], m.albedo[3], m.fuzz))
matrow(m::Dielectric{Float32}) = (RTW_DIELECTRIC, (1f0, 1f0, 1f0, m.ir))
matrow(m) = error("rtw_b200: unsupported material $(typeof(m)) (Float32 Lambertian/Metal/Dielectric only)")

function flatten(scene::HittableList)
    n = length(scene)
    geom = Matrix{Float32}(undef, 4, n)   # column k = {cx, cy, cz, radius}
    mat  = Matrix{Float32}(undef, 4, n)   # column k = {albedo r,g,b, fuzz|ir|0}
    kind = Vector{UInt32}(undef, n)
    for (k, s) in enumerate(scene)
        s isa Sphere{Float32} || error("rtw_b200: only Sphere{Float32} hittables are supported, got $(typeof(s))")
        geom[1, k], geom[2, k], geom[3, k] = s.center
        geom[4, k] = s.radius             # sign kept: a negative radius is a hollow glass shell (src/scenes.jl:35-36)
        kind[k], row = matrow(s.mat)
        mat[:, k] .= row
    end
    geom, mat, kind
end

"""
    render_b200(scene, cam::Camera{Float32}, image_width=400, n_samples=1; max_depth=16, seed=1, ctx, stats)

Drop-in for `render(scene, cam, image_width, n_samples)` (src/render.jl:8-44).  Returns `Matrix{RGB{Float32}}` of
size (image_width ÷ 16//9, image_width), gamma-2 encoded and unclamped exactly like the reference.
`max_depth` is `ray_color`'s `depth` (hard default 16 in the reference, src/ray_color.jl:14); `seed` plays the role
of `reseed!()` (src/render.jl:21): the same seed gives the same image on every call, for any number of GPUs.
"""
function render_b200(scene::HittableList, cam::Camera{Float32}, image_width::Integer = 400, n_samples::Integer = 1;
                     max_depth::Integer = 16, seed::Integer = 1, ctx::Context = default_context(),
                     stats::Union{Nothing,Ref{RtwStats}} = nothing)
    geom, mat, kind = flatten(scene)
    H = Int(ccall((:rtw_image_height, librtw), Cint, (Cint,), image_width))
    img = Matrix{RGB{Float32}}(undef, H, image_width)      # column-major H x W of 3 x Float32: the ABI's out_rgb layout
    st = stats === nothing ? Ref{RtwStats}() : stats
    camref = Ref(cam)                                      # Camera{Float32} is isbits: 22 x Float32 = rtw_camera
    GC.@preserve geom mat kind img camref begin
        status = ccall((:rtw_render_scene, librtw), Cint,
                       (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{UInt32}, UInt32, Ptr{Cvoid}, Cint, Cint, Cint, UInt64,
                        Ptr{Cvoid}, Ptr{RtwStats}),
                       ctx.ptr, geom, mat, kind, length(kind), camref, image_width, n_samples, max_depth, seed,
                       img, st)   # img::Matrix{RGB{Float32}} is H*W*3 contiguous Float32: exactly the ABI's out_rgb
        check(ctx.ptr, status)
    end
    img
end

# ---- Camera{Float64}: the Float64 instantiation of the kernels (rtw_render_scene_f64) ---------------------------
matrow64(m::Lambertian{Float64}) = (RTW_LAMBERTIAN, (m.albedo[1], m.albedo[2], m.albedo[3], 0.0))
matrow64(m::Metal{Float64})      = (RTW_METAL,      (m.albedo[1], m.albedo[2], m.albedo[3], m.fuzz))
matrow64(m::Dielectric{Float64}) = (RTW_DIELECTRIC, (1.0, 1.0, 1.0, m.ir))
matrow64(m) = error("rtw_b200: unsupported material $(typeof(m)) in a Float64 scene")

function flatten64(scene::HittableList)
    n = length(scene)
    geom, mat, kind = Matrix{Float64}(undef, 4, n), Matrix{Float64}(undef, 4, n), Vector{UInt32}(undef, n)
    for (k, s) in enumerate(scene)
        s isa Sphere{Float64} || error("rtw_b200: a Float64 camera needs Sphere{Float64} hittables, got $(typeof(s))")
        geom[1, k], geom[2, k], geom[3, k] = s.center
        geom[4, k] = s.radius
        kind[k], row = matrow64(s.mat)
        mat[:, k] .= row
    end
    geom, mat, kind
end

function render_b200(scene::HittableList, cam::Camera{Float64}, image_width::Integer = 400, n_samples::Integer = 1;
                     max_depth::Integer = 16, seed::Integer = 1, ctx::Context = default_context(),
                     stats::Union{Nothing,Ref{RtwStats}} = nothing)
    geom, mat, kind = flatten64(scene)
    H = Int(ccall((:rtw_image_height, librtw), Cint, (Cint,), image_width))
    img = Matrix{RGB{Float64}}(undef, H, image_width)
    st = stats === nothing ? Ref{RtwStats}() : stats
    camref = Ref(cam)                                      # Camera{Float64} is isbits: 22 x Float64 = rtw_camera_f64
    GC.@preserve geom mat kind img camref begin
        check(ctx.ptr, ccall((:rtw_render_scene_f64, librtw), Cint,
                             (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{UInt32}, UInt32, Ptr{Cvoid}, Cint, Cint, Cint, UInt64,
                              Ptr{Cvoid}, Ptr{RtwStats}),
                             ctx.ptr, geom, mat, kind, length(kind), camref, image_width, n_samples, max_depth, seed, img, st))
    end
    img
end

# ---- progressive rendering, image and scene files (C-ABI v2) ----------------------------------------------------
# render() split into passes over the samples: every draw is addressed by (pixel, sample, event) and the accumulator
# is an integer sum, so accumulate!(…, 0, a, n) + accumulate!(…, a, n - a, n) gives the bits of render(…, n).

set_scene!(ctx::Context, scene::HittableList) = begin
    geom, mat, kind = flatten(scene)
    GC.@preserve geom mat kind check(ctx.ptr, ccall((:rtw_set_scene, librtw), Cint,
        (Ptr{Cvoid}, Ptr{Float32}, Ptr{Float32}, Ptr{UInt32}, UInt32), ctx.ptr, geom, mat, kind, length(kind)))
end

"adds samples `first .. first+count-1` (0-based) of every pixel; `first == 0` starts a new image of `total` samples"
function accumulate!(ctx::Context, cam::Camera{Float32}, image_width::Integer, first::Integer, count::Integer,
                     total::Integer; max_depth::Integer = 16, seed::Integer = 1)
    st = Ref{RtwStats}()
    camref = Ref(cam)
    GC.@preserve camref check(ctx.ptr, ccall((:rtw_accumulate, librtw), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, Cint, Cint, UInt64, Ptr{RtwStats}),
        ctx.ptr, camref, image_width, first, count, total, max_depth, seed, st))
    st[]
end

"the image of the samples accumulated so far, as `render` returns it"
function resolve(ctx::Context)
    w, done, total = Ref{Cint}(0), Ref{Cint}(0), Ref{Cint}(0)
    check(ctx.ptr, ccall((:rtw_progress, librtw), Cint, (Ptr{Cvoid}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}), ctx.ptr, w, done, total))
    H = Int(ccall((:rtw_image_height, librtw), Cint, (Cint,), w[]))
    img = Matrix{RGB{Float32}}(undef, H, Int(w[]))
    GC.@preserve img check(ctx.ptr, ccall((:rtw_resolve, librtw), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), ctx.ptr, img))
    img
end

"`save(path, img)` without Images.jl's writers: 8-bit PNG of the progressive image (clamp01nan + N0f8 rounding)"
function save_png(ctx::Context, path::AbstractString)
    w, done, total = Ref{Cint}(0), Ref{Cint}(0), Ref{Cint}(0)
    check(ctx.ptr, ccall((:rtw_progress, librtw), Cint, (Ptr{Cvoid}, Ptr{Cint}, Ptr{Cint}, Ptr{Cint}), ctx.ptr, w, done, total))
    H = Int(ccall((:rtw_image_height, librtw), Cint, (Cint,), w[]))
    rgb8 = Vector{UInt8}(undef, 3 * H * Int(w[]))     # row-major, top row first
    GC.@preserve rgb8 begin
        check(ctx.ptr, ccall((:rtw_resolve_rgb8, librtw), Cint, (Ptr{Cvoid}, Ptr{UInt8}), ctx.ptr, rgb8))
        check(C_NULL, ccall((:rtw_write_png, librtw), Cint, (Cstring, Ptr{UInt8}, Cint, Cint), path, rgb8, w[], H))
    end
    path
end

"the progressive image of `ctx` as a self-describing checkpoint file (sums + seed, depth, camera, scene hash + CRC-32)"
save_checkpoint(ctx::Context, path::AbstractString) =
    (check(ctx.ptr, ccall((:rtw_checkpoint_save, librtw), Cint, (Ptr{Cvoid}, Cstring), ctx.ptr, path)); path)

"resume from a checkpoint file: `set_scene!` the scene it was rendered from first, then `accumulate!` from its sample on"
load_checkpoint!(ctx::Context, path::AbstractString) =
    check(ctx.ptr, ccall((:rtw_checkpoint_load, librtw), Cint, (Ptr{Cvoid}, Cstring), ctx.ptr, path))

"writes the flattened HittableList as a `.rtwscene` file (the fixture format shared with the Python harness / oracle)"
function save_scene(path::AbstractString, scene::HittableList)
    geom, mat, kind = flatten(scene)
    GC.@preserve geom mat kind check(C_NULL, ccall((:rtw_scene_save, librtw), Cint,
        (Cstring, Ptr{Float32}, Ptr{Float32}, Ptr{UInt32}, UInt32), path, geom, mat, kind, length(kind)))
    path
end

"reads a `.rtwscene` file back into a HittableList of Float32 spheres"
function load_scene(path::AbstractString)
    n = Ref{UInt32}(0)
    check(C_NULL, ccall((:rtw_scene_load, librtw), Cint,
        (Cstring, Ptr{Float32}, Ptr{Float32}, Ptr{UInt32}, UInt32, Ptr{UInt32}), path, C_NULL, C_NULL, C_NULL, 0, n))
    geom, mat, kind = Matrix{Float32}(undef, 4, n[]), Matrix{Float32}(undef, 4, n[]), Vector{UInt32}(undef, n[])
    GC.@preserve geom mat kind check(C_NULL, ccall((:rtw_scene_load, librtw), Cint,
        (Cstring, Ptr{Float32}, Ptr{Float32}, Ptr{UInt32}, UInt32, Ptr{UInt32}), path, geom, mat, kind, n[], n))
    scene = HittableList()
    for k in 1:Int(n[])
        albedo = SA[mat[1, k], mat[2, k], mat[3, k]]
        m = kind[k] == RTW_LAMBERTIAN ? Lambertian(albedo) :
            kind[k] == RTW_METAL ? Metal(albedo, mat[4, k]) : Dielectric(mat[4, k])
        push!(scene, Sphere(SA[geom[1, k], geom[2, k], geom[3, k]], geom[4, k], m))
    end
    scene
end

# ---- scene_random_spheres on the device (src/scenes.jl:49-84; rtw_scene_random_spheres) ---------------------------------
"""
    scene_random_spheres_device!(ctx = default_context(); half_extent = 11, rng = TRNG[Threads.threadid()])

Builds `scene_random_spheres(; elem_type=Float32)` on the GPU -- the list the reference's loop would produce from `rng`
in its current state, bit for bit, with `rng` advanced exactly as the loop would advance it -- installs it as the scene of
`ctx` (render it with `render_resident`) and returns it as a HittableList.  `half_extent = 158` gives ~100k spheres in
milliseconds instead of the seconds of the host loop.
"""
function scene_random_spheres_device!(ctx::Context = default_context(); half_extent::Integer = 11,
                                      rng = RayTracingWeekend.TRNG[Threads.threadid()])
    cap = 4 * half_extent^2 + 4
    geom, mat, kind = Matrix{Float32}(undef, 4, cap), Matrix{Float32}(undef, 4, cap), Vector{UInt32}(undef, cap)
    state = UInt64[rng.x, rng.y]          # Xoroshiro128Plus holds its state in the fields x, y (RandomNumbers.jl 1.5.3)
    n = Ref{UInt32}(0)
    GC.@preserve geom mat kind state check(ctx.ptr, ccall((:rtw_scene_random_spheres, librtw), Cint,
        (Ptr{Cvoid}, Ptr{UInt64}, Cint, Cint, Ptr{Float32}, Ptr{Float32}, Ptr{UInt32}, UInt32, Ptr{UInt32}),
        ctx.ptr, state, half_extent, 1, geom, mat, kind, cap, n))
    rng.x, rng.y = state[1], state[2]
    scene = HittableList()
    for k in 1:Int(n[])
        albedo = SA[mat[1, k], mat[2, k], mat[3, k]]
        m = kind[k] == RTW_LAMBERTIAN ? Lambertian(albedo) :
            kind[k] == RTW_METAL ? Metal(albedo, mat[4, k]) : Dielectric(mat[4, k])
        push!(scene, Sphere(SA[geom[1, k], geom[2, k], geom[3, k]], geom[4, k], m))
    end
    scene
end

"`render` of the scene already resident in `ctx` (rtw_set_scene / scene_random_spheres_device!): rtw_render"
function render_resident(ctx::Context, cam::Camera{Float32}, image_width::Integer = 400, n_samples::Integer = 1;
                         max_depth::Integer = 16, seed::Integer = 1)
    H = Int(ccall((:rtw_image_height, librtw), Cint, (Cint,), image_width))
    img = Matrix{RGB{Float32}}(undef, H, image_width)
    st = Ref{RtwStats}()
    camref = Ref(cam)
    GC.@preserve img camref check(ctx.ptr, ccall((:rtw_render, librtw), Cint,
        (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Cint, UInt64, Ptr{Cvoid}, Ptr{RtwStats}),
        ctx.ptr, camref, image_width, n_samples, max_depth, seed, img, st))
    img
end

# ---- the drop-in --------------------------------------------------------------------------------------------------------
"""
    RayTracingWeekendB200.install!(; max_depth = 16, seed = 1, ctx = default_context())

Defines `RayTracingWeekend.render(scene::HittableList, cam::Camera{Float32 | Float64}, image_width=400, n_samples=1)`
(src/render.jl:8-9) as a call into the GPU library, replacing the package's CPU method for those two camera types:
after this, every script that calls `render(scene, cam, w, n)` runs on the B200s unchanged.  `max_depth` and `seed`
stand for the two things the reference hard-codes (`ray_color`'s depth of 16, src/ray_color.jl:14; `reseed!()` at the
top of every render, src/render.jl:21).  Opt-in because it overwrites a method of another package (Julia prints a
"method overwritten" notice); `uninstall!()` is not possible -- restart Julia to get the CPU method back.
"""
function install!(; max_depth::Integer = 16, seed::Integer = 1, ctx::Context = default_context())
    @eval RayTracingWeekend begin
        function render(scene::HittableList, cam::Union{Camera{Float32},Camera{Float64}}, image_width = 400, n_samples = 1)
            $(render_b200)(scene, cam, image_width, n_samples; max_depth = $max_depth, seed = $seed, ctx = $ctx)
        end
    end
    nothing
end

export render_b200, render_resident, install!, Context, RtwStats, set_option!, set_scene!, accumulate!, resolve, save_png,
       save_scene, load_scene, save_checkpoint, load_checkpoint!, scene_random_spheres_device!,
       RTW_OPT_MODE, RTW_OPT_BLOCKS_PER_SM, RTW_OPT_COLLECT_TIMING, RTW_OPT_RAYS_PER_LANE, RTW_OPT_SWEEP, RTW_OPT_COOP,
       RTW_OPT_TAIL, RTW_OPT_WALK, RTW_OPT_GATHER, RTW_OPT_SMALL_RENDER, RTW_MODE_FUSED, RTW_MODE_WAVEFRONT,
       RTW_MODE_CTA_WAVEFRONT, RTW_MODE_GRID, RTW_GATHER_PEER, RTW_GATHER_NCCL

end # module
