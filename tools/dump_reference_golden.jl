# dump_reference_golden.jl -- run where Julia and RayTracingWeekend.jl are installed (they are not in the build image).
#
# Renders the reference's own smoke-test image with the REFERENCE's CPU code (test/runtests.jl:190-194:
# render(scene_2_spheres(; elem_type=T), default_camera(SA{T}[0,0,0]), 96, 16)), single-threaded so that the image is
# the one thread 1's Xoroshiro128Plus stream produces (src/rand.jl:2-13), times it, and writes
#     <out>                 : "RTWGOLD1", T code (0 = Float32, 1 = Float64), H, W (Int32), then H*W*3 values of type T in
#                             the memory order of Matrix{RGB{T}} (column-major) -- what rtw_render's out_rgb holds
#     <out>.stream.bin      : the first 4096 draws of Xoroshiro128Plus(1) after reseed!() as Float32 and as Float64
# so that a box with Julia can pin the oracle's xoroshiro mode (oracle/rtw_oracle.c, rtwo_xoroshiro_*), which restates
# RandomNumbers.jl from the published algorithm and is otherwise unverified (SURVEY.md section 8c, Appendix B).
#
#   julia --threads=1 tools/dump_reference_golden.jl out.bin [Float32|Float64]
using RayTracingWeekend
using StaticArrays
using Images: RGB

out = length(ARGS) >= 1 ? ARGS[1] : "reference_golden_cfg1.bin"
T = length(ARGS) >= 2 && ARGS[2] == "Float64" ? Float64 : Float32
Threads.nthreads() == 1 || @warn "run with --threads=1: the reference's image depends on the thread count"

reseed!()
open(out * ".stream.bin", "w") do io
    write(io, Float32[trand(Float32) for _ in 1:4096])
    reseed!()
    write(io, Float64[trand(Float64) for _ in 1:4096])
end

cam = default_camera(SA{T}[0, 0, 0])
scene = scene_2_spheres(; elem_type = T)
render(scene, cam, 96, 16)                       # compile
t = @elapsed img = render(scene, cam, 96, 16)    # render() reseeds by itself (src/render.jl:21)
println("reference render(scene_2_spheres, 96, 16), $T, 1 thread: $(round(t * 1e6)) us")
open(out, "w") do io
    write(io, b"RTWGOLD1")
    write(io, Int32(T == Float64 ? 1 : 0), Int32(size(img, 1)), Int32(size(img, 2)))
    write(io, reinterpret(T, vec(img)))
end
println("wrote $out and $out.stream.bin")

# the headline scene at a size the CPU finishes: timing only
reseed!()
big = scene_random_spheres(; elem_type = T)
cam1 = default_camera(SA{T}[13, 2, 3], SA{T}[0, 0, 0], SA{T}[0, 1, 0], T(20), T(16 / 9), T(0.1), T(10))
render(big, cam1, 96, 1)
t = @elapsed render(big, cam1, 400, 4)
println("reference render(scene_random_spheres, 400, 4), $T, $(Threads.nthreads()) thread(s): $(round(t * 1e3, digits = 1)) ms")
