"""versalignlib_b200 -- a B200 (sm_100a) CUDA kernel plug-in for versalignLib's batched
Smith-Waterman / "Needleman-Wunsch" DP hot path, behind the reference's own
AlignmentKernel plug-in boundary.  See DESIGN.md and INTEGRATION.md.

  build      in-tree nvcc / g++ builds of libCUDAKernel.so and libva_host.so
  host       driver-side loader: PluginHost (dlopen + the two virtual calls)
  capi       ctypes binding of the C ABI in include/versalign_cuda.h
  synth      seeded synthetic batches in the reference's input convention
"""
__version__ = "0.1.0"
