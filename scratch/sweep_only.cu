// compile-only probe: ptxas -v of the fused stage-C kernel without the rest of the library
#include "../atm_raytracer_b200/csrc/kernels.cuh"
namespace atmrt {
// (not a template any more)
//template __global__ void k_sweep_bits(const __grid_constant__ DevScene, DevBuffers, SweepLists, int, int);
}
