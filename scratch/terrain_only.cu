// compile-only probe: ptxas -v of stage A without the rest of the library
#include "../atm_raytracer_b200/csrc/kernels.cuh"
namespace atmrt {
template __global__ void k_terrain_profile<0>(const __grid_constant__ DevScene, DevTerrain, DevBuffers, int);
}
