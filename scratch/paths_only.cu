// compile-only probe: ptxas -v of the ray-path stage without the rest of the library
#include "../atm_raytracer_b200/csrc/kernels.cuh"
namespace atmrt {
template __global__ void k_ray_chain<false>(const __grid_constant__ DevScene, DevBuffers, PathRecords, int);
template __global__ void k_ray_elements<false>(const __grid_constant__ DevScene, DevBuffers, PathRecords);
}
