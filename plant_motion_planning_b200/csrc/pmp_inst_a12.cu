// Kernel instantiations for arms with up to 12 collision spheres.
#include "pmp_launch_impl.cuh"

PMP_DEFINE_LAUNCHERS(12)
