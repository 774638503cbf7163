// Kernel instantiations for arms with exactly 18 collision spheres (the 18 spheres fitted to the iiwa14 hulls).
#include "pmp_launch_impl.cuh"

PMP_DEFINE_LAUNCHERS(18)
