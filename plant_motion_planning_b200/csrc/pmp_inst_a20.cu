// Kernel instantiations for arms with up to 20 collision spheres (the 18 spheres fitted to the iiwa14 hulls).
#include "pmp_launch_impl.cuh"

PMP_DEFINE_LAUNCHERS(20)
