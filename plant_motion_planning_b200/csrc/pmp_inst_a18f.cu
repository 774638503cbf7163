// First-violation-only edge-kernel instantiations for sphere-count padding 18 (see pmp_launch.h).
#include "pmp_launch_impl.cuh"

PMP_DEFINE_FIRST_LAUNCHERS(18)
