// Specialised streaming kernels (filled in after the generic path is parity-green).
#include "kernels.cuh"
namespace b200q {}
