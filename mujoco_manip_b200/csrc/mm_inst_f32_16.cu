// one (precision, group) instantiation of the step / reset kernels
#include "mm_launch.cuh"
MM_DEFINE_INST(f32_16, float, 16)
