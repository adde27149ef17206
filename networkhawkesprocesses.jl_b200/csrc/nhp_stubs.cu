// nhp_stubs.cu -- every entry point of include/nhp.h is implemented; kept as a placeholder translation unit.
#include "nhp_internal.cuh"
