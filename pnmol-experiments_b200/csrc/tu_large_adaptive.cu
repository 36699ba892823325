// Translation unit of the LARGE_ADAPTIVE kernel group: the device code lives in the .cuh files; every kernel is DEFINED in
// exactly one translation unit (PNMOL_TU_LARGE_ADAPTIVE here) and only declared in the others, so that the groups compile in
// parallel.
#define PNMOL_TU_LARGE_ADAPTIVE
#include "ek1_kernels.cuh"
#include "ek1_large.cuh"
#include "ek1_small.cuh"
