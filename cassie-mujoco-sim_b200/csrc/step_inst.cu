// step_inst.cu -- one instance of the fused step kernel per compilation: nvcc ... -DINST_REAL=float -DINST_DR=0 (0 plain / 1 extended / 2 plain + estimator) -DINST_FEAT=0 -DINST_TAG=f00
// (build.py compiles the instances in parallel).  The library reaches an instance through its entry function below.
#include "step_kernel.cuh"

#define INST_CAT2(a, b) a##b
#define INST_CAT(a, b) INST_CAT2(a, b)

template __global__ void cassie::cassie_step_kernel<INST_REAL, INST_DR, INST_FEAT>(const cassie::DevModel<INST_REAL> *__restrict__, cassie::EnvArrays<INST_REAL>, int, int);

extern "C" const void *INST_CAT(cassie_step_entry_, INST_TAG)(void) { return (const void *)&cassie::cassie_step_kernel<INST_REAL, INST_DR, INST_FEAT>; }
