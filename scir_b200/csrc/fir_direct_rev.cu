// fir_direct_rev.cu -- the anticausal (DIR = -1) instantiations of the FP32 direct FIR kernels, compiled as their
// own translation unit so that the two halves of fir_direct.cu build in parallel (see the note there).
#define SCIR_FIR_DIRECT_REV 1
#include "fir_direct.cu"
