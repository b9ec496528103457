#include "dsc_internal.cuh"
namespace dsc {
int gemm_tf32x3_tcgen05(Ctx*, int, int, int, int, int, const float*, const float*, const float*, float*) { return 1; }
}
