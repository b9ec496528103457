#include "dsc_internal.cuh"
namespace dsc {
int gemm_f64_dmma(Ctx*, int, int, int, int, int, const double*, const double*, const double*, double*) { return 1; }
}
