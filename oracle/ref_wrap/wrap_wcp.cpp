// TEST INFRASTRUCTURE ONLY: C-ABI re-export of reference src/warren_cowley_parameter.cpp.
#include "wrap_common.h"
#include "warren_cowley_parameter.cpp"
extern "C" {
// warren_cowley_parameter.cpp:9 get_wcp
void ref_wcp(const int *verlet, int N, int M, const int *nn, const int *type_list, int T, double *WCP, int num_t)
{
    get_wcp(A2I(verlet, N, M), A1I(nn, N), A1I(type_list, N), T, W2D(WCP, T, T), num_t);
}
}
