// TEST INFRASTRUCTURE ONLY: C-ABI re-export of reference src/identify_fcc_planar_faults.cpp.
#include "wrap_common.h"
#include "identify_fcc_planar_faults.cpp"
extern "C" {
// identify_fcc_planar_faults.cpp:43 identify_sftb_fcc
void ref_identify_sftb_fcc(const int *hcp_indices, int n_hcp, int *hcp_neighbors, const int *ptm_indices,
                           const int *structure_types, int N, int *fault_types, int identify_esf, int num_t)
{
    identify_sftb_fcc(A1I(hcp_indices, n_hcp), W2I(hcp_neighbors, n_hcp, 12), A2I(ptm_indices, N, 12),
                      A1I(structure_types, N), W1I(fault_types, N), identify_esf != 0, num_t);
}
}
