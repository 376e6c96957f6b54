// TEST INFRASTRUCTURE ONLY: C-ABI re-export of reference src/polyhedral_template_matching.cpp
// (links the vendored extern/ptm/ptm_*.cpp compiled from /root/reference).
#include "wrap_common.h"
#include "polyhedral_template_matching.cpp"
extern "C" {
// polyhedral_template_matching.cpp:135 get_ptm
void ref_ptm(const char *structure, const double *x, const double *y, const double *z, int N, BOXARGS,
             const int *verlet, int M, const int *atom_types, int ntypes, double rmsd_threshold, double *output,
             int ocols, int *ptm_indices, int icols, int num_t)
{
    get_ptm(structure, A1D(x, N), A1D(y, N), A1D(z, N), BOXPASS, A2I(verlet, N, M), A1I(atom_types, ntypes),
            rmsd_threshold, W2D(output, N, ocols), W2I(ptm_indices, N, icols), num_t);
}
}
