// TEST INFRASTRUCTURE ONLY: C-ABI re-export of reference src/repeat_cell.cpp.
#include "wrap_common.h"
#include "repeat_cell.cpp"
extern "C" {
// repeat_cell.cpp:19 repeat_cell
void ref_repeat_cell(double *new_pos, const double *old_box, const double *old_pos, int n_old, int nx, int ny, int nz,
                     int num_t)
{
    repeat_cell(W1D(new_pos, (size_t)n_old * nx * ny * nz * 3), A2D(old_box, 3, 3), A2D(old_pos, n_old, 3), nx, ny, nz,
                num_t);
}
}
