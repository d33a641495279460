// eigen.cu -- placeholder, replaced below in this round
#include "krylov.cuh"
extern "C" int fdfd_eigenfrequency(fdfd_ctx* ctx, const fdfd_grid_t* g, int pol, double omega0, int nev, int which, int ncv,
                                   const fdfd_c128* eps_r, const fdfd_solve_opts_t* opts, fdfd_c128* omega_out,
                                   fdfd_c128* fields, fdfd_info_t* info) {
  fdfd_set_error(ctx, "fdfd_eigenfrequency: not built yet"); return FDFD_ERR_ARG;
}
