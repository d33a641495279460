// modulated.cu -- placeholder, replaced below in this round
#include "krylov.cuh"
extern "C" int fdfd_solve_modulated(fdfd_ctx* ctx, const fdfd_grid_t* g, double omega, double Omega, int nsidebands,
                                    int sharedpml, const fdfd_c128* eps_r, const fdfd_c128* deps_r, const fdfd_c128* src,
                                    const fdfd_solve_opts_t* opts, fdfd_c128* fields, fdfd_info_t* info) {
  fdfd_set_error(ctx, "fdfd_solve_modulated: not built yet"); return FDFD_ERR_ARG;
}
