/* TEST INFRASTRUCTURE: the reference's own set_scaledep_GM against shim/scaledep_gm_b200.c, in one process.
 *
 * Links EVERY translation unit of the reference program (src/pinocchio.c with its main renamed) over the one-task
 * MPI / restated GSL of oracle/ref_full, runs the reference's initialization() -- which ends with the reference's
 * set_scaledep_GM (src/initialization.c:115, :1533-2026: 3 x Nsmooth x NBINS adaptive integrals on the host) --
 * keeps what it left in Smoothing.* and SPLINE_INVGROW[], then calls set_scaledep_GM_b200() (one device call through
 * include/pinb200.h: libpinb200.so on a B200, the emulated ABI elsewhere) on the same globals and prints both.
 *     sdgm_{emu,b200}.x parameter_file  ->  one JSON object on the last line of stdout
 * Nothing in the product links this file. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pinocchio.h"
#include "def_splines.h"

int set_scaledep_GM_b200(void);

static double wall(void) { return MPI_Wtime(); }

int main(int argc, char **argv) {
  int got_level;
  MPI_Init_thread(&argc, &argv, MPI_THREAD_FUNNELED, &got_level);
  MPI_Comm_rank(MPI_COMM_WORLD, &ThisTask);
  MPI_Comm_size(MPI_COMM_WORLD, &NTasks);
  if (argc < 2) {
    printf("usage: %s parameter_file\n", argv[0]);
    return 2;
  }
  cputime.total = MPI_Wtime();
  memset(&params, 0, sizeof(param_data));
  strcpy(params.ParameterFile, argv[1]);
  const double t0 = wall();
  if (initialization()) return 1;
  const double t_init = wall() - t0;
  const int S = Smoothing.Nsmooth;
  double *ref_k = (double *)malloc(3 * S * sizeof(double)), *ref_rad = (double *)malloc(S * sizeof(double));
  double *ref_x = (double *)malloc((size_t)S * NBINS * sizeof(double));
  for (int r = 0; r < S; r++) {
    ref_k[r] = Smoothing.k_GM_dens[r];
    ref_k[S + r] = Smoothing.k_GM_displ[r];
    ref_k[2 * S + r] = Smoothing.k_GM_vel[r];
    ref_rad[r] = Smoothing.Rad_GM[r];
    for (int i = 0; i < NBINS; i++) ref_x[(size_t)r * NBINS + i] = SPLINE_INVGROW[r]->x[i];
  }
  /* the reference's set_scaledep_GM alone, timed (it only depends on the cosmology set up above) */
  const double t1 = wall();
  if (set_scaledep_GM()) return 1;
  const double t_ref = wall() - t1;
  /* twice: the first call pays for the CUDA context of this process */
  const double t2 = wall();
  if (set_scaledep_GM_b200()) return 1;
  const double t_b200_first = wall() - t2;
  const double t3 = wall();
  if (set_scaledep_GM_b200()) return 1;
  const double t_b200 = wall() - t3;

  double worst_vec = 0.0, worst_rad = 0.0;
  for (int r = 0; r < S; r++) {
    worst_rad = fmax(worst_rad, fabs(Smoothing.Rad_GM[r] - ref_rad[r]));
    for (int i = 0; i < NBINS; i++) {
      /* x = log10(vector): compare the normalised sqrt-variances themselves, relative */
      const double a = pow(10., SPLINE_INVGROW[r]->x[i]), b = pow(10., ref_x[(size_t)r * NBINS + i]);
      worst_vec = fmax(worst_vec, fabs(a - b) / b);
    }
  }
  printf("\n{\"nsmooth\": %d, \"nbins\": %d, \"nkbins\": %d, \"t_initialization_s\": %.3f, \"t_reference_s\": %.4f, \"t_b200_binding_first_call_s\": %.4f, \"t_b200_binding_s\": %.4f, "
         "\"invgrow_vector_max_rel\": %.3e, \"rad_gm_max_abs\": %.3e, \"k_gm\": [",
         S, NBINS, NkBINS, t_init, t_ref, t_b200_first, t_b200, worst_vec, worst_rad);
  const double *mine[3] = {Smoothing.k_GM_dens, Smoothing.k_GM_displ, Smoothing.k_GM_vel};
  for (int q = 0; q < 3; q++)
    for (int r = 0; r < S; r++) printf("%s[%d, %d, %.17g, %.17g]", (q || r) ? ", " : "", q, r, ref_k[q * S + r], mine[q][r]);
  printf("]}\n");
  MPI_Finalize();
  return 0;
}
