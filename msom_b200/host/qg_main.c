/*
 * qg_main.c -- `qg.e [params file]`, the process surface of msqg/qg.c:34-48:
 *   read_params(argv[1] | "params.in"); create_outdir(); init_grid(N); size(L0); run();
 * Extra, optional switches (the reference selects these at compile time):
 *   --modal       MODE_PV_INVERT 1 (msqg/qg.h:4)
 *   --stochastic  -D_STOCHASTIC=1  (msqg/qg.c:25)
 *   --device D    CUDA device
 */
#include "../../include/msqg.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char *argv[]) {
  const char *params = "params.in";
  for (int a = 1; a < argc; a++) {
    if (!strcmp(argv[a], "--modal")) qg_set_mode_pv_invert(1);
    else if (!strcmp(argv[a], "--stochastic")) qg_set_stochastic(1);
    else if (!strcmp(argv[a], "--device") && a + 1 < argc) qg_set_device(atoi(argv[++a]));
    else params = argv[a];
  }
  if (read_params((char *)params)) return 0; /* the reference exit(0)s on a missing file */
  if (create_outdir()) return 1;
  init_grid(qg_params()->N);
  int rc = run();
  return rc ? 1 : 0;
}
