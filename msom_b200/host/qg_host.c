/*
 * qg_host.c -- plain-C host side of the msqg driver, re-hosted on the CUDA C ABI.
 *
 * Mirrors the reference's global-state surface (same names, argument meaning
 * and messages) so that msqg/qg.c's main(), the SWIG module `qg`
 * (msqg/qg.i, qg_bfn.i) and the post-processing scripts keep working:
 *
 *   read_params      msqg/qg.h:689-761        create_outdir   msqg/qg.h:766-780
 *   backup_config    msqg/qg.h:782-835        set_vars        msqg/qg.h:837-925
 *   set_const        msqg/qg.h:931-1116 (CWD input files :940-984)
 *   trash_vars       msqg/qg.h:1130-1154      set_vars_bfn/trash_vars_bfn  msqg/qg_bfn.h:7-15
 *   pystep_bfn / pyq2p / pyp2q                msqg/qg_bfn.h:21-103
 *   run              [BASILISK] predictor-corrector.h run() + the events of
 *                    msqg/qg.c:53-173 (init, write_const, writestdout, output)
 *   .bas files       msqg/auxiliar_input.h:24-59 (input_matrixl), :101-149 (output_matrixl)
 *
 * All field arithmetic happens on the GPU through include/msqg.h layer (1);
 * this file only parses, schedules events, and moves float32 files.
 * Where the reference calls exit(0) this code prints the same message and
 * returns a negative code.
 */
#include "../../include/msqg.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/types.h>

#define TEPS 1e-9
#define HUGEV 1e30

static msqg_params P;
static int P_init = 0;
static msqg_model *M = NULL;
static int g_device = 0;
static char dpath[80] = "";
static double g_t = 0., g_dt = 0.;
static int g_i = 0;
static int g_verbose = 1;

static void ensure_defaults(void) {
  if (!P_init) { msqg_default_params(&P); P_init = 1; }
}
msqg_model *qg_model(void) { return M; }
msqg_params *qg_params(void) { ensure_defaults(); return &P; }
int qg_set_device(int device) { g_device = device; return MSQG_OK; }
int qg_set_mode_pv_invert(int mode) { ensure_defaults(); P.mode_pv_invert = mode; return MSQG_OK; }
int qg_set_stochastic(int on) { ensure_defaults(); P.stochastic = on; return MSQG_OK; }
int qg_set_verbose(int v) { g_verbose = v; return MSQG_OK; }
double qg_time(void) { return g_t; }
int qg_iter(void) { return g_i; }

static int fail(int rc) {
  fprintf(stdout, "%s\n", msqg_last_error());
  return rc;
}

/* ---------------------------------------------------------------- .bas files */
int qg_write_bas(const char *name, int nf, int N, double L0, const double *v) {
  FILE *fp = fopen(name, "w");
  if (!fp) return MSQG_ERR_FILE;
  float fn = N, Delta = L0 / fn;
  float *row = (float *)malloc(sizeof(float) * (size_t)(N + 1));
  for (int k = 0; k < nf; k++) {
    row[0] = fn;
    for (int j = 0; j < N; j++) row[j + 1] = Delta * j + 0. + Delta / 2.;
    fwrite(row, sizeof(float), (size_t)N + 1, fp);
    for (int i = 0; i < N; i++) {
      row[0] = Delta * i + 0. + Delta / 2.;
      for (int j = 0; j < N; j++) row[j + 1] = v[(size_t)N * N * k + (size_t)N * j + i];
      fwrite(row, sizeof(float), (size_t)N + 1, fp);
    }
  }
  free(row);
  fclose(fp);
  return MSQG_OK;
}

int qg_read_bas(const char *name, int nf, int N, double L0, double *v) {
  FILE *fp = fopen(name, "r");
  if (!fp) return MSQG_ERR_FILE;
  const double Delta = L0 / N;
  for (int k = 0; k < nf; k++) {
    float width = 0;
    if (fread(&width, sizeof(float), 1, fp) != 1) { fclose(fp); return MSQG_ERR_FILE; }
    const int pn = (int)width;
    if (pn <= 0) { fclose(fp); return MSQG_ERR_FILE; }
    float *buf = (float *)malloc(sizeof(float) * (size_t)pn * (pn + 1));
    float *yp = (float *)malloc(sizeof(float) * (size_t)pn);
    int ok = fread(yp, sizeof(float), pn, fp) == (size_t)pn;
    for (int i = 0; ok && i < pn; i++) ok = fread(buf + (size_t)i * (pn + 1), sizeof(float), pn + 1, fp) == (size_t)pn + 1;
    if (!ok) { free(buf); free(yp); fclose(fp); return MSQG_ERR_FILE; }
    for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++) {
        const double x = (i + 0.5) * Delta, y = (j + 0.5) * Delta;
        const int ii = (x - 0.) * width / L0, jj = (y - 0.) * width / L0;
        v[(size_t)N * N * k + (size_t)N * j + i] =
            (ii >= 0 && ii < width && jj >= 0 && jj < width) ? buf[(size_t)ii * (pn + 1) + 1 + jj] : 0.;
      }
    free(buf); free(yp);
  }
  fclose(fp);
  return MSQG_OK;
}

/* ---------------------------------------------------------------- setup */
int read_params(char *path2file) {
  ensure_defaults();
  int rc = msqg_read_params(path2file, &P);
  if (rc == MSQG_ERR_FILE) { fprintf(stdout, "file %s not found\n", path2file); return rc; }
  if (rc) return fail(rc);
  fprintf(stdout, "Config: N = %d, nl = %d, L0 = %g\n", P.N, P.nl, P.L0);
  return MSQG_OK;
}

int init_grid(int n) { ensure_defaults(); P.N = n; return MSQG_OK; }

int set_vars(void) {
  ensure_defaults();
  if (M) { msqg_destroy(M); M = NULL; }
  fprintf(stdout, "Create main variables .. ");
  int rc = msqg_create(&P, g_device, &M);
  if (rc) { fprintf(stdout, "\n"); return fail(rc); }
  if (P.stochastic) {
    fprintf(stdout, "Create stochastic variables .. ");
    fprintf(stdout, ".. ok\n");
    fprintf(stdout, "Read stochastic sigma files (std deviation of the noise):\n");
    char name[80];
    snprintf(name, sizeof(name), "s_stoch_%dl_N%d.bas", P.nl, P.N);
    size_t sz = (size_t)P.nl * P.N * P.N;
    double *buf = (double *)malloc(sizeof(double) * sz);
    if (qg_read_bas(name, P.nl, P.N, P.L0, buf) == MSQG_OK) {
      rc = msqg_set_field(M, MSQG_SSTOCH, buf);
      fprintf(stdout, "%s .. ok\n", name);
    }
    free(buf);
    if (rc) return fail(rc);
  }
  g_t = 0.; g_i = 0; g_dt = 0.;
  fprintf(stdout, "ok\n");
  return MSQG_OK;
}

static int read_list_file(const char *name, int id, int nf) {
  size_t sz = (size_t)nf * P.N * P.N;
  double *buf = (double *)malloc(sizeof(double) * sz);
  int rc = qg_read_bas(name, nf, P.N, P.L0, buf);
  if (rc == MSQG_OK) {
    rc = msqg_set_field(M, id, buf);
    if (rc == MSQG_OK) fprintf(stdout, "%s .. ok\n", name);
    free(buf);
    return rc == MSQG_OK ? 1 : rc;
  }
  free(buf);
  return 0; /* absent: silently skipped like the reference */
}

int set_const(void) {
  if (!M) return MSQG_ERR_ARG;
  fprintf(stdout, "Read input files:\n");
  char name[80];
  FILE *fp;
  int rc;
  snprintf(name, sizeof(name), "dh_%dl.bin", P.nl);
  if ((fp = fopen(name, "r"))) {
    float dh[MSQG_MAXL];
    size_t got = fread(dh, sizeof(float), P.nl, fp);
    fclose(fp);
    if (got == (size_t)P.nl) {
      double dhd[MSQG_MAXL];
      for (int l = 0; l < P.nl; l++) dhd[l] = dh[l];
      if ((rc = msqg_set_dh(M, dhd))) return fail(rc);
      fprintf(stdout, "%s .. ok\n", name);
    }
  }
  snprintf(name, sizeof(name), "psipg_%dl_N%d.bas", P.nl, P.N);
  if ((rc = read_list_file(name, MSQG_PSIPG, P.nl)) < 0) return fail(rc);
  snprintf(name, sizeof(name), "frpg_%dl_N%d.bas", P.nl, P.N);
  if ((rc = read_list_file(name, MSQG_FR, P.nl)) < 0) return fail(rc);
  snprintf(name, sizeof(name), "rdpg_%dl_N%d.bas", P.nl, P.N);
  if ((rc = read_list_file(name, MSQG_RD, 1)) < 0) return fail(rc);
  snprintf(name, sizeof(name), "topo.bas");
  if ((rc = read_list_file(name, MSQG_TOPO, 1)) < 0) return fail(rc);
  if (rc == 1) msqg_set_flag_topo(M, 1);
  snprintf(name, sizeof(name), "qforc_%dl_N%d.bas", P.nl, P.N);
  if ((rc = read_list_file(name, MSQG_QFORC, P.nl)) < 0) return fail(rc);
  rc = msqg_set_const(M);
  if (rc) return fail(rc);
  return MSQG_OK;
}

int create_outdir(void) {
  for (int i = 1; i < 10000; i++) {
    snprintf(dpath, sizeof(dpath), "outdir_%04d/", i);
    if (mkdir(dpath, 0777) == 0) {
      fprintf(stdout, "Writing output in %s\n", dpath);
      return MSQG_OK;
    }
  }
  return MSQG_ERR_FILE;
}
const char *qg_outdir(void) { return dpath; }
int qg_set_outdir(const char *d) { snprintf(dpath, sizeof(dpath), "%s", d); return MSQG_OK; }

static int write_list(const char *name, int id) {
  int nf = msqg_nfields(M, id);
  if (nf <= 0) return MSQG_ERR_ARG;
  size_t sz = (size_t)nf * P.N * P.N;
  double *buf = (double *)malloc(sizeof(double) * sz);
  int rc = msqg_get_field(M, id, buf);
  if (rc == MSQG_OK) rc = qg_write_bas(name, nf, P.N, P.L0, buf);
  free(buf);
  return rc;
}

/* write_field(list, name, rescale), auxiliar_input.h:153-167: psi[] *= rescale, then the float32 matrix */
static int write_list_scaled(const char *name, int id, double rescale) {
  const int nf = msqg_nfields(M, id);
  if (nf <= 0) return MSQG_ERR_ARG;
  size_t sz = (size_t)nf * P.N * P.N;
  double *buf = (double *)malloc(sizeof(double) * sz);
  int rc = msqg_get_field(M, id, buf);
  if (!rc) {
    if (rescale != 0) for (size_t c = 0; c < sz; c++) buf[c] *= rescale;
    rc = qg_write_bas(name, nf, P.N, P.L0, buf);
  }
  free(buf);
  return rc;
}

int backup_config(void) {
  if (!M) return MSQG_ERR_ARG;
  fprintf(stdout, "Backup config\n");
  char name[200];
  int ch;
  snprintf(name, sizeof(name), "%sparams.in", dpath);
  FILE *source = fopen("params.in", "r");
  FILE *target = fopen(name, "w");
  if (source && target)
    while ((ch = fgetc(source)) != EOF) fputc(ch, target);
  if (source) fclose(source);
  if (target) fclose(target);
  snprintf(name, sizeof(name), "%ssig_filt.bas", dpath);
  write_list(name, MSQG_SIGFILT);
  if (P.mode_pv_invert) {
    snprintf(name, sizeof(name), "%siBu.bas", dpath);
    write_list(name, MSQG_IBU);
  } else {
    snprintf(name, sizeof(name), "%srdpg_%dl_N%d.bas", dpath, P.nl, P.N);
    write_list(name, MSQG_RD);
  }
  snprintf(name, sizeof(name), "%spsipg_%dl_N%d.bas", dpath, P.nl, P.N);
  write_list(name, MSQG_PSIPG);
  snprintf(name, sizeof(name), "%sfrpg_%dl_N%d.bas", dpath, P.nl, P.N);
  write_list(name, MSQG_FR);
  snprintf(name, sizeof(name), "%sqforc_%dl_N%d.bas", dpath, P.nl, P.N);
  write_list(name, MSQG_QFORC);
  float dh[MSQG_MAXL];
  double dhd[MSQG_MAXL];
  msqg_get_dh(M, dhd);
  for (int l = 0; l < P.nl; l++) dh[l] = dhd[l];
  snprintf(name, sizeof(name), "%sdh_%dl.bin", dpath, P.nl);
  FILE *fp = fopen(name, "w");
  if (fp) { fwrite(dh, sizeof(float), P.nl, fp); fclose(fp); }
  return MSQG_OK;
}

int trash_vars(void) {
  if (M) { msqg_destroy(M); M = NULL; }
  return MSQG_OK;
}
/* bfn_tendl / bfn_forcl are the handle's DQ list here; nothing extra to allocate */
int set_vars_bfn(void) { return M ? MSQG_OK : MSQG_ERR_ARG; }
int trash_vars_bfn(void) { return MSQG_OK; }

/* ---------------------------------------------------------------- python entry points */
static int shape_ok(int a, int b, int c) { return a == P.nl && b == P.N && c == P.N; }

int pystep_bfn(double *varin_py, int len1, int len2, int len3, double *tend_py, int len4, int len5, int len6,
               double direction, int vartype) {
  if (!M || !shape_ok(len1, len2, len3) || !shape_ok(len4, len5, len6)) return MSQG_ERR_ARG;
  int rc;
  if (vartype == 0) {
    /* the reference still flips the dissipation signs and resets bfn_tendl */
    fprintf(stdout, "temporary disabled psi tendency\n");
    if ((rc = msqg_bfn_direction(M, direction))) return fail(rc);
    return MSQG_OK;
  } else if (vartype == 1) {
    if ((rc = msqg_set_field(M, MSQG_Q, varin_py))) return fail(rc);
    if ((rc = msqg_tendency_bfn(M, direction))) return fail(rc);
    if ((rc = msqg_get_field(M, MSQG_DQ, tend_py))) return fail(rc);
  }
  return MSQG_OK;
}

int pyq2p(double *po_py, int len7, int len8, int len9, double *qo_py, int len10, int len11, int len12) {
  if (!M || !shape_ok(len7, len8, len9) || !shape_ok(len10, len11, len12)) return MSQG_ERR_ARG;
  int rc;
  if ((rc = msqg_reset_field(M, MSQG_PSI))) return fail(rc); /* reset_layer_var: interior only */
  if ((rc = msqg_set_field(M, MSQG_Q, qo_py))) return fail(rc);
  if ((rc = msqg_invertq(M, MSQG_Q))) return fail(rc);
  if ((rc = msqg_get_field(M, MSQG_PSI, po_py))) return fail(rc);
  return MSQG_OK;
}

int pyp2q(double *po_py, int len13, int len14, int len15, double *qo_py, int len16, int len17, int len18) {
  if (!M || !shape_ok(len13, len14, len15) || !shape_ok(len16, len17, len18)) return MSQG_ERR_ARG;
  int rc;
  if ((rc = msqg_set_field(M, MSQG_PSI, po_py))) return fail(rc);
  if ((rc = msqg_comp_q(M))) return fail(rc);
  if ((rc = msqg_get_field(M, MSQG_Q, qo_py))) return fail(rc);
  return MSQG_OK;
}

/* ---------------------------------------------------------------- energy diagnostics, qg_energy.h / qg_energy.i */
int set_vars_energy(void) { return M ? fail(msqg_reset_energy(M)) : MSQG_ERR_ARG; }
int trash_vars_energy(void) { return MSQG_OK; } /* the lists live and die with the handle (trash_vars) */
/* pystep_de, qg_energy.h:294-340: ediag = 1, dt = 1 (locals).  filter_de is called with pol in the po_mft slot
 * (:330): psi is zero afterwards.  onlyKE zeroes the stretching field, as the reference does (permanently). */
int pystep_de(double *po_py, int len1, int len2, int len3, double *de_bf_py, int len4, int len5, int len6,
              double *de_vd_py, int len7, int len8, int len9, double *de_j1_py, int len10, int len11, int len12,
              double *de_j2_py, int len13, int len14, int len15, double *de_j3_py, int len16, int len17, int len18,
              double *de_ft_py, int len19, int len20, int len21, int onlyKE) {
  if (!M) return MSQG_ERR_ARG;
  if (!shape_ok(len1, len2, len3) || !shape_ok(len4, len5, len6) || !shape_ok(len7, len8, len9) ||
      !shape_ok(len10, len11, len12) || !shape_ok(len13, len14, len15) || !shape_ok(len16, len17, len18) ||
      !shape_ok(len19, len20, len21)) return MSQG_ERR_ARG;
  int rc;
  if ((rc = msqg_set_field(M, MSQG_PSI, po_py))) return fail(rc);
  if ((rc = msqg_reset_energy(M))) return fail(rc);
  if ((rc = msqg_comp_q(M))) return fail(rc);
  if (onlyKE == 1 && (rc = msqg_reset_field(M, MSQG_STR))) return fail(rc);
  if ((rc = msqg_energy_tend(M, 1., 1.))) return fail(rc);
  if ((rc = msqg_filter_de_pm(M, P.dtflt, 1., MSQG_PSI))) return fail(rc);
  double *out[6] = {de_bf_py, de_vd_py, de_j1_py, de_j2_py, de_j3_py, de_ft_py};
  const int ids[6] = {MSQG_DE_BF, MSQG_DE_VD, MSQG_DE_J1, MSQG_DE_J2, MSQG_DE_J3, MSQG_DE_FT};
  for (int k = 0; k < 6; k++) if ((rc = msqg_get_field(M, ids[k], out[k]))) return fail(rc);
  return MSQG_OK;
}

/* ---------------------------------------------------------------- run() */
/* init event, qg.c:53-72: psi from p0.bas or 1e-3*noise(), mean removed.
 * noise() = 1 - 2*rand()/RAND_MAX [BASILISK]; traversal x outer, y inner, layers innermost. */
int qg_init_event(void) {
  if (!M) return MSQG_ERR_ARG;
  const int N = P.N, nl = P.nl;
  size_t sz = (size_t)nl * N * N;
  double *psi = (double *)malloc(sizeof(double) * sz);
  if (qg_read_bas("p0.bas", nl, N, P.L0, psi) != MSQG_OK) {
    for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++)
        for (int l = 0; l < nl; l++) psi[(size_t)N * N * l + (size_t)N * j + i] = 1e-3 * (1. - 2. * rand() / (double)RAND_MAX);
  }
  const double Delta = P.L0 / N;
  for (int l = 0; l < nl; l++) {
    double *po = psi + (size_t)N * N * l;
    double sum = 0., volume = 0.;
    for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++) { volume += Delta * Delta; sum += Delta * Delta * po[(size_t)N * j + i]; }
    for (size_t c = 0; c < (size_t)N * N; c++) po[c] -= sum / volume;
  }
  int rc = msqg_set_field(M, MSQG_PSI, psi);
  free(psi);
  if (rc) return fail(rc);
  if (P.nptr > 0) { /* qg.c:75-90: tracers from ptr0.bas or 1e-3*noise() (the same rand() stream), relaxation field */
    const int nf = nl * P.nptr;
    size_t szp = (size_t)nf * N * N;
    double *tr = (double *)calloc(szp, sizeof(double));
    if (qg_read_bas("ptr0.bas", nf, N, P.L0, tr) != MSQG_OK) {
      for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++)
          for (int k = 0; k < nf; k++) tr[(size_t)N * N * k + (size_t)N * j + i] = 1e-3 * (1. - 2. * rand() / (double)RAND_MAX);
    }
    rc = msqg_set_field(M, MSQG_PTR, tr);
    if (!rc && qg_read_bas("ptr_relax.bas", nf, N, P.L0, tr) == MSQG_OK) rc = msqg_set_field(M, MSQG_PTR_RELAX, tr);
    free(tr);
    if (rc) return fail(rc);
  }
  return MSQG_OK;
}

/* One pass of the event loop + one predictor-corrector step; returns 1 while the
 * run continues, 0 when the last bounded event is exhausted, <0 on error. */
static double ev_out_t = 0.;
static int ev_out_alive = 1;
/* filter (t = dtflt; t <= tend+1e-10; t += dtflt), qg.h:655-658 and qg_energy.h:270-273 */
static double ev_flt_t = 0.;
static int ev_flt_alive = 0;

int qg_run_reset(void) {
  ev_out_t = 0.; ev_out_alive = 1; g_t = 0.; g_i = 0; g_dt = 0.;
  ev_flt_t = P.dtflt; ev_flt_alive = (P.dtflt > 0) && (ev_flt_t <= P.tend + 1e-10);
  return MSQG_OK;
}

int qg_run_iteration(int write_files) {
  int rc;
  if (ev_flt_alive && fabs(g_t - ev_flt_t) <= TEPS * g_t) {
    fprintf(stdout, "Filter solution\n");
    if ((rc = msqg_wavelet_filter(M, P.dtflt))) return fail(rc);
    if (P.ediag > -1 && (rc = msqg_filter_de(M, P.dtflt, (double)P.ediag))) return fail(rc);
    ev_flt_t += P.dtflt;
    if (!(ev_flt_t <= P.tend + 1e-10)) ev_flt_alive = 0;
  }
  /* comp_diag (i++), qg_energy.h:289-291: defined before qg.c's events, so it runs first.  `dt` is [BASILISK]'s
     global time step, 1. before the first step (common.h) */
  if (P.ediag > -1) {
    if ((rc = msqg_energy_tend(M, g_i == 0 ? 1. : g_dt, (double)P.ediag))) return fail(rc);
  }
  /* writestdout (i++), qg.c:101-109 */
  if (g_verbose) {
    double ke = 0.;
    if ((rc = msqg_ke1(M, &ke))) return fail(rc);
    fprintf(stdout, "i = %i, dt = %g, t = %g, ke_1 = %g\n", g_i, g_dt, g_t, ke);
  }
  /* output (t = 0; t <= tend+1e-10; t += dtout), qg.c:112-173 */
  double tnext = HUGEV;
  if (ev_out_alive && fabs(g_t - ev_out_t) <= TEPS * g_t) {
    fprintf(stdout, "write file\n");
    if ((rc = msqg_invertq(M, MSQG_Q))) return fail(rc);
    if (write_files) {
      char name[200];
      snprintf(name, sizeof(name), "%spo%09d.bas", dpath, g_i);
      write_list(name, MSQG_PSI);
      snprintf(name, sizeof(name), "%sqo%09d.bas", dpath, g_i);
      write_list(name, MSQG_Q);
    }
    if (P.dtflt > 0) { /* qg.c:124-129: invertq(tmpl, qofl), pf file, nbar = 0 */
      if ((rc = msqg_invert_filter_mean(M))) return fail(rc);
      if (write_files) {
        char name[200];
        snprintf(name, sizeof(name), "%spf%09d.bas", dpath, g_i);
        write_list(name, MSQG_TMP);
      }
    }
    if (P.nptr > 0 && write_files) { /* qg.c:168-171 */
      char name[200];
      snprintf(name, sizeof(name), "%sptr%09d.bas", dpath, g_i);
      write_list(name, MSQG_PTR);
    }
    if (P.ediag > -1) { /* qg.c:139-166: write_field(de_*, name, 1/dtout), then reset_layer_var */
      static const char *nm[6] = {"de_bf", "de_vd", "de_j1", "de_j2", "de_j3", "de_ft"};
      static const int ids[6] = {MSQG_DE_BF, MSQG_DE_VD, MSQG_DE_J1, MSQG_DE_J2, MSQG_DE_J3, MSQG_DE_FT};
      const double idtout = 1 / P.dtout;
      if (write_files) {
        char name[200];
        for (int k = 0; k < 6; k++) {
          snprintf(name, sizeof(name), "%s%s%09d.bas", dpath, nm[k], g_i);
          write_list_scaled(name, ids[k], idtout);
        }
      }
      if ((rc = msqg_reset_energy(M))) return fail(rc);
    }
    ev_out_t += P.dtout;
    if (!(ev_out_t <= P.tend + 1e-10)) ev_out_alive = 0;
  }
  if (!ev_out_alive) return 0;
  if (ev_out_t > g_t) tnext = ev_out_t;
  if (ev_flt_alive && ev_flt_t > g_t && ev_flt_t < tnext) tnext = ev_flt_t;
  double dt = 0., tn = 0.;
  if ((rc = msqg_step(M, g_t, tnext == HUGEV ? -1. : tnext, &dt, &tn))) return fail(rc);
  g_dt = dt; g_t = tn; g_i++;
  return 1;
}

int run(void) {
  int rc;
  /* defaults: set_vars(); init: qg.c init then set_const(); write_const: backup_config() */
  if ((rc = set_vars())) return rc;
  if ((rc = qg_init_event())) return rc;
  if ((rc = set_const())) return rc;
  if ((rc = backup_config())) return rc;
  qg_run_reset();
  while ((rc = qg_run_iteration(1)) > 0) {}
  if (rc < 0) return rc;
  trash_vars();
  return MSQG_OK;
}
