/*
 * layout.cuh -- HBM data layout of the msqg layer fields.
 *
 * A "list" (the reference's `scalar *`, msqg/qg.h:23-59) on multigrid level L is
 *     double [nf][ny+2][pitch]         nx = ny = 2^L on one GPU, x fastest
 * with one ghost ring ([BASILISK] allocates two, msqg only ever reads one).
 * Interior cell (x=0,y=0) of a plane sits at element  pitch + OX, OX = 16, so
 * every interior row starts 128-byte aligned; pitch is a multiple of 16 doubles.
 * This is the numpy (nl,N,N) = [layer][y][x] layout of pyset_field/pyget_field
 * (msqg/qg.h:1164-1189) plus padding; the reference's own AoS/y-fastest storage
 * is not observable through any interface.
 */
#pragma once
#include <cstddef>
#include <cstdint>

#define MSQG_OX 16
#define MSQG_MAXLEV 16
#define MSQG_FRAME 16 /* frame (deep halo capacity) of the planes of a tile, cells; <= MSQG_OX */
#define MSQG_NLMAX 12 /* device kernels are instantiated for 1..12 layers */

struct Geom {
  int nx, ny;   /* cells of this tile in x and y (single GPU: nx == ny == 2^level) */
  int bc;       /* bit mask of INTERNAL sides (1 left, 2 right, 4 bottom, 8 top): their ghosts come from a
                   neighbouring tile by halo exchange; physical sides get the boundary condition */
  int pitch;    /* doubles per row */
  size_t plane; /* doubles per scalar = (ny+2)*pitch */
  double Delta; /* L0/n */
  /* divisors used by the stencils and their correctly rounded reciprocals (for div_by) */
  double rD;          /* 1/Delta */
  double D2, rD2;     /* sq(Delta) */
  double D12, rD12;   /* 12.*Delta*Delta (jacobian macro) */
  double D2x, rD2x;   /* 2*Delta (beta_effect macro) */
};

static inline int msqg_pitch(int n) { return ((n + MSQG_OX + 1 + 15) / 16) * 16; }

#define GIDX(P, y, x) ((size_t)((y) + 1) * (size_t)(P) + (size_t)(MSQG_OX + (x)))
