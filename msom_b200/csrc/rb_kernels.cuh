/*
 * rb_kernels.cuh -- throughput smoother of the PV inversion: red-black ordering of relax_layer (sm_100a, fp64).
 *
 * The cell update is the reference's (msqg/poisson_layer.h:80-146: assemble the nl x nl tridiagonal system of one
 * column from the four horizontal neighbours, Thomas solve, write back in place; scalar Helmholtz variant
 * [BASILISK] poisson.h relax(), in-tree copy mspg/elliptic.h:294-301), expression by expression (-fmad=false,
 * divisions by the iterate-independent pivots through div_by(), which returns the IEEE quotient).  Only the
 * TRAVERSAL differs: a sweep is two half-sweeps, cells with (x + y) even ("red") first, then (x + y) odd ("black"),
 * each followed by boundary_level.  The reference documents that the result of its own sweep depends on traversal
 * order, OpenMP and MPI (poisson_layer.h:55-65); a red-black half-sweep only reads cells of the other colour, so its
 * result depends on none of them.  oracle/msqg_oracle.c restates this ordering (orc_set_smoother(m, 1)) and the
 * kernel is bit-exact against it.
 *
 * Temporal blocking: all ns <= 4 sweeps (nh = 2 ns half-sweeps) of a level are ONE pass over HBM.
 *   - a CTA owns an output block of TX = 128 - 2 nh columns x `rpc` rows and streams the rows of its 128-column
 *     window (output block + halo of nh cells on every side that has neighbours) through a shared-memory ring;
 *   - thread (k, p) applies half-sweep k to column pair p (columns 2p, 2p+1 of the window; exactly one of the two has
 *     the colour of half-sweep k in a given row); at step t it works on window row t - 1 - 2k, so half-sweep k runs two
 *     rows behind half-sweep k - 1 and all nh half-sweeps of a step are independent of each other: one
 *     __syncthreads() per step;
 *   - cells whose halo is incomplete (window edge) are computed anyway; the garbage moves inward one cell per
 *     half-sweep and reaches exactly the halo after nh of them -- no validity masks, only the output block is stored;
 *   - rows are fetched with cp.async RB_PF steps ahead and stored column-parity-split ([pair of layers][parity][pair]),
 *     so that every shared-memory access of a warp is a contiguous run of 16-byte words (conflict-free);
 *   - of the four neighbours only north and the one outside the pair are loaded: south and the in-pair neighbour are
 *     the values the thread loaded one step earlier (it moves one row up and switches column every step).
 * HBM traffic per pass: read da and res once (plus halo redundancy 128/TX), write da once, whatever ns is.  The pass
 * is OUT of place (da -> da_out): the output block of one CTA is halo of its neighbours, which must read the
 * pre-pass values; the host swaps the two buffers afterwards.
 * Homogeneous dirichlet ghosts of da are evaluated as -(value of the cell itself before its update), which is what
 * boundary_level left in the ghost ring (the ghost of a cell only mirrors that cell, and the cell only changes in its
 * own half-sweep).
 */
#pragma once
#include "mg_kernels.cuh"

#define RB_WX 128     /* window columns */
#define RB_NP 64      /* column pairs of the window = threads per half-sweep stage */
#define RB_PF 4       /* rows fetched ahead */
#define RB_NSMAX 4    /* sweeps fused in one pass */

struct RbArgs {
  const double *da;   /* in: [NL] planes of the level */
  double *da_out;     /* out: a DIFFERENT buffer of the same geometry (the output block of a CTA is halo of its neighbours) */
  const double *res;  /* rhs of the correction equation */
  Geom g;
  int ns;             /* sweeps of this pass, 1..RB_NSMAX */
  int R;              /* ring rows = 4 ns + 1 + RB_PF */
  int TX;             /* output columns per strip = RB_WX - 4 ns */
  int rpc;            /* output rows per chunk */
  int ox_lo, ox_hi, oy_lo, oy_hi; /* cells stored by this launch (own cells; a tile may add a ring of halo cells) */
  int xlo, xhi, ylo, yhi;         /* cells that exist in memory: own cells + deep halo on internal sides */
  int par0;           /* parity of the tile origin (x0 + y0) & 1 */
  const double *coef; /* RCOEF: per-row [ny][6][NL] or per-cell [ny][nx][6][NL] Thomas coefficients (k_rowcoef) */
  int coef_cell;
  int reuse;          /* 1: south / in-pair neighbours from registers (default); 0: always from shared memory */
};

__device__ __forceinline__ void cp_async8(unsigned smem_addr, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_addr), "l"(gsrc));
}

template <int NL>
struct RbCfg {
  static constexpr int NL2 = (NL + 1) / 2;          /* layer pairs (16-byte words per cell) */
  static constexpr int ARR2 = NL2 * RB_WX;          /* double2 per array per ring row */
  static constexpr int ROW2 = 2 * ARR2;             /* double2 per ring row: da, then res */
  static constexpr size_t row_bytes = (size_t)ROW2 * 16;
  /* sweeps per pass: bounded by the ring (4 ns + 1 + RB_PF rows <= 227 KB) and by registers (threads = 128 ns) */
  static constexpr int NSMAX = NL <= 4 ? RB_NSMAX : (NL <= 8 ? 2 : 1);
  static_assert((size_t)(4 * NSMAX + 1 + RB_PF) * row_bytes <= 227 * 1024, "ring does not fit");
};

template <int NL, bool RCOEF>
__global__ void __launch_bounds__(RbCfg<NL>::NSMAX * 2 * RB_NP, 1)
k_relax_rb(RbArgs A, RelaxCoef<NL> C) {
  using Cfg = RbCfg<NL>;
  constexpr int WX = RB_WX, NP = RB_NP, PF = RB_PF;
  constexpr int NL2 = Cfg::NL2, ARR2 = Cfg::ARR2, ROW2 = Cfg::ROW2;
  extern __shared__ double2 ring[];
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int nh = 2 * A.ns, H = nh, R = A.R;
  const int k = tid / NP, p = tid % NP;
  const int pitch = A.g.pitch, nx = A.g.nx, ny = A.g.ny, bc = A.g.bc;
  const size_t plane = A.g.plane;
  /* output block and window of this CTA */
  const int ox0 = A.ox_lo + blockIdx.x * A.TX;
  const int ox1 = min(ox0 + A.TX, A.ox_hi);
  const int x0w = ox0 - H;
  const int oy0 = A.oy_lo + blockIdx.y * A.rpc;
  const int oy1 = min(oy0 + A.rpc, A.oy_hi);
  const int jw0 = max(A.ylo, oy0 - H), jw1 = min(A.yhi, oy1 + H);
  const int nrw = jw1 - jw0; /* window rows */

  /* ---- streaming: thread <-> (window column wx, planes q0, q0 + qstep, ...) */
  const int wx = tid & (WX - 1), q0 = tid >> 7, qstep = nthreads >> 7;
  const int gxl = x0w + wx;
  const bool lvalid = gxl >= A.xlo && gxl < A.xhi;
  const bool ovalid = gxl >= ox0 && gxl < ox1;
  const int colofs = ((wx & 1) * NP + (wx >> 1)) * 2; /* doubles; parity-split column inside a layer-pair block */
  double *ring_d = reinterpret_cast<double *>(ring);
  const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring_d);

  auto load_row = [&](int r, int slot) {
    if (r < nrw && lvalid) {
      const long long go = (long long)(jw0 + r + 1) * pitch + MSQG_OX + gxl;
      const unsigned dst0 = ring_s + (unsigned)((slot * ROW2 * 2 + colofs) * 8);
      for (int q = q0; q < 2 * NL; q += qstep) {
        const int arr = q >= NL ? 1 : 0, l = arr ? q - NL : q;
        const double *src = (arr ? A.res : A.da) + (long long)l * (long long)plane + go;
        cp_async8(dst0 + (unsigned)((arr * 2 * ARR2 + (l >> 1) * 2 * WX + (l & 1)) * 8), src);
      }
    }
    cp_async_commit();
  };

  /* prologue: rows 0 .. PF-1 */
#pragma unroll 1
  for (int r = 0; r < PF; r++) load_row(r, r % R);
  int ld_slot = PF % R;                 /* slot of window row t + PF */
  int sl = ((-1 - 2 * k) % R + R) % R;  /* slot of this stage's row t - 1 - 2k */
  int out_slot = ((-2 * nh) % R + R) % R; /* slot of the row completed in the previous step, t - 2 nh */
  const int T = nrw + 2 * nh;

  double Ireg[NL], Nreg[NL]; /* carried: raw in-pair neighbour of the last step's row, raw north of the last step */
#pragma unroll
  for (int l = 0; l < NL; l++) { Ireg[l] = 0.; Nreg[l] = 0.; }
  bool have_prev = false;

#pragma unroll 1
  for (int t = 0; t < T; t++) {
    load_row(t + PF, ld_slot);
    ld_slot = ld_slot + 1 == R ? 0 : ld_slot + 1;

    /* ---- half-sweep k on window row r */
    const int r = t - 1 - 2 * k;
    if (r >= 0 && r < nrw) {
      const int j = jw0 + r;
      const int upd = (j + k + A.par0 + x0w) & 1; /* which column of the pair has the colour of half-sweep k in this row */
      const int gx = x0w + 2 * p + upd;
      const bool exist = gx >= A.xlo && gx < A.xhi;
      const int slN = sl + 1 == R ? 0 : sl + 1, slS = sl == 0 ? R - 1 : sl - 1;
      const double2 *rowc = ring + sl * ROW2, *rown = ring + slN * ROW2, *rows = ring + slS * ROW2;
      const int cu = upd * NP + p, ci = (1 - upd) * NP + p;
      int po = p + (upd ? 1 : -1);
      po = po < 0 ? 0 : (po > NP - 1 ? NP - 1 : po);
      const int co = (1 - upd) * NP + po;
      double b[NL], vi[NL], vo[NL], vn[NL], vs[NL];
      const bool reuse = A.reuse && have_prev;
#pragma unroll
      for (int lp = 0; lp < NL2; lp++) {
        const double2 tb = rowc[ARR2 + lp * WX + cu];
        const double2 tn = rown[lp * WX + cu];
        const double2 to = rowc[lp * WX + co];
        b[2 * lp] = tb.x; vn[2 * lp] = tn.x; vo[2 * lp] = to.x;
        if (2 * lp + 1 < NL) { b[2 * lp + 1] = tb.y; vn[2 * lp + 1] = tn.y; vo[2 * lp + 1] = to.y; }
      }
      if (reuse) {
#pragma unroll
        for (int l = 0; l < NL; l++) { vs[l] = Ireg[l]; vi[l] = Nreg[l]; }
      } else {
#pragma unroll
        for (int lp = 0; lp < NL2; lp++) {
          const double2 ti = rowc[lp * WX + ci];
          const double2 ts = rows[lp * WX + cu];
          vi[2 * lp] = ti.x; vs[2 * lp] = ts.x;
          if (2 * lp + 1 < NL) { vi[2 * lp + 1] = ti.y; vs[2 * lp + 1] = ts.y; }
        }
      }
#pragma unroll
      for (int l = 0; l < NL; l++) { Ireg[l] = vi[l]; Nreg[l] = vn[l]; }
      have_prev = true;
      /* homogeneous dirichlet ghosts on the physical sides: -(this cell before its update) */
      const bool gl = gx == 0 && !(bc & 1), gr = gx == nx - 1 && !(bc & 2);
      const bool gb = j == 0 && !(bc & 4), gt = j == ny - 1 && !(bc & 8);
      if (gl || gr || gb || gt) {
        /* west is the outer neighbour when the even column of the pair is updated, the in-pair one otherwise */
        const bool o_ghost = (gl && upd == 0) || (gr && upd == 1);
        const bool i_ghost = (gl && upd == 1) || (gr && upd == 0);
#pragma unroll
        for (int lp = 0; lp < NL2; lp++) {
          const double2 tc = rowc[lp * WX + cu];
          const double g0 = -tc.x, g1 = -tc.y;
          if (o_ghost) vo[2 * lp] = g0;
          if (i_ghost) vi[2 * lp] = g0;
          if (gb) vs[2 * lp] = g0;
          if (gt) vn[2 * lp] = g0;
          if (2 * lp + 1 < NL) {
            if (o_ghost) vo[2 * lp + 1] = g1;
            if (i_ghost) vi[2 * lp + 1] = g1;
            if (gb) vs[2 * lp + 1] = g1;
            if (gt) vn[2 * lp + 1] = g1;
          }
        }
      }
      /* relax_layer, poisson_layer.h:80-146 (same expression order as k_relax_lex) */
      double rhs[NL], out[NL];
      if (!RCOEF) {
#pragma unroll
        for (int l = 0; l < NL; l++) {
          double rr = C.msd2 * b[l];
          rr += vi[l] + vo[l];
          rr += vn[l] + vs[l];
          rhs[l] = rr;
        }
#pragma unroll
        for (int l = 1; l < NL; l++) rhs[l] -= div_by(C.t0[l] * rhs[l - 1], C.t1p[l - 1], C.rinv[l - 1]);
        out[NL - 1] = div_by(rhs[NL - 1], C.t1p[NL - 1], C.rinv[NL - 1]);
#pragma unroll
        for (int l = NL - 2; l >= 0; l--) out[l] = div_by(rhs[l] - C.t2[l] * out[l + 1], C.t1p[l], C.rinv[l]);
      } else {
        /* horizontally varying stretching: coefficients of this row / cell from the k_rowcoef table (own cells only) */
        const int jc = j < 0 ? 0 : (j > ny - 1 ? ny - 1 : j), xc = gx < 0 ? 0 : (gx > nx - 1 ? nx - 1 : gx);
        const double *ct = A.coef + (A.coef_cell ? (size_t)jc * nx + xc : (size_t)jc) * 6 * NL;
        double t0[NL], t2[NL], t1p[NL], rinv[NL];
#pragma unroll
        for (int l = 0; l < NL; l++) { t0[l] = ct[l]; t2[l] = ct[NL + l]; t1p[l] = ct[2 * NL + l]; rinv[l] = ct[3 * NL + l]; }
#pragma unroll
        for (int l = 0; l < NL; l++) {
          double rr = C.msd2 * b[l];
          rr += vi[l] + vo[l];
          rr += vn[l] + vs[l];
          rhs[l] = rr;
        }
#pragma unroll
        for (int l = 1; l < NL; l++) rhs[l] -= div_by(t0[l] * rhs[l - 1], t1p[l - 1], rinv[l - 1]);
        out[NL - 1] = div_by(rhs[NL - 1], t1p[NL - 1], rinv[NL - 1]);
#pragma unroll
        for (int l = NL - 2; l >= 0; l--) out[l] = div_by(rhs[l] - t2[l] * out[l + 1], t1p[l], rinv[l]);
      }
      if (exist) {
        double2 *wr = ring + sl * ROW2;
#pragma unroll
        for (int lp = 0; lp < NL2; lp++)
          wr[lp * WX + cu] = make_double2(out[2 * lp], (2 * lp + 1 < NL) ? out[2 * lp + 1] : 0.);
      }
    } else
      have_prev = false;
    sl = sl + 1 == R ? 0 : sl + 1;

    /* ---- store the row that the last half-sweep completed in the previous step */
    {
      const int ro = t - 2 * nh;
      if (ro >= 0 && ovalid) {
        const int j = jw0 + ro;
        if (j >= oy0 && j < oy1) {
          const long long go = (long long)(j + 1) * pitch + MSQG_OX + gxl;
          const double *srow = ring_d + (size_t)out_slot * ROW2 * 2 + colofs;
          for (int q = q0; q < NL; q += qstep)
            A.da_out[(long long)q * (long long)plane + go] = srow[(q >> 1) * 2 * WX + (q & 1)];
        }
      }
      out_slot = out_slot + 1 == R ? 0 : out_slot + 1;
    }
    cp_async_wait<PF - 1>();
    __syncthreads();
  }
  cp_async_wait<0>();
}
