/*
 * oracle/align_oracle.c -- CPU restatement of AmBiVErT's lib/align pairwise aligner.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path.
 * Nothing under ambivert_b200/ may import, link or call it; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against
 *   (1) the known-answer vectors of the reference's own tests
 *       (ambivert/tests/test_ambivert.py:111-192, via tests/golden/kat_reference_tests.json),
 *   (2) outputs of the reference C sources themselves, compiled unmodified into
 *       oracle/_ref/libalign_ref.so (tests/golden/fuzz_ref_vectors.json, made by
 *       oracle/make_golden.py), and, when oracle/_ref exists, a live differential fuzz.
 *
 * What is restated (all arithmetic is 32-bit signed integer):
 *   sa = first sequence (query, outer loop "col"), sb = second (target, inner loop "row").
 *   H  = best score of a cell, F = gap running down a column (consumes sb only, A_GAP),
 *   E[row] = gap running along a row (consumes sa only, B_GAP).
 *   jump[col*lb+row] : 0 = diagonal, <0 = -(A_GAP length), >0 = B_GAP length,
 *                      INT_MIN = local-alignment stop marker.
 *   Cell rule common to all modes (reference lib/align/global_align.c:180-209,
 *   local_align.c:162-206, glocal_align.c:139-168, align.c:141-169):
 *       h = diag + S(sa[col], sb[row]); jump = 0
 *       if (h < F)      { h = F;      jump = -(row - Fpos); }    diag wins ties over F
 *       if (h < E[row]) { h = E[row]; jump = col - Epos[row]; }  diag/F win ties over E
 *       F      <- (h+go >= F+ge)      ? open (pos=row) : extend     open wins ties
 *       E[row] <- (h+go >= E[row]+ge) ? open (pos=col) : extend
 */
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include "align_oracle.h"

#define NEG_HALF (INT_MIN / 2)

typedef struct {
  int la, lb, alpha, go, ge;
  const unsigned char *sa, *sb;
  const int *M;
  int *jump;  /* la*lb */
  int *Ha;    /* lb : H of previous column */
  int *Hb;    /* lb : H of current column  */
  int *E;     /* lb */
  int *Epos;  /* lb */
} dp_t;

static inline int sub(const dp_t *d, int col, int row) {
  return d->M[(int)d->sa[col] * d->alpha + (int)d->sb[row]];
}

/* choose the source of a cell: diagonal candidate, column gap F, row gap E */
static inline int choose(int cand, int F, int Fpos, int E, int Epos, int col, int row, int *jump) {
  int h = cand, j = 0;
  if (h < F) { h = F; j = -(row - Fpos); }
  if (h < E) { h = E; j = col - Epos; }
  *jump = j;
  return h;
}

/* advance one gap state past a cell whose score is h */
static inline void gap_advance(int h, int go, int ge, int *G, int *Gpos, int here) {
  if (h + go >= *G + ge) { *G = h + go; *Gpos = here; }
  else *G += ge;
}

static inline void swap_cols(dp_t *d) { int *t = d->Ha; d->Ha = d->Hb; d->Hb = t; }

/* ---- global: lib/align/global_align.c:136-217 -------------------------------------------- */
static int fill_global(dp_t *d, int *end_col, int *end_row) {
  const int go = d->go, ge = d->ge, lb = d->lb;
  for (int row = 0; row < lb; row++) {
    d->Ha[row] = go + ge * row;            /* virtual column -1 (:137) */
    d->Epos[row] = -1;
    d->E[row] = d->Ha[row] + go;           /* (:139) */
  }
  for (int col = 0; col < d->la; col++) {
    int Fpos = -1, F = 2 * go + ge * col;  /* (:145-146) */
    int *J = d->jump + (size_t)col * lb;
    for (int row = 0; row < lb; row++) {
      int diag = row ? d->Ha[row - 1] : (col ? go + ge * (col - 1) : 0);   /* (:148, :180) */
      int h = choose(diag + sub(d, col, row), F, Fpos, d->E[row], d->Epos[row], col, row, &J[row]);
      d->Hb[row] = h;
      gap_advance(h, go, ge, &F, &Fpos, row);
      gap_advance(h, go, ge, &d->E[row], &d->Epos[row], col);
    }
    swap_cols(d);
  }
  *end_col = d->la - 1; *end_row = lb - 1;
  return d->Ha[lb - 1];                    /* (:217) */
}

/* ---- local: lib/align/local_align.c:119-214 ---------------------------------------------- */
static int fill_local(dp_t *d, int *end_col, int *end_row) {
  const int go = d->go, ge = d->ge, lb = d->lb;
  int best = 0, bc = 0, br = 0;            /* (:119-121) */
  for (int row = 0; row < lb; row++) { d->Ha[row] = 0; d->Epos[row] = -1; d->E[row] = go; }
  for (int col = 0; col < d->la; col++) {
    int Fpos = -1, F = go;                 /* (:134-135) */
    int *J = d->jump + (size_t)col * lb;
    for (int row = 0; row < lb; row++) {
      int diag = row ? d->Ha[row - 1] : 0; /* (:137, :182) */
      int h = choose(diag + sub(d, col, row), F, Fpos, d->E[row], d->Epos[row], col, row, &J[row]);
      if (h < 0) { h = 0; J[row] = INT_MIN; }             /* clamp + stop marker (:150-152) */
      else if (h >= best) { best = h; bc = col; br = row; } /* last maximum wins (:156-160) */
      d->Hb[row] = h;
      gap_advance(h, go, ge, &F, &Fpos, row);
      gap_advance(h, go, ge, &d->E[row], &d->Epos[row], col);
    }
    swap_cols(d);
  }
  *end_col = bc; *end_row = br;
  return best;
}

/* ---- glocal: lib/align/glocal_align.c:115-191 (requires lb >= 2; lb==1 is UB upstream) ----- */
static int fill_glocal(dp_t *d, int *end_col, int *end_row) {
  const int go = d->go, ge = d->ge, lb = d->lb;
  int best = 0, bc = 0, br = 0;            /* (:115-117) */
  for (int row = 0; row < lb; row++) { d->Ha[row] = NEG_HALF; d->Epos[row] = -1; d->E[row] = NEG_HALF; }
  for (int col = 0; col < d->la; col++) {
    int *J = d->jump + (size_t)col * lb;
    int h = sub(d, col, 0);                /* first target row is a forced match (:128-136) */
    J[0] = 0; d->Hb[0] = h;
    int F = h + go, Fpos = 0;
    int row;
    for (row = 1; row < lb - 1; row++) {   /* (:138-169) */
      h = choose(d->Ha[row - 1] + sub(d, col, row), F, Fpos, d->E[row], d->Epos[row], col, row, &J[row]);
      d->Hb[row] = h;
      gap_advance(h, go, ge, &F, &Fpos, row);
      gap_advance(h, go, ge, &d->E[row], &d->Epos[row], col);
    }
    h = d->Ha[row - 1] + sub(d, col, row); /* last target row is a forced match (:172-176) */
    J[row] = 0; d->Hb[row] = h;
    if (h >= best) { best = h; bc = col; br = row; }      /* (:179-183) */
    swap_cols(d);
  }
  *end_col = bc; *end_row = br;
  return best;
}

/* ---- overlap ("align"): lib/align/align.c:109-232 (requires la >= 2; la==1 is UB upstream) - */
static int fill_overlap(dp_t *d, int *end_col, int *end_row) {
  const int go = d->go, ge = d->ge, lb = d->lb, la = d->la;
  int best = INT_MIN, bc = -1, br = -1;    /* (:109-111) */
  int h = 0;
  /* first column (:114-125) */
  for (int row = 0; row < lb; row++) {
    h = sub(d, 0, row);
    d->Hb[row] = h; d->Epos[row] = 0; d->E[row] = h + go; d->jump[row] = 0;
  }
  if (h >= best) { best = h; bc = 0; br = lb - 1; }
  swap_cols(d);
  /* interior columns (:131-180): a gap may only be re-opened from a cell that was not itself
     reached through the same kind of gap */
  int col;
  for (col = 1; col < la - 1; col++) {
    int *J = d->jump + (size_t)col * lb;
    h = sub(d, col, 0);
    d->Hb[0] = h; J[0] = 0;
    int Fpos = 0, F = h + go;
    for (int row = 1; row < lb; row++) {
      h = choose(d->Ha[row - 1] + sub(d, col, row), F, Fpos, d->E[row], d->Epos[row], col, row, &J[row]);
      d->Hb[row] = h;
      if (J[row] >= 0 && h + go >= F + ge) { F = h + go; Fpos = row; } else F += ge;
      if (J[row] <= 0 && h + go >= d->E[row] + ge) { d->E[row] = h + go; d->Epos[row] = col; } else d->E[row] += ge;
    }
    if (h >= best) { best = h; bc = col; br = lb - 1; }   /* bottom cell of the column (:172-176) */
    swap_cols(d);
  }
  /* last column (:183-230): every cell is an admissible end point, plain gap updates */
  {
    int *J = d->jump + (size_t)col * lb;
    h = sub(d, col, 0);
    d->Hb[0] = h; J[0] = 0;
    int Fpos = 0, F = h + go;
    if (h >= best) { best = h; bc = la - 1; br = 0; }
    for (int row = 1; row < lb; row++) {
      h = choose(d->Ha[row - 1] + sub(d, col, row), F, Fpos, d->E[row], d->Epos[row], col, row, &J[row]);
      d->Hb[row] = h;
      if (h >= best) { best = h; bc = la - 1; br = row; }
      gap_advance(h, go, ge, &F, &Fpos, row);
      gap_advance(h, go, ge, &d->E[row], &d->Epos[row], col);
    }
  }
  *end_col = bc; *end_row = br;
  return best;
}

/* ---- traceback: global_align.c:25-92, local_align.c:23-72, glocal_align.c:25-70, align.c:25-65.
 * Emits fragments walking backwards; caller reverses.  Returns the number of fragments. */
static int walk_back(const dp_t *d, int mode, int col, int row, orc_frag *out, int cap) {
  int n = 0;
  const int lb = d->lb;
  while (col >= 0 && row >= 0) {
    int j = d->jump[(size_t)col * lb + row];
    if (mode == ORC_LOCAL && j == INT_MIN) break;
    orc_frag f;
    if (j < 0)      { row += j; f.type = ORC_A_GAP; f.hsp_len = -j; }
    else if (j > 0) { col -= j; f.type = ORC_B_GAP; f.hsp_len = j; }
    else {
      int run = 0;
      do { --col; --row; ++run; }
      while (col >= 0 && row >= 0 && d->jump[(size_t)col * lb + row] == 0);
      f.type = ORC_MATCH; f.hsp_len = run;
    }
    f.sa_start = col + 1; f.sb_start = row + 1;
    if (n < cap) out[n] = f;
    n++;
  }
  if (mode == ORC_GLOBAL && (col >= 0 || row >= 0)) {   /* leading end gap (global_align.c:69-86) */
    orc_frag f; f.sa_start = 0; f.sb_start = 0;
    if (col >= 0) { f.type = ORC_B_GAP; f.hsp_len = col + 1; }
    else          { f.type = ORC_A_GAP; f.hsp_len = row + 1; }
    if (n < cap) out[n] = f;
    n++;
  }
  return n;
}

int orc_align_raw(int mode, const unsigned char *sa, int la, const unsigned char *sb, int lb,
                  int alpha_len, const int *matrix, int gap_open, int gap_extend,
                  int *score, orc_frag *frags, int cap) {
  if (la <= 0 || lb <= 0 || !sa || !sb || !matrix || gap_open > 0 || gap_extend > 0) return ORC_ERR_ARGS;
  if (mode == ORC_GLOCAL && lb < 2) return ORC_ERR_UNDEFINED;
  if (mode == ORC_OVERLAP && la < 2) return ORC_ERR_UNDEFINED;
  if (mode < 0 || mode > 3) return ORC_ERR_ARGS;

  dp_t d; memset(&d, 0, sizeof d);
  d.la = la; d.lb = lb; d.alpha = alpha_len; d.go = gap_open; d.ge = gap_extend;
  d.sa = sa; d.sb = sb; d.M = matrix;
  d.jump = (int *)malloc((size_t)la * lb * sizeof(int));
  d.Ha = (int *)malloc((size_t)lb * sizeof(int));
  d.Hb = (int *)malloc((size_t)lb * sizeof(int));
  d.E = (int *)malloc((size_t)lb * sizeof(int));
  d.Epos = (int *)malloc((size_t)lb * sizeof(int));
  int n = ORC_ERR_ARGS;
  if (d.jump && d.Ha && d.Hb && d.E && d.Epos) {
    int ec = 0, er = 0, s;
    switch (mode) {
      case ORC_GLOBAL: s = fill_global(&d, &ec, &er); break;
      case ORC_LOCAL:  s = fill_local(&d, &ec, &er); break;
      case ORC_GLOCAL: s = fill_glocal(&d, &ec, &er); break;
      default:         s = fill_overlap(&d, &ec, &er); break;
    }
    if (score) *score = s;
    n = walk_back(&d, mode, ec, er, frags, frags ? cap : 0);
    int m = n < cap ? n : cap;              /* reverse into sequence order */
    for (int i = 0; frags && i < m / 2; i++) { orc_frag t = frags[i]; frags[i] = frags[m - 1 - i]; frags[m - 1 - i] = t; }
    if (n > cap && frags) {
      /* caller's buffer was too small: the tail (earliest fragments) was dropped; report count only */
    }
  }
  free(d.jump); free(d.Ha); free(d.Hb); free(d.E); free(d.Epos);
  return n;
}

/* to_raw (lib/align/helpers.h:26-32) with the byte taken as unsigned */
void orc_encode(const char *in, unsigned char *out, int n, const unsigned char *map256) {
  for (int i = 0; i < n; i++) out[i] = map256[(unsigned char)in[i]];
}

int orc_align(int mode, const char *seqa, int la, const char *seqb, int lb,
              int alpha_len, const unsigned char *map256, const int *matrix,
              int gap_open, int gap_extend, int *score, orc_frag *frags, int cap) {
  if (la <= 0 || lb <= 0 || !seqa || !seqb || !map256) return ORC_ERR_ARGS;
  unsigned char *a = (unsigned char *)malloc((size_t)la), *b = (unsigned char *)malloc((size_t)lb);
  int n = ORC_ERR_ARGS;
  if (a && b) {
    orc_encode(seqa, a, la, map256); orc_encode(seqb, b, lb, map256);
    n = orc_align_raw(mode, a, la, b, lb, alpha_len, matrix, gap_open, gap_extend, score, frags, cap);
  }
  free(a); free(b);
  return n;
}

/* Best target per query: the selection loop of ambivert/ambivert.py:478-485 (seed -1000000) and
 * :504-511 (seed 0): strict '>' in target order, so the first maximum wins; idx -1 = no hit. */
int orc_best_hits(int mode,
                  const unsigned char *q, const long long *q_off, int n_q,
                  const unsigned char *t, const long long *t_off, int n_t,
                  int alpha_len, const int *matrix, int gap_open, int gap_extend,
                  int seed_score, int *best_score, int *best_idx) {
  for (int i = 0; i < n_q; i++) {
    int bs = seed_score, bi = -1;
    for (int j = 0; j < n_t; j++) {
      int s = 0;
      int rc = orc_align_raw(mode, q + q_off[i], (int)(q_off[i + 1] - q_off[i]),
                             t + t_off[j], (int)(t_off[j + 1] - t_off[j]),
                             alpha_len, matrix, gap_open, gap_extend, &s, NULL, 0);
      if (rc < 0) return rc;
      if (s > bs) { bs = s; bi = j; }
    }
    best_score[i] = bs; best_idx[i] = bi;
  }
  return 0;
}
