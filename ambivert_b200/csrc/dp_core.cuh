// dp_core.cuh -- the per-cell recurrences, direction encoding and traceback walk shared by the
// sm_100a kernels (kernels.cu).  Everything here is a small __host__ __device__ inline function so
// that tests/emu can drive the very same cell logic lane by lane on a CPU and compare it with the
// oracle before any GPU time is spent; the shipped library only ever runs it on the device.
//
// Notation follows the reference: sa = query, indexed by col; sb = target,
// indexed by row; H = cell score; F = gap running down a column (consumes sb, A_GAP); E = gap
// running along a row (consumes sa, B_GAP).  Reference recurrences:
//   lib/align/global_align.c:136-217, local_align.c:119-214, glocal_align.c:115-191, align.c:109-232.
#pragma once
#include <limits.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define ABV_HD __host__ __device__ __forceinline__
#else
#define ABV_HD inline
#endif

namespace abv {

enum Mode { OVERLAP = 0, LOCAL = 1, GLOBAL = 2, GLOCAL = 3 };
enum { FRAG_A_GAP = 0, FRAG_B_GAP = 1, FRAG_MATCH = 2 };

constexpr int NEG_HALF = INT_MIN / 2;   // glocal_align.c:120-122 "minus infinity"

// ---------------------------------------------------------------------------------------------
// Direction nibble.  The reference stores a jump length per cell (0 diag, <0 A_GAP length,
// >0 B_GAP length, INT_MIN local stop; e.g. local_align.c:183-197).  The jump length of a gap
// is "distance back to where that gap state was last opened", so it is recoverable from one
// open/extend bit per gap state per cell: 4 bits per cell instead of 32.
constexpr unsigned SRC_DIAG = 0, SRC_F = 1, SRC_E = 2, SRC_STOP = 3, SRC_MASK = 3;
constexpr unsigned BIT_FOPEN = 4;   // after this cell the column gap restarts here (vgap_pos = row)
constexpr unsigned BIT_EOPEN = 8;   // after this cell the row gap restarts here    (hgap_pos = col)

// End-cell candidate of one lane (one column).  Cross-lane order is (score, col): the reference
// scans column-major with '>=' so among equal maxima the last scanned cell wins.
struct EndCell {
  int score, col, row;
};

ABV_HD bool end_better(const EndCell &a, const EndCell &b) {   // a replaces b ?
  return a.score > b.score || (a.score == b.score && a.col > b.col);
}

// ---------------------------------------------------------------------------------------------
// One cell of the 32-bit wavefront kernel.
//   c,r      : column (query position) and row (target position)
//   sub      : score_matrix[sa[c]*alpha + sb[r]]
//   hdiag    : H[c-1][r-1] as handed over by the left neighbour (for c==0 the caller passes the
//              mode's virtual column -1); row 0 is overridden here where the mode defines it
//   e_in     : E[r] entering column c
//   F        : column gap state entering row r (lane-resident, updated in place)
//   h,e_out  : H[c][r] and E[r] leaving column c
//   cand     : lane's end-cell candidate (updated per the mode's rule)
// returns the direction nibble.
template <int MODE>
ABV_HD unsigned wf32_cell(int c, int r, int la, int lb, int sub, int hdiag, int e_in, int &F,
                          int go, int ge, int &h, int &e_out, EndCell &cand) {
  unsigned nib;
  if (MODE == GLOBAL) {
    if (r == 0) {                       // global_align.c:145-148
      hdiag = c ? go + ge * (c - 1) : 0;
      F = 2 * go + ge * c;
    }
  } else if (MODE == LOCAL) {
    if (r == 0) {                       // local_align.c:134-137
      hdiag = 0;
      F = go;
    }
  } else if (MODE == GLOCAL) {
    if (r == 0) {                       // first target row: forced match (glocal_align.c:128-136)
      h = sub;
      F = h + go;
      e_out = e_in;
      return SRC_DIAG | BIT_FOPEN;
    }
    if (r == lb - 1) {                  // last target row: forced match (glocal_align.c:172-183)
      h = hdiag + sub;
      e_out = e_in;
      if (h >= 0) { cand.score = h; cand.col = c; cand.row = r; }
      return SRC_DIAG;
    }
  } else {                              // OVERLAP
    if (c == 0) {                       // first column (align.c:114-125)
      h = sub;
      e_out = h + go;
      if (r == lb - 1) { cand.score = h; cand.col = c; cand.row = r; }
      return SRC_DIAG | BIT_EOPEN;
    }
    if (r == 0) {                       // top cell of later columns (align.c:135-138, :185-192)
      h = sub;
      F = h + go;
      e_out = e_in;
      if (c == la - 1 || lb == 1) { cand.score = h; cand.col = c; cand.row = 0; }
      return SRC_DIAG | BIT_FOPEN;
    }
  }

  // the common choice: diagonal wins ties over F, both win ties over E (global_align.c:180-192)
  int s = hdiag + sub;
  unsigned src = SRC_DIAG;
  if (s < F) { s = F; src = SRC_F; }
  if (s < e_in) { s = e_in; src = SRC_E; }

  if (MODE == LOCAL) {
    if (s < 0) { s = 0; src = SRC_STOP; }                       // local_align.c:150-152
    else if (s >= cand.score) { cand.score = s; cand.col = c; cand.row = r; }   // :156-160
  }
  if (MODE == OVERLAP) {
    if (c == la - 1) {                                          // align.c:211-215
      if (s >= cand.score) { cand.score = s; cand.col = c; cand.row = r; }
    } else if (r == lb - 1) {                                   // align.c:172-176
      cand.score = s; cand.col = c; cand.row = r;
    }
  }

  // gap state updates: opening wins ties over extending (global_align.c:198-209)
  bool fo = (s + go >= F + ge), eo = (s + go >= e_in + ge);
  if (MODE == OVERLAP && c < la - 1) {                          // guarded re-open (align.c:157, :164)
    fo = fo && (src != SRC_F);
    eo = eo && (src != SRC_E);
  }
  F = fo ? s + go : F + ge;
  e_out = eo ? s + go : e_in + ge;
  h = s;
  nib = src | (fo ? BIT_FOPEN : 0u) | (eo ? BIT_EOPEN : 0u);
  return nib;
}

// virtual column -1 handed to column 0: H[-1][r] and E[r] entering column 0
template <int MODE>
ABV_HD void wf32_virtual_column(int r, int go, int ge, int &h_left, int &e_in) {
  if (MODE == GLOBAL) { h_left = go + ge * r; e_in = h_left + go; }        // global_align.c:136-140
  else if (MODE == LOCAL) { h_left = 0; e_in = go; }                       // local_align.c:123-128
  else if (MODE == GLOCAL) { h_left = NEG_HALF; e_in = NEG_HALF; }         // glocal_align.c:119-123
  else { h_left = 0; e_in = 0; }                                           // overlap: column 0 ignores both
}

template <int MODE>
ABV_HD EndCell wf32_initial_candidate() {
  EndCell e;
  e.score = (MODE == OVERLAP) ? INT_MIN : -1;   // local/glocal: -1 = "no cell reached best >= 0"
  e.col = -1;
  e.row = -1;
  return e;
}

// final (score, end cell) of a pair from the warp-wide best candidate / the corner cell
template <int MODE>
ABV_HD void wf32_finish(const EndCell &best, int corner_h, int la, int lb, int &score, int &ec, int &er) {
  if (MODE == GLOBAL) { score = corner_h; ec = la - 1; er = lb - 1; }      // global_align.c:217
  else if (MODE == OVERLAP) { score = best.score; ec = best.col; er = best.row; }
  else if (best.score < 0 || best.col < 0) { score = 0; ec = 0; er = 0; }  // max starts at 0,(0,0)
  else { score = best.score; ec = best.col; er = best.row; }
}

// ---------------------------------------------------------------------------------------------
// Direction matrix layout of the wavefront kernel: the query is cut into blocks of 32 columns,
// lane t owns column 32*b+t and at step s works on row s-t.  Each lane packs the nibbles of 8
// consecutive steps into one 32-bit word, so a warp writes one coalesced 128-byte line per 8
// steps:   word(b, s, t) = (b*steps8 + s/8)*32 + t ,  nibble s%8.
ABV_HD int dir_steps8(int lb) { return (lb + 31 + 7) >> 3; }
ABV_HD long long dir_words(int la, int lb) { return (long long)((la + 31) >> 5) * dir_steps8(lb) * 32; }

struct DirReader {
  const uint32_t *w;
  int steps8;
  ABV_HD unsigned get(int col, int row) const {
    int b = col >> 5, t = col & 31, s = row + t;
    uint32_t v = w[((long long)b * steps8 + (s >> 3)) * 32 + t];
    return (v >> ((s & 7) * 4)) & 15u;
  }
};

struct FragOut {
  int type, sa_start, sb_start, hsp_len;
};

// Traceback (global_align.c:25-92, local_align.c:23-72, glocal_align.c:25-70, align.c:25-65).
// Fragments are found in REVERSE sequence order; emit(j, frag) receives the j-th one found.  Returns how
// many there are.
template <int MODE, class Reader, class Emit>
ABV_HD int walk_back_emit(const Reader &rd, int col, int row, Emit emit) {
  int n = 0;
  while (col >= 0 && row >= 0) {
    unsigned nib = rd.get(col, row);
    unsigned src = nib & SRC_MASK;
    if (MODE == LOCAL && src == SRC_STOP) break;
    FragOut f;
    if (src == SRC_F) {              // jump = -(row - vgap_pos): back to where F was last opened
      int p = row - 1;
      while (p >= 0 && !(rd.get(col, p) & BIT_FOPEN)) --p;
      f.type = FRAG_A_GAP; f.hsp_len = row - p;
      row = p;
    } else if (src == SRC_E) {       // jump = col - hgap_pos[row]
      int p = col - 1;
      while (p >= 0 && !(rd.get(p, row) & BIT_EOPEN)) --p;
      f.type = FRAG_B_GAP; f.hsp_len = col - p;
      col = p;
    } else {                         // diagonal run
      int run = 0;
      do { --col; --row; ++run; }
      while (col >= 0 && row >= 0 && (rd.get(col, row) & SRC_MASK) == SRC_DIAG);
      f.type = FRAG_MATCH; f.hsp_len = run;
    }
    f.sa_start = col + 1; f.sb_start = row + 1;
    emit(n, f);
    ++n;
  }
  if (MODE == GLOBAL && (col >= 0 || row >= 0)) {   // leading end gap (global_align.c:69-86)
    FragOut f; f.sa_start = 0; f.sb_start = 0;
    if (col >= 0) { f.type = FRAG_B_GAP; f.hsp_len = col + 1; }
    else          { f.type = FRAG_A_GAP; f.hsp_len = row + 1; }
    emit(n, f);
    ++n;
  }
  return n;
}

struct FragStore {      // out[j] = j-th fragment found, while j < cap
  FragOut *out; int cap;
  ABV_HD void operator()(int j, const FragOut &f) const { if (j < cap) out[j] = f; }
};

// Writes fragments in REVERSE sequence order into out[0..cap) and returns how many there are (which may
// exceed cap; the excess is counted but not stored).
template <int MODE, class Reader>
ABV_HD int walk_back(const Reader &rd, int col, int row, FragOut *out, int cap) {
  FragStore st; st.out = out; st.cap = cap;
  return walk_back_emit<MODE>(rd, col, row, st);
}

// ---------------------------------------------------------------------------------------------
// Packed 16-bit arithmetic of the many-to-many score-only kernel: each 32-bit word carries the
// same DP cell of TWO alignments (one query against two targets).  On sm_100a these are single
// instructions (VIADD.16x2, VIMNMX.S16x2, VIMNMX3.S16x2, VIADDMNMX.S16x2[.RELU]).
namespace p16 {

ABV_HD uint32_t pack(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
ABV_HD int lo(uint32_t v) { return (int)(int16_t)(v & 0xffffu); }
ABV_HD int hi(uint32_t v) { return (int)(int16_t)(v >> 16); }

#if defined(__CUDA_ARCH__)
ABV_HD uint32_t add(uint32_t a, uint32_t b) { return __vadd2(a, b); }
ABV_HD uint32_t max2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
ABV_HD uint32_t max3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
ABV_HD uint32_t max3_relu(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2_relu(a, b, c); }
ABV_HD uint32_t addmax(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }   // max(a+b, c)
#else
ABV_HD int16_t w16(int v) { return (int16_t)(uint16_t)(v & 0xffff); }   // wrap like the hardware
ABV_HD uint32_t add(uint32_t a, uint32_t b) { return pack(w16(lo(a) + lo(b)), w16(hi(a) + hi(b))); }
ABV_HD int mx(int a, int b) { return a > b ? a : b; }
ABV_HD uint32_t max2(uint32_t a, uint32_t b) { return pack(mx(lo(a), lo(b)), mx(hi(a), hi(b))); }
ABV_HD uint32_t max3(uint32_t a, uint32_t b, uint32_t c) { return max2(max2(a, b), c); }
ABV_HD uint32_t max3_relu(uint32_t a, uint32_t b, uint32_t c) { return max2(max3(a, b, c), 0u); }
ABV_HD uint32_t addmax(uint32_t a, uint32_t b, uint32_t c) { return max2(add(a, b), c); }
#endif

// "minus infinity" of the packed kernel and the bound that makes it exact: every true DP value v
// of a pair satisfies |v| <= P16_LIMIT, so NEG16 + anything reachable never wins a max against a
// true value and never wraps.
constexpr int NEG16 = -16000;
constexpr int P16_LIMIT = 8000;

}  // namespace p16

// Layout of the per-query substitution profile in shared memory.  Lane t of the warp owns query
// columns [t*K, (t+1)*K).  Its K packed substitution scores for one target-pair symbol are read
// with K/V vector loads of V words; chunk i of lane t sits at word (i*32 + t)*V, so the 8 (V=4)
// or 16 (V=2) lanes of one shared-memory phase touch 32 consecutive banks whatever symbol each
// lane is looking up (the row stride 32*K words is a multiple of 32 banks).
// One row of one lane of the packed kernel, plain Gotoh form (5 packed ops per cell pair):
//   d   : in  H[c0-1][r-1]                          (diag of the lane's first column)
//   e   : in  E[r] entering c0; out E[r] leaving the strip
//   Hp  : in  H[.][r-1]; out H[.][r]
//   F   : column gap states
//   S   : packed substitution scores of the K columns for this row's target symbols
template <int K, bool RELU, bool TRACK_MAX>
ABV_HD void mm16_row(uint32_t d, uint32_t &e, uint32_t (&Hp)[K], uint32_t (&F)[K], const uint32_t (&S)[K],
                     uint32_t go2, uint32_t ge2, uint32_t &best) {
#pragma unroll
  for (int k = 0; k < K; ++k) {
    uint32_t dd = p16::add(d, S[k]);
    d = Hp[k];
    uint32_t h = RELU ? p16::max3_relu(dd, e, F[k]) : p16::max3(dd, e, F[k]);
    uint32_t hg = p16::add(h, go2);
    e = p16::addmax(e, ge2, hg);
    F[k] = p16::addmax(F[k], ge2, hg);
    Hp[k] = h;
    if (TRACK_MAX) best = p16::max2(best, h);
  }
}

// glocal rows on which at least one half is on its last target row: that half takes the forced
// diagonal (glocal_align.c:172-176), the other half the plain update.  `last_mask` has 0xffff in
// the halves that are on their last row.  colmask[k] is 0xffff.. for real query columns.
template <int K>
ABV_HD void mm16_row_glocal_last(uint32_t d, uint32_t &e, uint32_t (&Hp)[K], uint32_t (&F)[K], const uint32_t (&S)[K],
                                 uint32_t go2, uint32_t ge2, uint32_t last_mask, int c0, int lq, uint32_t &best) {
#pragma unroll
  for (int k = 0; k < K; ++k) {
    uint32_t dd = p16::add(d, S[k]);
    d = Hp[k];
    uint32_t hn = p16::max3(dd, e, F[k]);
    uint32_t h = (dd & last_mask) | (hn & ~last_mask);
    uint32_t hg = p16::add(h, go2);
    e = p16::addmax(e, ge2, hg);
    F[k] = p16::addmax(F[k], ge2, hg);
    Hp[k] = h;
    if (c0 + k < lq) {   // only real query columns may end the alignment
      uint32_t cand = (h & last_mask) | (p16::pack(p16::NEG16, p16::NEG16) & ~last_mask);
      best = p16::max2(best, cand);
    }
  }
}


// Overlap mode (align.c:68-243) in the packed kernel, score only.  The cells follow the common recurrences except
//   * the first query column is forced, H[0][r] = S, and opens the row gap with H + go (align.c:114-125): the lane that owns it
//     takes diagonal 0 and "minus infinity" for both gap states on every row;
//   * the first target row is forced, H[c][0] = S, F = H + go (align.c:135-138); its row gap only feeds forced cells;
//   * gap states are not re-opened from themselves before the last column (align.c:157, :164): for go <= ge that never changes a
//     value (re-opening from the state's own value loses or ties against extending it), which is what the host requires;
//   * the score is the best cell of the last query column (any row) or of the last target row (any column) (align.c:172-176,
//     :185-192, :211-215).  With two targets per word the rows of the shorter one end earlier: rowmask has 0xffff in the halves
//     whose target still has row r, lastmask in those for which r is the last row; kstar >= 0 on the lane that owns column lq-1.
template <int K>
ABV_HD void mm16_overlap_cands(const uint32_t (&Hp)[K], int c0, int lq, int kstar, uint32_t rowmask, uint32_t lastmask, uint32_t &best) {
  const uint32_t NEG2 = p16::pack(p16::NEG16, p16::NEG16);
  if (kstar >= 0) {
    uint32_t h = Hp[0];
#pragma unroll
    for (int k = 1; k < K; ++k) h = (k == kstar) ? Hp[k] : h;
    best = p16::max2(best, (h & rowmask) | (NEG2 & ~rowmask));
  }
  if (lastmask) {
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (c0 + k < lq) best = p16::max2(best, (Hp[k] & lastmask) | (NEG2 & ~lastmask));
  }
}

// ---------------------------------------------------------------------------------------------
// Biased packed form used by the streaming global kernel (mm16g_kernel).  Every score-like value
// (H, E, F) is stored as true + BIAS in each half, so both halves are always positive.  Then
//   * adding a small signed term to both halves is ONE ordinary 32-bit add of the "arithmetically
//     packed" addend hi*65536 + lo: no borrow can cross the half boundary because the low half stays
//     in [0, 65535].  Plain IADD/IMAD can issue on the FMA pipe, leaving the half-rate ALU/DPX pipe
//     to the three max-type instructions of a cell;
//   * max-type DPX instructions compare the biased halves as signed 16-bit (all in [0, 32767]).
namespace b16 {
constexpr int BIAS = 16384;
constexpr int LIMIT = 16000;      // |true value| bound that keeps every biased half in [384, 32384]
ABV_HD uint32_t biased(int lo, int hi) { return p16::pack(lo + BIAS, hi + BIAS); }
ABV_HD uint32_t addend(int lo, int hi) { return (uint32_t)(hi * 65536 + lo); }
ABV_HD int lo(uint32_t v) { return p16::lo(v) - BIAS; }
ABV_HD int hi(uint32_t v) { return p16::hi(v) - BIAS; }
}  // namespace b16

// Drift-normalised recurrences of the streaming global kernel.  With g = gap_extend, every value of
// cell (c, r) is stored minus g*(c + r):  H~ = H - g(c+r), E~ = E - g(c+r), F~ = F - g(c+r).  Moving one
// cell right or down changes the offset by exactly g, which cancels the "+ gap_extend" of a gap
// extension, and a diagonal move changes it by 2g, which is folded into the substitution profile:
//     H~[c][r]   = max(H~[c-1][r-1] + (S - 2g), E~[c][r], F~[c][r])
//     E~[c+1][r] = max(E~[c][r], H~[c][r] + (go - g))          F~[c][r+1] likewise
// One ordinary add and three DPX instructions per cell pair instead of two adds and three DPX, and a
// two-instruction dependent chain.  The global boundary becomes constant: H~[c][-1] = H~[-1][r] = go + g,
// F~[c][0] = E~[0][r] = 2go, H~[-1][-1] = 2g (from global_align.c:136-148).
//   S[]   : in  addend-packed (S - 2g) of this row's symbols; scratch
//   hleft : H~[c0-1][r-1]       e : in E~[r] entering the strip, out E~[r] leaving it
//   gop2  : p16::pack(go - g, go - g)
// K = query columns per lane; the profile is stored with KP >= K words per lane (StripCfg) so that odd
// K still loads with 16-byte vectors; S[] has KP entries of which the first K are used.
template <int K>
struct StripCfg {
  static constexpr int KP = (K % 2) ? ((K + 3) & ~3) : K;   // padded strip width in the profile
  static constexpr int V = (KP % 4 == 0) ? 4 : 2;           // words per shared-memory vector load
  static constexpr int P = 32 * KP;                         // profile row length in words
};

#ifndef ABV_ROW_ADDS
#define ABV_ROW_ADDS 0
#endif
#ifndef ABV_MAX3_ORDER
#define ABV_MAX3_ORDER 1      // max3(S, F, e) and the F update before the e update: cheapest of 108 scanned forms at unroll 8 for K = 7, 8, 9,
#endif                        // global and glocal (static model), measured +1.0 % global / +0.8 % glocal against max3(S, e, F), e first
#ifndef ABV_FE_ORDER
#define ABV_FE_ORDER 1
#endif
template <int K, int KP>
ABV_HD void mm16d_row(uint32_t hleft, uint32_t &e, uint32_t (&Hp)[K], uint32_t (&F)[K], uint32_t (&S)[KP], uint32_t gop2, uint32_t gope) {
#if ABV_ROW_ADDS == 2
  uint32_t Hn[K];
#endif
  // The statement and operand orders below are semantically equivalent; they steer ptxas' register allocation, i.e. the number of
  // moves and register-bank conflicts of the steady loop (tools/sass_cost.py; scanned with tools/scan_row_variants.sh).
#if ABV_ROW_ADDS == 1
  S[0] += hleft;
#pragma unroll
  for (int k = 1; k < K; ++k) S[k] += Hp[k - 1];
#elif ABV_ROW_ADDS == 0
#pragma unroll
  for (int k = K - 1; k >= 1; --k) S[k] += Hp[k - 1];
  S[0] += hleft;
#endif
#pragma unroll
  for (int k = 0; k < K; ++k) {
#if ABV_ROW_ADDS == 2
    S[k] += k ? Hp[k - 1] : hleft;        // Hp[k-1] still holds the previous row here only if written after use: see below
#endif
#if ABV_MAX3_ORDER == 1
    const uint32_t h = p16::max3(S[k], F[k], e);
#elif ABV_MAX3_ORDER == 2
    const uint32_t h = p16::max3(e, S[k], F[k]);
#elif ABV_MAX3_ORDER == 3
    const uint32_t h = p16::max3(e, F[k], S[k]);
#elif ABV_MAX3_ORDER == 4
    const uint32_t h = p16::max3(F[k], S[k], e);
#elif ABV_MAX3_ORDER == 5
    const uint32_t h = p16::max3(F[k], e, S[k]);
#else
    const uint32_t h = p16::max3(S[k], e, F[k]);
#endif
#if ABV_FE_ORDER == 1
    F[k] = p16::addmax(h, gop2, F[k]);
    e = p16::addmax(h, gope, e);
#else
    e = p16::addmax(h, gope, e);
    F[k] = p16::addmax(h, gop2, F[k]);
#endif
#if ABV_ROW_ADDS == 2
    Hn[k] = h;
#else
    Hp[k] = h;
#endif
  }
#if ABV_ROW_ADDS == 2
#pragma unroll
  for (int k = 0; k < K; ++k) Hp[k] = Hn[k];
#endif
}
template <int K, int KP>
ABV_HD void mm16d_row(uint32_t hleft, uint32_t &e, uint32_t (&Hp)[K], uint32_t (&F)[K], uint32_t (&S)[KP], uint32_t gop2) {
  mm16d_row<K, KP>(hleft, e, Hp, F, S, gop2, gop2);
}

// glocal row 0 through the common row (glocal_align.c:128-136: H[c][0] = S, F = H + go, no row gap).  The lane that
// starts a pair sets, in drift form, H~[c][-1] := g*(1 - c) (so that the diagonal alone gives S - g*c), F~ := -inf,
// takes E~ = -inf from the left and runs the row with the E-chain addend MM16G_NO_EGAP: h + (-32000) stays below the
// -inf sentinel for every representable h, so no row gap is opened and E leaves the strip as -inf, while F is
// opened normally (F~' = h + go - g).
constexpr int MM16G_NO_EGAP = -32000;

// ---------------------------------------------------------------------------------------------
// tb16: one-to-one alignments WITH traceback, two pairs per warp in the 16-bit halves (local and global;
// the merge stage and the final alignment, ambivert.py:301-331, :543-574).  Same strip geometry as the packed
// scorer (lane t owns columns [tK, (t+1)K), rows skewed by lane) and the same biased operands (namespace
// b16: every half holds true + 16384, so bit 15 of a half is always 0).  The decisions of a cell
// (global_align.c:180-209, local_align.c:150-207) are taken for both halves at once WITHOUT predicates:
//     ge(a, b) = a + 0x80008000 - b        bit 15 / bit 31 = (a >= b) in the low / high half
// (one 3-input add; no borrow crosses the halves because a + 0x8000 >= b), and combined with 3-input logic ops:
//   pa = ge(diag, F) (diag wins ties), pb = ge(max(diag,F), E) (both win ties over E), pc = ge(value, 0) (local: else stop),
//   fo = ge(h + go, F + ge), eo = ge(h + go, E + ge) (opening wins ties)
//   direction = SRC_* of wf32: bit0 = (!pa & pb) | !pc, bit1 = !(pb & pc) (stored complemented); plus the two "gap re-opened" bits
// The four direction bits of a cell ("planes" 0..3: bit0, complemented bit1, F re-opened, E re-opened) of both halves share ONE
// word per column and 4 steps.  Two byte permutes with sign replication turn the bits 15 / 31 of four decision words into whole
// bytes (X = planes 0,1; Y = planes 2,3; byte = 2*(plane & 1) + half), one logic op interleaves them (X on the odd bits, Y on the
// even ones, every bit pair of a byte alike) and one more keeps the pair of this step: 4 instructions per cell pair where one word
// per plane and 16 steps took a shift and a logic op per plane (8).
//     bit of (plane, half, step s) in word s / 4 of the column = 8 * (2*(plane & 1) + half) + 2 * (s % 4) + (plane < 2)
namespace p16 {
#if defined(__CUDA_ARCH__)
ABV_HD uint32_t bmax(uint32_t a, uint32_t b, bool &p_hi, bool &p_lo) { return __vibmax_s16x2(a, b, &p_hi, &p_lo); }
#else
ABV_HD uint32_t bmax(uint32_t a, uint32_t b, bool &p_hi, bool &p_lo) {
  p_lo = lo(a) >= lo(b); p_hi = hi(a) >= hi(b);
  return pack(p_lo ? lo(a) : lo(b), p_hi ? hi(a) : hi(b));
}
#endif
}  // namespace p16

template <int K>
struct Tb16Cfg {
  static constexpr int V = (K % 4 == 0) ? 4 : ((K % 2 == 0) ? 2 : 1);   // words per shared-memory load of the profile
  static constexpr int P = 32 * K;                                     // profile row length in words
  static constexpr int MAX_LA = 32 * K;
};
constexpr int TB16_MAX_ROWS = 1024;
// Local mode keeps lane 31 free of columns (queries up to 31*K): it never runs a row, so its outputs stay the constants of
// column -1 (H = 0, E = gap_open: local_align.c:123-128) and lane 0 reads them through the same indexed shuffle that hands every
// other lane its left neighbour's border -- no select after the shuffles.  Global's column -1 changes with the row: selects.
ABV_HD constexpr int tb16_max_la(bool local, int K) { return (local ? 31 : 32) * K; }
constexpr int TB16_PLANES = 4;                    // bit0, bit1 of the source, F re-opened, E re-opened
constexpr uint32_t TB16_H = 0x80008000u;
ABV_HD int tb16_steps16(int rows_cap) { return (rows_cap + 31 + 15) >> 4; }
// word of the profile row that holds column k of lane t: the lanes of one shared-memory phase read 32 distinct banks
ABV_HD int tb16_slot(int t, int k, int V) { return ((k / V) * 32 + t) * V + k % V; }
ABV_HD uint32_t tb16_ge(uint32_t a, uint32_t b) { return a + TB16_H - b; }
// acc | (x & m) as ONE 3-input logic op (left to itself the compiler may mask first)
ABV_HD uint32_t tb16_insert(uint32_t acc, uint32_t x, uint32_t m) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(x), "r"(m), "r"(acc));
  return d;
#else
  return acc | (x & m);
#endif
}
// (~pa & pb) | ~pc as one logic op on pa, pb, pc themselves (left to itself the compiler builds ~pc with a second add)
ABV_HD uint32_t tb16_b0_local(uint32_t pa, uint32_t pb, uint32_t pc) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0x5D;" : "=r"(d) : "r"(pa), "r"(pb), "r"(pc));
  return d;
#else
  return (~pa & pb) | ~pc;
#endif
}
// bytes (a.bit15, a.bit31, b.bit15, b.bit31), each replicated over its whole byte (prmt with the sign-replication flag)
ABV_HD uint32_t tb16_signs(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, 0xFDB9;" : "=r"(d) : "r"(a), "r"(b));
  return d;
#else
  return ((a >> 15) & 1u ? 0x000000ffu : 0u) | ((a >> 31) & 1u ? 0x0000ff00u : 0u) | ((b >> 15) & 1u ? 0x00ff0000u : 0u) | ((b >> 31) & 1u ? 0xff000000u : 0u);
#endif
}
// (x & 0xAAAAAAAA) | (y & 0x55555555) as one logic op
ABV_HD uint32_t tb16_interleave(uint32_t x, uint32_t y) {
#if defined(__CUDA_ARCH__)
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(d) : "r"(0xAAAAAAAAu), "r"(x), "r"(y));      // a ? b : c
  return d;
#else
  return (x & 0xAAAAAAAAu) | (y & 0x55555555u);
#endif
}
ABV_HD int tb16_steps4(int rows_cap) { return tb16_steps16(rows_cap) * 4; }

// One row of one lane.  All score-like values are b16-biased.
//   hleft : H[c0-1][r-1]        e : in E[r] entering the strip, out E[r] leaving it
//   Hp,F  : H[.][r-1] -> H[.][r]; column gap states
//   Slo,Shi : addend-packed substitution scores of the two pairs' columns for this row's symbols
//   gow,gew : b16::addend(go,go), b16::addend(ge,ge);  zerob = b16::biased(0,0)
//   mj    : 0x03030303 << 2*(step % 4): the bit pair of this step in every byte of the direction word acc[k]; first = step % 4 == 0
//   LOCAL : cand[k] = running best value of column c0+k (both halves), crl[k] / crh[k] = its row in the low / high half;
//           r = this row.  '>=' so that the last row wins ties within a column (local_align.c:156-160)
template <int MODE, int K>
ABV_HD void tb16_row(uint32_t hleft, uint32_t &e, uint32_t (&Hp)[K], uint32_t (&F)[K], const uint32_t *Slo, const uint32_t *Shi,
                     uint32_t gow, uint32_t gew, uint32_t zerob, uint32_t mj, bool first, uint32_t (&acc)[K],
                     uint32_t (&cand)[K], int (&crl)[K], int (&crh)[K], int r) {
  uint32_t d = hleft;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const uint32_t s0 = d + Slo[k] + Shi[k];
    d = Hp[k];
    const uint32_t t = p16::max2(s0, F[k]);
    const uint32_t pa = tb16_ge(s0, F[k]);                // s0 >= F: the diagonal wins ties
    const uint32_t hr = p16::max2(t, e);
    const uint32_t pb = tb16_ge(t, e);                    // max(diag, F) >= E
    uint32_t h = hr, b0, b1;
    if (MODE == LOCAL) {
      h = p16::max2(hr, zerob);
      const uint32_t pc = tb16_ge(hr, zerob);             // value >= 0, else the cell is a stop
      b0 = tb16_b0_local(pa, pb, pc);                     // (~pa & pb) | ~pc
      b1 = pb & pc;                                       // plane 1 holds the COMPLEMENT of source bit 1
      bool pn_h, pn_l;
      cand[k] = p16::bmax(hr, cand[k], pn_h, pn_l);       // a stop cell (hr < 0) never beats a candidate (>= -1)
      if (pn_l) crl[k] = r;                                // one predicated move per half (a row packed into the halves of
      if (pn_h) crh[k] = r;                                // one word took two logic ops per half: 8 % of the kernel's issue cycles)
    } else {
      b0 = ~pa & pb;
      b1 = pb;
    }
    const uint32_t hg = h + gow, fe = F[k] + gew, ee = e + gew;
    F[k] = p16::max2(hg, fe);
    const uint32_t fo = tb16_ge(hg, fe);                  // opening wins ties
    e = p16::max2(hg, ee);
    const uint32_t eo = tb16_ge(hg, ee);
    Hp[k] = h;
    // the first step of a direction word overwrites it (nothing to clear after the store of the previous one); a lane that
    // enters the matrix later in the group still holds the zero of the item start, one that has left it is never read there
    const uint32_t T = tb16_interleave(tb16_signs(b0, b1), tb16_signs(fo, eo));
    acc[k] = first ? (T & mj) : tb16_insert(acc[k], T, mj);
  }
}

// directions of one warp's two pairs: word(k, s, t) = (k*steps4 + s/4)*32 + t holds the four bits of both halves for the 4 steps
// s/4*4 .. +3 (bit layout above), where column c = t*K + k and s = row + t is the step at which lane t worked on that row
template <int K>
struct Tb16Reader {
  const uint32_t *w;
  int steps4, half;
  ABV_HD uint32_t cell(int col, int row) const {      // the word of the cell, shifted so that bits 0,1 = planes 2,0 and bits 16,17 = planes 3,1
    const int t = col / K, k = col - t * K, s = row + t;
    return w[((long long)k * steps4 + (s >> 2)) * 32 + t] >> (8 * half + 2 * (s & 3));
  }
  ABV_HD unsigned plane_bit(int plane, int col, int row) const { return (cell(col, row) >> (16 * (plane & 1) + (plane < 2))) & 1u; }
  ABV_HD unsigned src(int col, int row) const {
    const uint32_t v = cell(col, row);
    return ((v >> 1) & 1u) | ((((v >> 17) & 1u) ^ 1u) << 1);
  }
  ABV_HD unsigned get(int col, int row) const {       // the nibble of wf32 (for the generic walk)
    const uint32_t v = cell(col, row);
    return ((v >> 1) & 1u) | ((((v >> 17) & 1u) ^ 1u) << 1) | ((v & 1u) ? BIT_FOPEN : 0u) | (((v >> 16) & 1u) ? BIT_EOPEN : 0u);
  }
};

// ---------------------------------------------------------------------------------------------
// Variant scan of one aligned (sample row, reference row) pair: the step after the aligner
// (ambivert/call_variants.py:32-128, caller()).  The reference walks the
// two gapped rows site by site with three accumulating strings (mismatch, deletion, insertion) and
// yields (type, start in the gapped rows, one-based position in the ungapped reference, bases).  The
// strings only ever grow over consecutive sites, so a length and the site of the first base stand
// for each; a record is {kind, start, pos, len, run_start}: `start` and `pos` are the numbers the
// reference yields (call_variants.py:83-127, including its end-of-row arithmetic), the bases are
// row[run_start, run_start + len) of the sample row (X, I) or of the reference row (D).
//   softmask rules (call_variants.py:58-76): lower-case reference = primer; leading primer sites only
//   advance the position; a gap inside the leading primer is skipped, one that abuts unmasked sequence is
//   scanned; the first lower-case site after unmasked sequence ends the scan.
// Returns 0, or VAR_ERR_INDEX where the reference's look-ahead over a gap run leaves the row
// (IndexError, call_variants.py:66).  Python's str.lower()/upper() on ASCII.
enum { VAR_X = 0, VAR_D = 1, VAR_I = 2 };
enum { VAR_OK = 0, VAR_ERR_LENGTH = -1, VAR_ERR_INDEX = -2 };

ABV_HD bool var_is_own_lower(unsigned char c) { return !(c >= 'A' && c <= 'Z'); }        // c == c.lower()
ABV_HD unsigned char var_upper(unsigned char c) { return (c >= 'a' && c <= 'z') ? (unsigned char)(c - 32) : c; }

template <class Emit>
ABV_HD int variants_walk(const unsigned char *v, const unsigned char *r, int n, bool softmask, unsigned char gap, Emit emit) {
  int mm = 0, del = 0, ins = 0, mm_at = 0, del_at = 0, ins_at = 0, pos = 0, site = 0;
  bool masked = true;
  for (int i = 0; i < n; ++i) {
    site = i;
    const unsigned char rc = r[i], vc = v[i];
    const bool r_gap = rc == gap, r_low = var_is_own_lower(rc);
    if (softmask && masked && r_low && !r_gap) { ++pos; continue; }                  // :58-61 leading primer
    else if (softmask && masked && r_gap) {                                           // :62-69
      int g = 0;
      for (;;) {
        if (i + g >= n) return VAR_ERR_INDEX;
        if (r[i + g] != gap) break;
        ++g;
      }
      if (var_is_own_lower(r[i + g])) continue;                                       // gap within the primer
    } else if (softmask && !masked && !r_gap && r_low) break;                         // :70-72 trailing primer
    else masked = false;                                                              // :73-75

    if (var_upper(vc) == var_upper(rc) && !r_gap && vc != gap) {                      // :77-88 match
      ++pos;
      if (mm) { emit(VAR_X, i - mm, pos - mm, mm, mm_at); mm = 0; }
      else if (del) { emit(VAR_D, i - del, pos - del, del, del_at); del = 0; }
      else if (ins) { emit(VAR_I, i - ins, pos, ins, ins_at); ins = 0; }
    } else if (r_gap) {                                                               // :89-96 insertion grows
      if (!ins) ins_at = i;
      ++ins;
      if (mm) { emit(VAR_X, i - mm, pos - mm, mm, mm_at); mm = 0; }
      else if (del) { emit(VAR_D, i - del, pos - del + 1, del, del_at); del = 0; }
    } else if (vc == gap) {                                                           // :97-105 deletion grows
      if (!del) del_at = i;
      ++del; ++pos;
      if (mm) { emit(VAR_X, i - mm, pos - mm, mm, mm_at); mm = 0; }
      else if (ins) { emit(VAR_I, i - ins, pos, ins, ins_at); ins = 0; }
    } else {                                                                          // :106-117 substitution
      ++pos;
      if (mm) { emit(VAR_X, i - mm, pos - mm, mm, mm_at); }
      mm = 1; mm_at = i;
      if (ins) { emit(VAR_I, i - ins, pos, ins, ins_at); ins = 0; }
      else if (del) { emit(VAR_D, i - del, pos - del, del, del_at); del = 0; }
    }
  }
  ++pos;                                                                              // :119-128 states left at the end
  if (mm) emit(VAR_X, site - mm + 1, pos - mm, mm, mm_at);
  if (del) emit(VAR_D, site - del + 1, pos - del, del, del_at);
  if (ins) emit(VAR_I, site - ins + 1, pos, ins, ins_at);
  return VAR_OK;
}

// The BORDER LANE of the streaming kernel.  Column -1 of the DP matrix (what lane 0 receives "from the left") is a constant
// pair in drift form: (H~, E~) = (Hc, Ec) = (go + g, 2go) + level for global, (-inf, -inf) for glocal.  Selecting it into
// lane 0's shuffle results costs two SEL per step on the pipe that bounds the kernel (7.7 % of its ALU-pipe instructions at
// K = 8) and sits on the step's dependent chain.  Instead lane 31 owns no query columns: its profile is all zero, it receives
// its OWN outputs (shuffle by index: lane 0 reads lane 31, lane 31 reads itself, lane t reads lane t-1), and with
//   Hp[k] = hleft = Hc for all k,  e = Ec,  F[k] <= Hc      (Hc + (go - g) = Ec for global; all -inf for glocal)
// the common row maps this state onto itself: h_k = max3(Hc + 0, Ec, F) = Hc, e' = max(Hc + go - g, Ec) = Ec.  So lane 31
// emits (Hc, Ec) every step without a single extra instruction in the loops; its state is (re)written at the warp-uniform
// points where the constants change (pair start of an explicitly reset pair, around the separator row of an injected one).
// A strip width K then serves queries of up to 31K columns; the queries of (31K, 32K] columns, the widest strip and glocal take the
// instantiation that keeps the selects and all 32 lanes (template parameter BORDER of mm16g_kernel).
#ifndef ABV_NO_BORDER
#define ABV_NO_BORDER 0
#endif
#ifndef ABV_GLOCAL_BORDER
#define ABV_GLOCAL_BORDER 1
#endif
// Measured (B200, 20k x 2k panel, profiles/r02_ab_*.jsonl): global +2.7 % per executed cell; glocal +1 % per executed cell, less than
// the queries that would move up a strip width cost (its switching steps and parked rows dominate): glocal keeps the selects.
ABV_HD constexpr bool mm16g_border(int K, bool glocal) { return !ABV_NO_BORDER && K < 16 && (ABV_GLOCAL_BORDER || !glocal); }
constexpr int MM16G_BORDER_LANE = 31;

// pair-row bookkeeping of the streaming kernel, shared by the kernel and its CPU mirror
constexpr int MM16G_MIN_ROWS = 32;   // every pair occupies at least 32 steps: lane t switches pair t steps after lane 0
ABV_HD int mm16g_rows(int l1, int l2) {
  int m = l1 > l2 ? l1 : l2;
  return m < MM16G_MIN_ROWS ? MM16G_MIN_ROWS : m;
}

// Level plan of the streaming GLOBAL kernel: "reset by injection".
// A lane that starts a pair needs H~[c][-1] = Hc and F~[c][0] = Ec = Hc + (go - g) in all of its K columns
// (global_align.c:136-148 in drift form).  Writing 2K registers under a one-lane branch in each of the 32
// switching steps costs a third of such a step.  Instead, a pair that is FOLLOWED by an "injected" pair sweeps one
// extra row (the separator, a pad row at index `rows` of its buffer) on which lane 0 feeds E~ = C from the left,
// with C above everything the warp holds: the common row then leaves h = C in every column, E~ = C for the next
// lane, and F~ = max(F~, C + go - g) = C + go - g.  That IS the boundary state of a pair whose values are stored
// `level` higher, level' + go + g = C.  So pairs climb a staircase of levels and only every n-th pair is reset
// explicitly (back to the lowest level); the corner capture subtracts the pair's level.
//   W: |true drifted value| bound of a pair (the host's exactness bound, align_abi.cu packed_exact_streaming)
//   U: upper bound of every stored value above its level: d diagonal steps add at most d*(maxpos + 2|g|) and
//      gap steps add nothing in drift form; d <= min(P, rows) + 1, plus the row's own diagonal term
struct Mm16gLevels { int n, lvl0, step; };
ABV_HD Mm16gLevels mm16g_levels(int P, int max_rows, int go, int ge, int maxabs, int maxpos) {
  Mm16gLevels L{1, 0, 0};
  const int age = ge < 0 ? -ge : ge;
  const long long W = (long long)(P + max_rows + 12) * (maxabs + age);
  if (W > 16000 || go > ge) return L;                      // not the 16-bit streaming kernel, or go - g > 0: E~ would grow along the separator
  const int d = (P < max_rows + 1 ? P : max_rows + 1) + 3;
  const long long U = (long long)d * (maxpos + 2 * age);
  const long long step = U + (ge - go) + (-go - ge) + 4;   // C >= level + U + |go - g|,  level' = C - (go + g)
  const long long lvl0 = W - 16000;
  if (lvl0 + U > 16000) return L;
  const long long n = 1 + (16000 - U - lvl0) / step;
  if (n < 2) return L;
  L.n = (int)(n > 1024 ? 1024 : n); L.lvl0 = (int)lvl0; L.step = (int)step;
  return L;
}

}  // namespace abv
