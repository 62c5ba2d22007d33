// kernels.cuh -- hand-written sm_100a kernels of the AmBiVErT alignment backend.
//
//   encode_kernel        to_raw (lib/align/helpers.h:26-32): bytes -> symbol indices, on the device
//   wf32_kernel          general kernel: one warp per (sa, sb) pair, 32-bit, anti-diagonal wavefront
//                        over 32-column blocks with shuffle exchange; all four modes with the reference's
//                        tie rules; optional 4-bit direction matrix + traceback (bit-exact fragment lists)
//   tb16_kernel          one-to-one alignments WITH traceback, two pairs per warp in 16x2 DPX words (local, global):
//                        predicate-free tie decisions, byte-packed direction words (4 steps x 4 bits x 2 halves), warp-parallel walk back
//   variants_kernel      variant scan of aligned row pairs (call_variants.py caller()), one thread per pair
//   md5_pairs_kernel     MD5 keys of read pairs (the clustering front end), grouped by radix sort
//   assemble_*_kernel    gapped rows / merged reads built from the fragment pool (the string work that
//                        follows the aligner in merge_overlaps and align_to_reference)
//   mm16g_kernel         THE BULK KERNEL: many-to-many score-only, global and glocal.  One warp per
//                        (query, target PAIR) with both alignments packed in 16x2 DPX words; target pairs
//                        stream through the lane pipeline; drift-normalised recurrences (1 add + 3 DPX
//                        per cell pair); per-query substitution profile in shared memory; target rows
//                        staged by bulk async copies (TMA 1-D, mbarrier); fused per-query arg-max
//   mm16_kernel          first-generation packed kernel (one pair at a time per warp): local mode, and
//                        the fallback for value ranges beyond the biased 16-bit range
//   argmax_rows_kernel, best_key_finalize_kernel   first-maximum selection + seed rule (ambivert.py:478-485)
//   int_peak_kernel      register-only chains of the DP cell's instruction mix: the measured roofline denominator
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dp_core.cuh"

namespace abv {

#define ABV_FULL 0xffffffffu
#ifndef ABV_UNROLL
#define ABV_UNROLL 8
#endif
constexpr int kHotUnroll = ABV_UNROLL;   // steps of the steady-state loop unrolled together
#ifndef ABV_SWITCH_UNROLL
#define ABV_SWITCH_UNROLL 2
#endif
constexpr int kSwitchUnroll = ABV_SWITCH_UNROLL;
#ifndef ABV_GL_SW_UNROLL
#define ABV_GL_SW_UNROLL 1
#endif
constexpr int kGlSwitchUnroll = ABV_GL_SW_UNROLL;   // glocal (deferred last rows): steps of a switching segment unrolled together
#ifndef ABV_INJ_UNROLL
#define ABV_INJ_UNROLL 2   // round 1 build: 1: 8.065, 2: 8.040, 4: 8.084, 8: 8.088 TCUPS; round 2 build (20k x 2k): 1: 8.706, 2: 8.748, 4: 8.725, 8: 8.680
#endif
constexpr int kInjUnroll = ABV_INJ_UNROLL;         // the same for the (lighter) switching segments of injected pairs   // steps of the pair-switching segments unrolled together (code size: instruction cache)
#ifndef ABV_LDS_PTX
#define ABV_LDS_PTX 1      // streaming kernel: profile vectors by ld.shared with one multiply-add of address arithmetic per step
#endif
#ifndef ABV_DYN_PAIRS
#define ABV_DYN_PAIRS 1   // streaming kernel: the warps of a CTA take the target pairs of a work item from a shared counter (1) or strided (0)
#endif
constexpr bool kDynPairs = ABV_DYN_PAIRS != 0;
#ifndef ABV_MINB8
#define ABV_MINB8 3   // measured on B200: 3 CTAs (24 warps, <= 85 registers) beat 4 (32 warps, 64 registers) by 4 %; 2 and 5 are slower
#endif

// ---------------------------------------------------------------------------------------------
__global__ void encode_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, long long n,
                              const uint8_t *__restrict__ map256) {
  __shared__ uint8_t m[256];
  if (threadIdx.x < 256) m[threadIdx.x] = map256[threadIdx.x];
  __syncthreads();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = m[in[i]];
}

// ---------------------------------------------------------------------------------------------
struct Wf32Params {
  const uint8_t *a; const long long *a_off;      // encoded sa sequences (queries)
  const uint8_t *b; const long long *b_off;      // encoded sb sequences (targets)
  long long n_pairs;                             // pairs of this launch
  int n_t;                                       // 0: pair p = (a[p], b[p]); >0: p -> (row q_base + p / n_t, p % n_t)
  long long q_base;
  const int *q_list;                             // optional: query of a row = q_list[row]
  int score_rows_by_query;                       // many-to-many: scores[query*n_t + target] instead of scores[p]
  const int *matrix; int alpha, go, ge;
  int *scores;                                   // [n_pairs]
  // traceback (TRACE only)
  uint32_t *dir; long long dir_words_per_warp;   // per-warp direction scratch
  FragOut *frag_tmp; int frag_tmp_cap;           // per-warp reversed fragment scratch
  FragOut *frags; long long frag_cap;            // output pool
  unsigned long long *frag_total;
  long long *frag_start; int *frag_count;
  // column-block border hand-over (queries longer than 32): 2*border_len ints per warp
  int *border; int border_len;
  int stage_len;                                 // STAGED kernels: per-warp shared-memory rows of this many entries (multiple of 16)
  unsigned long long *pair_counter;
  const int *pair_list;                          // one-to-one jobs: work item i is pair pair_list[i] (NULL: pair i)
};

constexpr int WF32_WARPS = 8;
constexpr int WF32_MAX_SMEM_ALPHA = 48;

// The fill of one pair by one warp: 32-column blocks, lane = column, rows skewed by lane; the block's
// right border (H, E per row) goes through borderH/borderE to the next block.  Warp-uniform results.
template <int MODE, bool TRACE, bool STAGED>
__device__ __forceinline__ void wf32_fill(const uint8_t *sa, int la, const uint8_t *sb, int lb, const int *M, int alpha,
                                          int go, int ge, uint32_t *dir, int *borderH, int *borderE, uint8_t *ssb, int lane,
                                          int &score, int &ec, int &er) {
  const int nblk = (la + 31) >> 5;
  const int steps8 = dir_steps8(lb);
  if (STAGED) {
    __syncwarp();
    for (int i = lane; i < lb; i += 32) ssb[i] = sb[i];
    __syncwarp();
  }
  EndCell wbest = wf32_initial_candidate<MODE>();
  int corner = 0;
  for (int b = 0; b < nblk; ++b) {
    const int c = (b << 5) + lane;
    const bool colvalid = c < la;
    const int qa = colvalid ? (int)sa[c] * alpha : 0;
    int F = 0, hleft_prev = 0, h_out = 0, e_out = 0;
    EndCell cand = wf32_initial_candidate<MODE>();
    uint32_t acc = 0;
    const int lanes_used = min(32, la - (b << 5));
    const int nsteps = lb + lanes_used - 1;
    for (int s = 0; s < nsteps; ++s) {
      int h_sh = __shfl_up_sync(ABV_FULL, h_out, 1);
      int e_sh = __shfl_up_sync(ABV_FULL, e_out, 1);
      const int r = s - lane;
      if (lane == 0 && s < lb) {
        if (b == 0) wf32_virtual_column<MODE>(s, go, ge, h_sh, e_sh);
        else { h_sh = borderH[s]; e_sh = borderE[s]; }
      }
      unsigned nib = 0;
      if (r >= 0 && r < lb && colvalid) {
        const int sub = M[qa + (int)(STAGED ? ssb[r] : sb[r])];
        int h, eo;
        nib = wf32_cell<MODE>(c, r, la, lb, sub, hleft_prev, e_sh, F, go, ge, h, eo, cand);
        hleft_prev = h_sh;
        h_out = h;
        e_out = eo;
        if (MODE == GLOBAL && c == la - 1 && r == lb - 1) corner = h;
        if (lane == 31 && b + 1 < nblk) { borderH[r] = h; borderE[r] = eo; }
      }
      if (TRACE) {
        acc |= nib << ((s & 7) * 4);
        if ((s & 7) == 7 || s == nsteps - 1) {
          dir[((long long)b * steps8 + (s >> 3)) * 32 + lane] = acc;
          acc = 0;
        }
      }
    }
    // best end cell of this block, then fold into the running best (later blocks win ties)
    for (int off = 16; off > 0; off >>= 1) {
      EndCell o;
      o.score = __shfl_xor_sync(ABV_FULL, cand.score, off);
      o.col = __shfl_xor_sync(ABV_FULL, cand.col, off);
      o.row = __shfl_xor_sync(ABV_FULL, cand.row, off);
      if (end_better(o, cand)) cand = o;
    }
    if (cand.col >= 0 && cand.score >= wbest.score) wbest = cand;
    __syncwarp();
  }
  if (MODE == GLOBAL) corner = __shfl_sync(ABV_FULL, corner, (la - 1) & 31);
  wf32_finish<MODE>(wbest, corner, la, lb, score, ec, er);
}

// STAGED: the target symbols and the column-block border arrays of each warp live in dynamic shared
// memory (stage_len bytes + 2*stage_len ints per warp) instead of global memory; the host picks it when
// the longest target of the batch fits (Wf32Params::stage_len > 0).
template <int MODE, bool TRACE, bool STAGED>
__global__ void __launch_bounds__(WF32_WARPS * 32) wf32_kernel(Wf32Params p) {
  extern __shared__ __align__(16) unsigned char wf32_smem[];
  __shared__ int s_matrix[WF32_MAX_SMEM_ALPHA * WF32_MAX_SMEM_ALPHA];
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const long long gwarp = (long long)blockIdx.x * WF32_WARPS + warp_in_block;
  const int alpha = p.alpha, go = p.go, ge = p.ge;
  const bool smem_matrix = alpha <= WF32_MAX_SMEM_ALPHA;
  if (smem_matrix)
    for (int i = threadIdx.x; i < alpha * alpha; i += blockDim.x) s_matrix[i] = p.matrix[i];
  __syncthreads();
  const int *M = (STAGED || smem_matrix) ? s_matrix : p.matrix;   // STAGED launches always have alpha <= WF32_MAX_SMEM_ALPHA

  uint32_t *dir = TRACE ? p.dir + gwarp * p.dir_words_per_warp : nullptr;
  int *borderH, *borderE;
  uint8_t *ssb = nullptr;
  if (STAGED) {
    unsigned char *mine = wf32_smem + (size_t)warp_in_block * (size_t)p.stage_len * 9;
    borderH = reinterpret_cast<int *>(mine);
    borderE = borderH + p.stage_len;
    ssb = mine + (size_t)p.stage_len * 8;
  } else {
    borderH = p.border + gwarp * 2 * (long long)p.border_len;
    borderE = borderH + p.border_len;
  }

  for (;;) {
    unsigned long long pidx = 0;
    if (lane == 0) pidx = atomicAdd(p.pair_counter, 1ull);
    pidx = __shfl_sync(ABV_FULL, pidx, 0);
    if ((long long)pidx >= p.n_pairs) break;
    long long ia, ib, out_idx = (long long)pidx;
    if (p.n_t > 0) {
      const long long row = p.q_base + (long long)(pidx / (unsigned)p.n_t);
      ia = p.q_list ? (long long)p.q_list[row] : row;
      ib = (long long)(pidx % (unsigned)p.n_t);
      if (p.score_rows_by_query) out_idx = ia * p.n_t + ib;
    } else {
      if (p.pair_list) pidx = (unsigned long long)p.pair_list[pidx];
      ia = (long long)pidx; ib = (long long)pidx; out_idx = (long long)pidx;
    }
    const uint8_t *sa = p.a + p.a_off[ia];
    const uint8_t *sb = p.b + p.b_off[ib];
    const int la = (int)(p.a_off[ia + 1] - p.a_off[ia]);
    const int lb = (int)(p.b_off[ib + 1] - p.b_off[ib]);
    const int steps8 = dir_steps8(lb);
    int score, ec, er;
    wf32_fill<MODE, TRACE, STAGED>(sa, la, sb, lb, M, alpha, go, ge, dir, borderH, borderE, ssb, lane, score, ec, er);
    if (lane == 0) p.scores[out_idx] = score;

    if (TRACE) {
      __syncwarp();
      FragOut *tmp = p.frag_tmp + gwarp * (long long)p.frag_tmp_cap;
      int n = 0;
      unsigned long long base = 0;
      if (lane == 0) {
        DirReader rd; rd.w = dir; rd.steps8 = steps8;
        n = walk_back<MODE>(rd, ec, er, tmp, p.frag_tmp_cap);
        base = atomicAdd(p.frag_total, (unsigned long long)n);
        p.frag_start[pidx] = (long long)base;
        p.frag_count[pidx] = n;
      }
      n = __shfl_sync(ABV_FULL, n, 0);
      base = __shfl_sync(ABV_FULL, base, 0);
      __syncwarp();
      for (int i = lane; i < n; i += 32)
        if ((long long)(base + i) < p.frag_cap) p.frags[base + i] = tmp[n - 1 - i];
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// tb16_kernel<MODE,K>: one-to-one alignments with traceback, TWO pairs per warp (16-bit halves), local and
// global.  Lane t owns query columns [tK, (t+1)K) of both pairs and sweeps the target rows one row behind its
// left neighbour; strip borders (H, E) move with two shuffles per step.  Per warp in shared memory: the two
// pairs' substitution profiles (prof[half][symbol][column], addend-packed so that hdiag + lo + hi is two plain
// adds), and per target row the offsets of the two halves' profile rows (8 bits each, in units of 128 bytes).  Sequences shorter than their partner's
// are padded with an extra symbol whose score is negative against everything: with gap_open < 0 a padded cell
// is always strictly below the best real cell it derives from, so it never becomes the local end cell, and
// real cells never depend on padded ones.  Directions: one word per column and 4 steps with the four direction bits of
// both halves (dp_core.cuh::Tb16Reader), written as whole 128-byte lines every 4 steps.  The walk back is done by the whole warp, one pair after
// the other: the lanes probe 32 cells of the current diagonal (or 32 cells up / left for a gap) at once and a
// ballot finds where the run ends, so a fragment costs one or two memory round trips instead of one per cell.
struct Tb16Params {
  const uint8_t *a; const long long *a_off;      // encoded sequences (sa = query side, columns)
  const uint8_t *b; const long long *b_off;      // (sb = target side, rows)
  const int *pair_list;                          // item i = pairs pair_list[2i], pair_list[2i+1] (-1: none)
  int n_items;
  const int *matrix_ext;                         // (alpha+1)^2, last row/column = the padding symbol
  int a1, go, ge;
  int rows_cap;                                  // staged target rows per warp (multiple of 16, <= TB16_MAX_ROWS)
  int *scores;
  uint32_t *dir; long long dir_words_per_warp;   // K * steps4 * 32 words (steps4 = 4 * steps16)
  FragOut *frag_tmp; int frag_tmp_cap;           // frag_tmp_cap entries per warp
  FragOut *frags; long long frag_cap;
  unsigned long long *frag_total;
  long long *frag_start; int *frag_count;
  unsigned long long *item_counter;
};
constexpr int TB16_WARPS = 4;
#ifndef TB16_MINB
#define TB16_MINB 6
#endif
// resident CTAs the register allocation aims at: the shared memory of the profiles allows 6 CTAs of 4 warps at K <= 5 (85 registers)
#define TB16_MIN_BLOCKS(K) ((K) <= 5 ? TB16_MINB : ((K) <= 8 ? 4 : 3))
template <int K>
__host__ __device__ inline int tb16_smem_per_warp(int a1, int rows_cap) {
  // profiles + one area that holds the query codes while the profiles are built and the target rows afterwards
  const int rows_bytes = rows_cap * 2, qc_bytes = 2 * Tb16Cfg<K>::P;
  return (2 * a1 * Tb16Cfg<K>::P * 4 + (rows_bytes > qc_bytes ? rows_bytes : qc_bytes) + 15) & ~15;
}

template <int K>
__device__ __forceinline__ uint32_t tb16_pick(const uint32_t (&Hp)[K], int kstar) {
  uint32_t v = Hp[0];
#pragma unroll
  for (int k = 1; k < K; ++k) v = (k == kstar) ? Hp[k] : v;
  return v;
}

// Traceback of one half by the whole warp; same fragments, same order as dp_core.cuh::walk_back_emit
// (global_align.c:25-92, local_align.c:23-72).  tmp receives them in REVERSE sequence order; returns the count.
template <int MODE, int K>
__device__ __forceinline__ int tb16_walk(const Tb16Reader<K> &rd, int col, int row, FragOut *tmp, int cap, int lane) {
  int n = 0;
  while (col >= 0 && row >= 0) {
    const unsigned src0 = rd.src(col, row);                     // warp-uniform address: one broadcast load for all four bits
    if (MODE == LOCAL && src0 == SRC_STOP) break;
    FragOut f;
    if (src0 == SRC_F) {              // back to the nearest row above where the column gap was (re)opened
      int p = row - 1;
      for (;;) {
        const int pi = p - lane;
        const unsigned hit = __ballot_sync(ABV_FULL, pi >= 0 && rd.plane_bit(2, col, pi));
        if (hit) { p -= __ffs(hit) - 1; break; }
        p -= 32;
        if (p < 0) { p = -1; break; }
      }
      f.type = FRAG_A_GAP; f.hsp_len = row - p;
      row = p;
    } else if (src0 == SRC_E) {       // back to the nearest column to the left where the row gap was (re)opened
      int p = col - 1;
      for (;;) {
        const int pi = p - lane;
        const unsigned hit = __ballot_sync(ABV_FULL, pi >= 0 && rd.plane_bit(3, pi, row));
        if (hit) { p -= __ffs(hit) - 1; break; }
        p -= 32;
        if (p < 0) { p = -1; break; }
      }
      f.type = FRAG_B_GAP; f.hsp_len = col - p;
      col = p;
    } else {                          // diagonal run: the start cell, then every following cell whose source is the diagonal
      int run = 1;
      --col; --row;
      for (;;) {
        const int ci = col - lane, ri = row - lane;
        const unsigned diag = __ballot_sync(ABV_FULL, ci >= 0 && ri >= 0 && rd.src(ci, ri) == SRC_DIAG);
        const int cnt = diag == ABV_FULL ? 32 : __ffs(~diag) - 1;
        run += cnt; col -= cnt; row -= cnt;
        if (cnt < 32) break;
      }
      f.type = FRAG_MATCH; f.hsp_len = run;
    }
    f.sa_start = col + 1; f.sb_start = row + 1;
    if (lane == 0 && n < cap) tmp[n] = f;
    ++n;
  }
  if (MODE == GLOBAL && (col >= 0 || row >= 0)) {   // leading end gap (global_align.c:69-86)
    FragOut f; f.sa_start = 0; f.sb_start = 0;
    if (col >= 0) { f.type = FRAG_B_GAP; f.hsp_len = col + 1; }
    else          { f.type = FRAG_A_GAP; f.hsp_len = row + 1; }
    if (lane == 0 && n < cap) tmp[n] = f;
    ++n;
  }
  return n;
}

template <int MODE, int K>
__global__ void __launch_bounds__(TB16_WARPS * 32, TB16_MIN_BLOCKS(K)) tb16_kernel(Tb16Params p) {
  static_assert(MODE == LOCAL || MODE == GLOBAL, "tb16 implements local and global");
  static_assert(15 * K < 256, "profile-row offsets of a target row are 8 bits per half (alphabet + padding symbol <= 8 rows per half)");
  constexpr int V = Tb16Cfg<K>::V, P = Tb16Cfg<K>::P;
  extern __shared__ __align__(16) unsigned char tb_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long gwarp = (long long)blockIdx.x * TB16_WARPS + warp;
  const int a1 = p.a1, go = p.go, ge = p.ge, pad = a1 - 1;
  int *mext = reinterpret_cast<int *>(tb_smem);
  unsigned char *mine = tb_smem + ((a1 * a1 * 4 + 15) & ~15) + (size_t)warp * tb16_smem_per_warp<K>(a1, p.rows_cap);
  uint32_t *prof = reinterpret_cast<uint32_t *>(mine);                  // [half][symbol][P]
  // [rows_cap] per target row the offsets of the two halves' profile rows in units of 128 bytes (a row is K units; at most
  // 2*8 rows of 16 units): lo half | hi half << 8.  The same bytes hold the query codes [half][P] while the profiles are built.
  uint16_t *tb = reinterpret_cast<uint16_t *>(mine + 2 * a1 * P * 4);
  uint8_t *qc = mine + 2 * a1 * P * 4;
  for (int i = threadIdx.x; i < a1 * a1; i += blockDim.x) mext[i] = p.matrix_ext[i];
  __syncthreads();

  const uint32_t gow = b16::addend(go, go), gew = b16::addend(ge, ge), zerob = b16::biased(0, 0);
  const int steps4 = tb16_steps4(p.rows_cap);
  uint32_t *dir = p.dir + gwarp * p.dir_words_per_warp;
  FragOut *tmp = p.frag_tmp + gwarp * (long long)p.frag_tmp_cap;
  const int c0 = lane * K;

  for (;;) {
    unsigned long long item = 0;
    if (lane == 0) item = atomicAdd(p.item_counter, 1ull);
    item = __shfl_sync(ABV_FULL, item, 0);
    if ((long long)item >= p.n_items) break;
    const int i0 = p.pair_list[2 * item], i1 = p.pair_list[2 * item + 1];
    const int j1 = i1 >= 0 ? i1 : i0;                           // no partner: the pair rides in both halves
    const uint8_t *qa0 = p.a + p.a_off[i0], *qa1 = p.a + p.a_off[j1];
    const uint8_t *tb0 = p.b + p.b_off[i0], *tb1 = p.b + p.b_off[j1];
    const int la0 = (int)(p.a_off[i0 + 1] - p.a_off[i0]), la1 = (int)(p.a_off[j1 + 1] - p.a_off[j1]);
    const int lb0 = (int)(p.b_off[i0 + 1] - p.b_off[i0]), lb1 = (int)(p.b_off[j1 + 1] - p.b_off[j1]);
    const int rows = max(lb0, lb1);
    const int tl = (max(la0, la1) - 1) / K;                     // last lane that owns a real column
    const int nsteps = rows + tl;

    __syncwarp();
    for (int c = lane; c < 32 * K; c += 32) {
      qc[c] = c < la0 ? qa0[c] : (uint8_t)pad;
      qc[P + c] = c < la1 ? qa1[c] : (uint8_t)pad;
    }
    __syncwarp();
    for (int idx = lane; idx < a1 * P; idx += 32) {
      const int sym = idx / P, pos = idx - sym * P;
      const int j = pos % V, it = pos / V, t = it & 31, i = it >> 5;
      const int k = i * V + j;
      prof[idx] = b16::addend(mext[qc[t * K + k] * a1 + sym], 0);
      prof[a1 * P + idx] = b16::addend(0, mext[qc[P + t * K + k] * a1 + sym]);
    }
    __syncwarp();                                 // the query codes are dead: their bytes take the target rows
    for (int r = lane; r < rows; r += 32)
      tb[r] = (uint16_t)(((r < lb0 ? tb0[r] : pad) * K) | (((a1 + (r < lb1 ? tb1[r] : pad)) * K) << 8));
    __syncwarp();

    uint32_t Hp[K], F[K], acc[K], cand[K];
    int crl[K], crh[K];
    uint32_t hleft, h_out, e_out;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const int c = c0 + k;
      if (MODE == GLOBAL) { Hp[k] = b16::biased(go + ge * c, go + ge * c); F[k] = b16::biased(2 * go + ge * c, 2 * go + ge * c); }
      else { Hp[k] = zerob; F[k] = b16::biased(go, go); }
      acc[k] = 0;
      cand[k] = b16::biased(-1, -1); crl[k] = 0; crh[k] = 0;
    }
    if (MODE == GLOBAL) hleft = c0 ? b16::biased(go + ge * (c0 - 1), go + ge * (c0 - 1)) : zerob;
    else hleft = zerob;
    // column -1 as seen by lane 0 at row r: H = go + ge*r, E = 2go + ge*r (global_align.c:136-140); 0, go (local_align.c:123-128)
    const uint32_t Hv0 = MODE == GLOBAL ? b16::biased(go, go) : zerob;
    const uint32_t Ev0 = MODE == GLOBAL ? b16::biased(2 * go, 2 * go) : b16::biased(go, go);
    // local: lane 31 owns no column (tb16_max_la) and keeps the constants of column -1 for lane 0's indexed shuffle
    h_out = zerob; e_out = (MODE == LOCAL && lane == 31) ? Ev0 : zerob;
    const int t0 = (la0 - 1) / K, k0 = (la0 - 1) - t0 * K, t1 = (la1 - 1) / K, k1 = (la1 - 1) - t1 * K;
    int corner0 = 0, corner1 = 0;

    for (int s4 = 0; s4 < nsteps; s4 += 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {              // fully unrolled: a direction word holds 4 steps, the first one overwrites it
        const int s = s4 + j;
        uint32_t h_sh, e_sh;
        const int r = s - lane;
        if (MODE == LOCAL) {
          h_sh = __shfl_sync(ABV_FULL, h_out, (lane + 31) & 31);
          e_sh = __shfl_sync(ABV_FULL, e_out, (lane + 31) & 31);
        } else {
          h_sh = __shfl_up_sync(ABV_FULL, h_out, 1);
          e_sh = __shfl_up_sync(ABV_FULL, e_out, 1);
          if (lane == 0) { h_sh = Hv0 + (uint32_t)r * gew; e_sh = Ev0 + (uint32_t)r * gew; }
        }
        if (r >= 0 && r < rows && lane <= tl) {
          const uint32_t code = tb[r];
          uint32_t Slo[K], Shi[K];
          const uint32_t *rlo = prof + (code & 0xffu) * 32u, *rhi = prof + (code >> 8) * 32u;
          if (V == 4) {
#pragma unroll
            for (int i = 0; i < K / 4; ++i) {
              const uint4 u = reinterpret_cast<const uint4 *>(rlo)[i * 32 + lane], w = reinterpret_cast<const uint4 *>(rhi)[i * 32 + lane];
              Slo[4 * i] = u.x; Slo[4 * i + 1] = u.y; Slo[4 * i + 2] = u.z; Slo[4 * i + 3] = u.w;
              Shi[4 * i] = w.x; Shi[4 * i + 1] = w.y; Shi[4 * i + 2] = w.z; Shi[4 * i + 3] = w.w;
            }
          } else if (V == 2) {
#pragma unroll
            for (int i = 0; i < K / 2; ++i) {
              const uint2 u = reinterpret_cast<const uint2 *>(rlo)[i * 32 + lane], w = reinterpret_cast<const uint2 *>(rhi)[i * 32 + lane];
              Slo[2 * i] = u.x; Slo[2 * i + 1] = u.y; Shi[2 * i] = w.x; Shi[2 * i + 1] = w.y;
            }
          } else {
#pragma unroll
            for (int i = 0; i < K; ++i) { Slo[i] = rlo[i * 32 + lane]; Shi[i] = rhi[i * 32 + lane]; }
          }
          uint32_t e = e_sh;
          tb16_row<MODE, K>(hleft, e, Hp, F, Slo, Shi, gow, gew, zerob, 0x03030303u << (2 * j), j == 0, acc, cand, crl, crh, r);
          hleft = h_sh; e_out = e; h_out = Hp[K - 1];
          if (MODE == GLOBAL) {                                   // the corner cell is the score (global_align.c:217)
            if (r == lb0 - 1 && lane == t0) corner0 = b16::lo(tb16_pick<K>(Hp, k0));
            if (r == lb1 - 1 && lane == t1) corner1 = b16::hi(tb16_pick<K>(Hp, k1));
          }
        }
      }
      const long long wbase = (long long)(s4 >> 2) * 32 + lane;
#pragma unroll
      for (int k = 0; k < K; ++k) dir[(long long)k * steps4 * 32 + wbase] = acc[k];
    }

    // end cell and score of each half
    int score0, ec0, er0, score1, ec1, er1;
    if (MODE == GLOBAL) {
      score0 = __shfl_sync(ABV_FULL, corner0, t0); score1 = __shfl_sync(ABV_FULL, corner1, t1);
      ec0 = la0 - 1; er0 = lb0 - 1; ec1 = la1 - 1; er1 = lb1 - 1;
    } else {
      // per lane: best over its columns, later columns win ties; then over lanes by (score, column)
      int b0 = -1, bc0 = 0, br0 = 0, b1 = -1, bc1 = 0, br1 = 0;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int v0 = b16::lo(cand[k]), v1 = b16::hi(cand[k]);
        if (v0 >= 0 && v0 >= b0) { b0 = v0; bc0 = c0 + k; br0 = crl[k]; }
        if (v1 >= 0 && v1 >= b1) { b1 = v1; bc1 = c0 + k; br1 = crh[k]; }
      }
      long long key0 = b0 < 0 ? -1 : (((long long)b0 << 40) | ((long long)bc0 << 20) | br0);
      long long key1 = b1 < 0 ? -1 : (((long long)b1 << 40) | ((long long)bc1 << 20) | br1);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        key0 = max(key0, __shfl_xor_sync(ABV_FULL, key0, off));
        key1 = max(key1, __shfl_xor_sync(ABV_FULL, key1, off));
      }
      // max_score starts at 0 with end cell (0,0) (local_align.c:119-121)
      if (key0 < 0) { score0 = 0; ec0 = 0; er0 = 0; } else { score0 = (int)(key0 >> 40); ec0 = (int)((key0 >> 20) & 0xfffff); er0 = (int)(key0 & 0xfffff); }
      if (key1 < 0) { score1 = 0; ec1 = 0; er1 = 0; } else { score1 = (int)(key1 >> 40); ec1 = (int)((key1 >> 20) & 0xfffff); er1 = (int)(key1 & 0xfffff); }
    }
    if (lane == 0) { p.scores[i0] = score0; if (i1 >= 0) p.scores[i1] = score1; }
    __syncwarp();      // the direction words of all lanes are visible to the whole warp

#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && i1 < 0) break;
      Tb16Reader<K> rd; rd.w = dir; rd.steps4 = steps4; rd.half = h;
      const int n = tb16_walk<MODE, K>(rd, h ? ec1 : ec0, h ? er1 : er0, tmp, p.frag_tmp_cap, lane);
      unsigned long long base = 0;
      if (lane == 0) {
        base = atomicAdd(p.frag_total, (unsigned long long)n);
        const int pi = h ? i1 : i0;
        p.frag_start[pi] = (long long)base;
        p.frag_count[pi] = n;
      }
      base = __shfl_sync(ABV_FULL, base, 0);
      __syncwarp();
      for (int i = lane; i < n; i += 32)
        if ((long long)(base + i) < p.frag_cap) p.frags[base + i] = tmp[n - 1 - i];
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// One pair, one launch: what a call through the reference's own ABI (global_align, local_align, ...) costs.
// The call's whole input arrives in ONE host->device copy (`in`: 256-byte map, alpha*alpha ints, sa bytes, sb
// bytes) and its whole result leaves in ONE device->host copy (`out`: score, fragment count, fragments in
// sequence order).  One warp encodes, fills, walks back.  Host threads call concurrently on their own streams.
struct Wf32OneParams {
  const unsigned char *in;       // [map 256][matrix alpha*alpha ints][sa la bytes][sb lb bytes]
  int la, lb, alpha, go, ge, use_map;
  uint8_t *codes;                // la + lb bytes of scratch
  uint32_t *dir;                 // dir_words(la, lb)
  int *border;                   // 2 * lb ints per 32-column block
  FragOut *tmp; int tmp_cap;     // la + lb + 2
  int *out;                      // [score, n_frags, pad, pad][frags ...]
  int out_frag_cap;
};

constexpr int WF32_ONE_WARPS = 8;
constexpr int WF32_ONE_CHUNK = 16;      // border rows handed from one column block to the next at a time

// The column blocks of the one pair are pipelined over the warps of the CTA: warp w fills blocks w, w+8, ...
// and block b may run as soon as block b-1 has published the next chunk of its right border (a progress
// counter in shared memory; forward dependencies only, all warps resident, so the spin cannot deadlock).
// SMEM: the encoded sequences and the block borders live in dynamic shared memory
// ([border: nblk*2*lb ints][sa: la bytes][sb: lb bytes]); otherwise in the global scratch of the call.
template <int MODE, bool SMEM>
__global__ void __launch_bounds__(WF32_ONE_WARPS * 32) wf32_one_kernel(Wf32OneParams p) {
  extern __shared__ __align__(16) unsigned char one_smem[];
  __shared__ int s_matrix[WF32_MAX_SMEM_ALPHA * WF32_MAX_SMEM_ALPHA];
  __shared__ uint8_t s_map[256];
  __shared__ volatile int s_progress[WF32_ONE_WARPS];     // rows of the right border published by the block a warp is on
  __shared__ volatile int s_block_of[WF32_ONE_WARPS];     // which block that is
  __shared__ int s_best[WF32_ONE_WARPS * 3];
  __shared__ int s_corner;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int alpha = p.alpha, la = p.la, lb = p.lb, go = p.go, ge = p.ge;
  const int *matrix = reinterpret_cast<const int *>(p.in + 256);
  const unsigned char *ra = p.in + 256 + (size_t)alpha * alpha * 4, *rb = ra + la;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_map[i] = p.in[i];
  for (int i = threadIdx.x; i < alpha * alpha; i += blockDim.x) s_matrix[i] = matrix[i];
  if (threadIdx.x < WF32_ONE_WARPS) { s_progress[threadIdx.x] = 0; s_block_of[threadIdx.x] = -1; }
  __syncthreads();
  const int nblk = (la + 31) >> 5;
  int *border = SMEM ? reinterpret_cast<int *>(one_smem) : p.border;
  uint8_t *sa = SMEM ? one_smem + (size_t)nblk * 2 * lb * 4 : p.codes;
  uint8_t *sb = sa + la;
  for (int i = threadIdx.x; i < la; i += blockDim.x) sa[i] = (uint8_t)min((int)(p.use_map ? s_map[ra[i]] : ra[i]), alpha - 1);
  for (int i = threadIdx.x; i < lb; i += blockDim.x) sb[i] = (uint8_t)min((int)(p.use_map ? s_map[rb[i]] : rb[i]), alpha - 1);
  __syncthreads();

  const int steps8 = dir_steps8(lb);
  EndCell wbest = wf32_initial_candidate<MODE>();
  int corner = 0;
  for (int b = warp; b < nblk; b += WF32_ONE_WARPS) {
    // borders: block b reads rows of border b-1 and writes border b (2*lb ints each)
    int *outH = border + (size_t)b * 2 * lb, *outE = outH + lb;
    const int *inH = border + (size_t)(b - 1) * 2 * lb, *inE = inH + lb;
    const int prod = (b - 1) & (WF32_ONE_WARPS - 1);       // the warp that fills block b-1
    if (lane == 0) { s_progress[warp] = 0; __threadfence_block(); s_block_of[warp] = b; }
    const int c = (b << 5) + lane;
    const bool colvalid = c < la;
    const int qa = colvalid ? (int)sa[c] * alpha : 0;
    int F = 0, hleft_prev = 0, h_out = 0, e_out = 0;
    EndCell cand = wf32_initial_candidate<MODE>();
    uint32_t acc = 0;
    const int lanes_used = min(32, la - (b << 5));
    const int nsteps = lb + lanes_used - 1;
    for (int s = 0; s < nsteps; ++s) {
      if (b > 0 && s < lb && (s % WF32_ONE_CHUNK) == 0) {   // wait for the next chunk of the left block's border
        const int need = min(lb, s + WF32_ONE_CHUNK);
        if (lane == 0) while (s_block_of[prod] < b - 1 || (s_block_of[prod] == b - 1 && s_progress[prod] < need)) { }
        __syncwarp();
        __threadfence_block();
      }
      int h_sh = __shfl_up_sync(ABV_FULL, h_out, 1);
      int e_sh = __shfl_up_sync(ABV_FULL, e_out, 1);
      const int r = s - lane;
      if (lane == 0 && s < lb) {
        if (b == 0) wf32_virtual_column<MODE>(s, go, ge, h_sh, e_sh);
        else if (SMEM) { h_sh = inH[s]; e_sh = inE[s]; }
        else { h_sh = __ldcg(inH + s); e_sh = __ldcg(inE + s); }
      }
      unsigned nib = 0;
      if (r >= 0 && r < lb && colvalid) {
        const int sub = s_matrix[qa + (int)sb[r]];
        int h, eo;
        nib = wf32_cell<MODE>(c, r, la, lb, sub, hleft_prev, e_sh, F, go, ge, h, eo, cand);
        hleft_prev = h_sh;
        h_out = h;
        e_out = eo;
        if (MODE == GLOBAL && c == la - 1 && r == lb - 1) corner = h;
        if (lane == 31 && b + 1 < nblk) {
          if (SMEM) { outH[r] = h; outE[r] = eo; } else { __stcg(outH + r, h); __stcg(outE + r, eo); }
          if (((r + 1) % WF32_ONE_CHUNK) == 0 || r + 1 == lb) { __threadfence_block(); s_progress[warp] = r + 1; }
        }
      }
      acc |= nib << ((s & 7) * 4);
      if ((s & 7) == 7 || s == nsteps - 1) {
        p.dir[((long long)b * steps8 + (s >> 3)) * 32 + lane] = acc;
        acc = 0;
      }
    }
    for (int off = 16; off > 0; off >>= 1) {
      EndCell o;
      o.score = __shfl_xor_sync(ABV_FULL, cand.score, off);
      o.col = __shfl_xor_sync(ABV_FULL, cand.col, off);
      o.row = __shfl_xor_sync(ABV_FULL, cand.row, off);
      if (end_better(o, cand)) cand = o;
    }
    if (cand.col >= 0 && cand.score >= wbest.score) wbest = cand;
    if (MODE == GLOBAL && b == nblk - 1) {
      corner = __shfl_sync(ABV_FULL, corner, (la - 1) & 31);
      if (lane == 0) s_corner = corner;
    }
    __syncwarp();
  }
  if (lane == 0) { s_best[3 * warp] = wbest.score; s_best[3 * warp + 1] = wbest.col; s_best[3 * warp + 2] = wbest.row; }
  __threadfence();
  __syncthreads();
  if (warp != 0) return;
  // warps own interleaved blocks: the overall end cell is the best by (score, then larger column)
  EndCell best = wf32_initial_candidate<MODE>();
  for (int w = 0; w < WF32_ONE_WARPS; ++w) {
    EndCell o; o.score = s_best[3 * w]; o.col = s_best[3 * w + 1]; o.row = s_best[3 * w + 2];
    if (o.col >= 0 && (best.col < 0 || end_better(o, best))) best = o;
  }
  int score, ec, er;
  wf32_finish<MODE>(best, MODE == GLOBAL ? s_corner : 0, la, lb, score, ec, er);
  int n = 0;
  if (lane == 0) {
    DirReader rd; rd.w = p.dir; rd.steps8 = steps8;
    n = walk_back<MODE>(rd, ec, er, p.tmp, p.tmp_cap);
    p.out[0] = score; p.out[1] = n;
  }
  n = __shfl_sync(ABV_FULL, n, 0);
  __syncwarp();
  FragOut *of = reinterpret_cast<FragOut *>(p.out + 4);
  for (int i = lane; i < n && i < p.out_frag_cap; i += 32) of[i] = p.tmp[n - 1 - i];
}

// ---------------------------------------------------------------------------------------------
// String assembly after the aligner, on the device:
//   ASSEMBLE_GAPPED  the two gapped rows of an alignment, as ambivert.py:95-111 / :136-152 build them
//   ASSEMBLE_MERGE   the merged read of a forward/reverse pair, ambivert.py:324-327:
//                    fwd[:fwd_start] + flatten_paired_alignment(rows) + rev[rev_start + rev bases aligned:]
//                    (flatten_paired_alignment / encode_ambiguous: sequence_utilities.py:130-169)
enum { ASSEMBLE_GAPPED = 0, ASSEMBLE_MERGE = 1 };
// per-pair status: 1 assembled; 0 merge skipped (aligned length < minimum_overlap: "unmergable");
// -1 empty alignment (the reference dereferences a NULL fragment: ValueError);
// -2 a mismatching base pair outside ACGTN (encode_ambiguous raises KeyError)
struct AssembleParams {
  int what, n, min_overlap;
  const uint8_t *a; const long long *a_off;      // original characters
  const uint8_t *b; const long long *b_off;
  const FragOut *frags; const long long *frag_start; const int *frag_count;
  long long *lens, *offs;                        // [n+1]
  int *status, *start_a, *start_b;
  uint8_t *out1, *out2;
};

__global__ void assemble_len_kernel(AssembleParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) p.lens[p.n] = 0;
  if (i >= p.n) return;
  const int cnt = p.frag_count[i];
  const FragOut *f = p.frags + p.frag_start[i];
  long long aligned = 0, b_used = 0;
  for (int k = 0; k < cnt; ++k) {
    aligned += f[k].hsp_len;
    if (f[k].type != FRAG_B_GAP) b_used += f[k].hsp_len;
  }
  const int sa = cnt ? f[0].sa_start : 0, sb = cnt ? f[0].sb_start : 0;
  p.start_a[i] = sa; p.start_b[i] = sb;
  int status = 1;
  long long len = aligned;
  if (p.what == ASSEMBLE_MERGE) {
    const long long lb = p.b_off[i + 1] - p.b_off[i];
    if (cnt == 0) { status = -1; len = 0; }
    else if (aligned < p.min_overlap) { status = 0; len = 0; }           // ambivert.py:324-325
    else len = sa + aligned + (lb - (sb + b_used));
  } else if (cnt == 0) status = -1;
  p.status[i] = status;
  p.lens[i] = len;
}

__device__ __forceinline__ int acgt_index(uint8_t c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : -1; }
__device__ __forceinline__ uint8_t ascii_upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? c - 32 : c; }

// one column of flatten_paired_alignment where neither row has a gap; 0 = KeyError in the reference
__device__ __forceinline__ uint8_t flatten_bases(uint8_t a, uint8_t b) {
  if (a == b) return a;
  const uint8_t ua = ascii_upper(a), ub = ascii_upper(b);
  if (ua == ub) return ua;
  if (ua == 'N' || ub == 'N') return 'N';
  const int ia = acgt_index(ua), ib = acgt_index(ub);
  if (ia < 0 || ib < 0) return 0;
  const char table[17] = "?MRWM?SYRS?KWYK?";       // IUPAC code of two different bases, indexed [ia*4+ib]
  return (uint8_t)table[ia * 4 + ib];
}

__global__ void assemble_kernel(AssembleParams p) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= p.n || p.status[i] != 1) return;
  const uint8_t *a = p.a + p.a_off[i], *b = p.b + p.b_off[i];
  const int lb = (int)(p.b_off[i + 1] - p.b_off[i]);
  const int cnt = p.frag_count[i];
  const FragOut *f = p.frags + p.frag_start[i];
  uint8_t *o1 = p.out1 + p.offs[i];
  uint8_t *o2 = p.what == ASSEMBLE_GAPPED ? p.out2 + p.offs[i] : nullptr;
  long long pos = 0;
  bool bad = false;
  if (p.what == ASSEMBLE_MERGE) {
    const int sa = f[0].sa_start;
    for (int j = lane; j < sa; j += 32) o1[j] = a[j];
    pos = sa;
  }
  int b_end = cnt ? f[0].sb_start : 0;
  for (int k = 0; k < cnt; ++k) {
    const FragOut fr = f[k];
    for (int j = lane; j < fr.hsp_len; j += 32) {
      if (p.what == ASSEMBLE_GAPPED) {
        o1[pos + j] = fr.type == FRAG_A_GAP ? (uint8_t)'-' : a[fr.sa_start + j];
        o2[pos + j] = fr.type == FRAG_B_GAP ? (uint8_t)'-' : b[fr.sb_start + j];
      } else if (fr.type == FRAG_MATCH) {
        const uint8_t c = flatten_bases(a[fr.sa_start + j], b[fr.sb_start + j]);
        if (c == 0) bad = true;
        o1[pos + j] = c;
      } else {
        o1[pos + j] = fr.type == FRAG_A_GAP ? b[fr.sb_start + j] : a[fr.sa_start + j];   // a gap takes the other row's base
      }
    }
    pos += fr.hsp_len;
    if (fr.type != FRAG_B_GAP) b_end += fr.hsp_len;
  }
  if (p.what == ASSEMBLE_MERGE) {
    for (int j = b_end + lane; j < lb; j += 32) o1[pos + (j - b_end)] = b[j];
    if (__any_sync(ABV_FULL, bad) && lane == 0) p.status[i] = -2;
  }
}

// ---------------------------------------------------------------------------------------------
// Variant scan of aligned row pairs (call_variants.py:32-128): one thread walks one
// pair of gapped rows with variants_walk (dp_core.cuh).  Two passes over the same walk: COUNT sizes the
// record list of every pair (then one exclusive scan places them), EMIT writes the records.  Byte work
// bound by HBM/L2: each row is read once per pass, 2 x row length bytes per pair; threads of a warp read
// neighbouring rows, whose 128-byte lines stay in L1 while the thread walks them.
struct VariantRec { int kind, start, pos, len, run_start; };
struct VariantParams {
  const uint8_t *sample; const long long *sample_off;     // gapped sample rows
  const uint8_t *ref; const long long *ref_off;           // gapped reference rows
  int n, softmask, gap;
  int *status;                 // [n]  VAR_OK / VAR_ERR_LENGTH / VAR_ERR_INDEX
  long long *count;            // [n+1] records per pair (COUNT pass; element n = 0 for the scan)
  const long long *start;      // [n+1] exclusive scan of count (EMIT pass)
  VariantRec *recs;
};

template <bool EMIT>
__global__ void variants_kernel(VariantParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (!EMIT && i == 0) p.count[p.n] = 0;
  if (i >= p.n) return;
  const long long s0 = p.sample_off[i], r0 = p.ref_off[i];
  const long long ls = p.sample_off[i + 1] - s0, lr = p.ref_off[i + 1] - r0;
  if (ls != lr) {                                      // RuntimeError, call_variants.py:54-55
    if (!EMIT) { p.status[i] = VAR_ERR_LENGTH; p.count[i] = 0; }
    return;
  }
  if (EMIT) {
    if (p.status[i] != VAR_OK) return;
    VariantRec *out = p.recs + p.start[i];
    int k = 0;
    variants_walk(p.sample + s0, p.ref + r0, (int)ls, p.softmask != 0, (unsigned char)p.gap,
                  [&](int kind, int start, int pos, int len, int at) { out[k++] = VariantRec{kind, start, pos, len, at}; });
  } else {
    int k = 0;
    const int rc = variants_walk(p.sample + s0, p.ref + r0, (int)ls, p.softmask != 0, (unsigned char)p.gap,
                                 [&](int, int, int, int, int) { ++k; });
    p.status[i] = rc;
    p.count[i] = rc == VAR_OK ? k : 0;
  }
}

// ---------------------------------------------------------------------------------------------
// Clustering front end: the key of a read pair is the MD5 (RFC 1321) of its trimmed
// forward + reverse sequence, md5(f_seq[trim5:trim3] + r_seq[trim5:trim3]) (ambivert.py:259); identical
// keys form one cluster.  One thread hashes one pair straight from the read buffers.
struct Md5Params {
  const uint8_t *f; const long long *f_off;
  const uint8_t *r; const long long *r_off;
  int n;
  int has5, trim5;      // Python slice start: None or an int (negative counts from the end)
  int has3, trim3;      // Python slice stop: None or -abs(trim3)
  unsigned long long *lo, *hi;   // digest bytes 0-7 and 8-15 as little-endian words
  unsigned int *index;           // 0..n-1 (the sort's payload)
};
__constant__ uint32_t c_md5_k[64];

__host__ __device__ inline void py_slice(long long len, int has_start, long long start, int has_stop, long long stop,
                                         long long &b, long long &e) {
  b = has_start ? (start < 0 ? (start + len < 0 ? 0 : start + len) : (start > len ? len : start)) : 0;
  e = has_stop ? (stop < 0 ? (stop + len < 0 ? 0 : stop + len) : (stop > len ? len : stop)) : len;
  if (e < b) e = b;
}

__global__ void md5_pairs_kernel(Md5Params p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const uint8_t *f = p.f + p.f_off[i], *r = p.r + p.r_off[i];
  long long fb, fe, rb, re;
  py_slice(p.f_off[i + 1] - p.f_off[i], p.has5, p.trim5, p.has3, p.trim3, fb, fe);
  py_slice(p.r_off[i + 1] - p.r_off[i], p.has5, p.trim5, p.has3, p.trim3, rb, re);
  const long long lf = fe - fb, L = lf + (re - rb);
  f += fb; r += rb;
  uint32_t a0 = 0x67452301u, b0 = 0xefcdab89u, c0 = 0x98badcfeu, d0 = 0x10325476u;
  const long long nblocks = (L + 9 + 63) / 64;
  const unsigned long long bits = (unsigned long long)L * 8ull;
  for (long long blk = 0; blk < nblocks; ++blk) {
    uint32_t M[16];
#pragma unroll
    for (int w = 0; w < 16; ++w) {
      uint32_t v = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const long long pos = blk * 64 + w * 4 + k;
        uint32_t byte;
        if (pos < lf) byte = f[pos];
        else if (pos < L) byte = r[pos - lf];
        else if (pos == L) byte = 0x80u;
        else if (blk == nblocks - 1 && w >= 14) byte = (uint32_t)((bits >> (8 * ((w - 14) * 4 + k))) & 0xffull);
        else byte = 0;
        v |= byte << (8 * k);
      }
      M[w] = v;
    }
    uint32_t A = a0, B = b0, C = c0, D = d0;
#pragma unroll
    for (int t = 0; t < 64; ++t) {
      uint32_t F; int g;
      if (t < 16) { F = (B & C) | (~B & D); g = t; }
      else if (t < 32) { F = (D & B) | (~D & C); g = (5 * t + 1) & 15; }
      else if (t < 48) { F = B ^ C ^ D; g = (3 * t + 5) & 15; }
      else { F = C ^ (B | ~D); g = (7 * t) & 15; }
      const int sh = t < 16 ? (t % 4 == 0 ? 7 : t % 4 == 1 ? 12 : t % 4 == 2 ? 17 : 22)
                   : t < 32 ? (t % 4 == 0 ? 5 : t % 4 == 1 ? 9 : t % 4 == 2 ? 14 : 20)
                   : t < 48 ? (t % 4 == 0 ? 4 : t % 4 == 1 ? 11 : t % 4 == 2 ? 16 : 23)
                            : (t % 4 == 0 ? 6 : t % 4 == 1 ? 10 : t % 4 == 2 ? 15 : 21);
      F = F + A + c_md5_k[t] + M[g];
      A = D; D = C; C = B;
      B = B + ((F << sh) | (F >> (32 - sh)));
    }
    a0 += A; b0 += B; c0 += C; d0 += D;
  }
  p.lo[i] = (unsigned long long)a0 | ((unsigned long long)b0 << 32);
  p.hi[i] = (unsigned long long)c0 | ((unsigned long long)d0 << 32);
  p.index[i] = (unsigned)i;
}

// after the two stable radix sorts (by lo, then by hi) equal digests are adjacent and, inside a run, in input order
__global__ void cluster_flag_kernel(const unsigned long long *lo, const unsigned long long *hi, int n, int *flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flag[i] = (i == 0 || lo[i] != lo[i - 1] || hi[i] != hi[i - 1]) ? 1 : 0;
}
// per run: where it starts in the sorted order and which input pair comes first
__global__ void cluster_runs_kernel(const int *flag, const int *seg, const unsigned int *index, int n, int *run_start, unsigned int *run_first, int *run_id) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  const int s = seg[i] - 1;
  run_start[s] = i; run_first[s] = index[i]; run_id[s] = s;
}
// clusters are numbered in order of first occurrence (dict insertion order of AmpliconData.data)
__global__ void cluster_emit_kernel(const int *order_run, int n_runs, const int *run_start, const unsigned int *run_first_sorted, int n,
                                    const unsigned long long *lo, const unsigned long long *hi,
                                    unsigned char *digest, int *count, int *first, int *rank_of_run) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_runs) return;
  const int s = order_run[k];
  const int b = run_start[s];
  // the run ends where the next run (in sorted order) starts
  rank_of_run[s] = k;
  first[k] = (int)run_first_sorted[k];
  const unsigned long long l = lo[b], h = hi[b];
  for (int j = 0; j < 8; ++j) { digest[16 * k + j] = (unsigned char)(l >> (8 * j)); digest[16 * k + 8 + j] = (unsigned char)(h >> (8 * j)); }
  count[k] = 0;   // filled by cluster_assign_kernel
}
__global__ void cluster_assign_kernel(const int *seg, const unsigned int *index, int n, const int *rank_of_run, int *cluster_of, int *count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = rank_of_run[seg[i] - 1];
  cluster_of[index[i]] = k;
  atomicAdd(count + k, 1);
}

// ---------------------------------------------------------------------------------------------
// per-query selection over a score matrix: strict '>' in target order from the seed
// (ambivert.py:478-485): first maximum wins; idx -1 when nothing beats the seed.
__global__ void argmax_rows_kernel(const int *__restrict__ scores, int n_rows, int n_t, int seed,
                                   int *__restrict__ best_score, int *__restrict__ best_idx, long long row_base,
                                   const int *__restrict__ row_map, int rows_by_query) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const long long qi = row_map ? (long long)row_map[row_base + row] : row_base + row;
  const int *s = scores + (rows_by_query ? qi : (long long)row) * n_t;
  int bs = INT_MIN, bi = INT_MAX;
  for (int j = lane; j < n_t; j += 32) {
    int v = s[j];
    if (v > bs) { bs = v; bi = j; }       // ascending j within a lane: first maximum kept
  }
  for (int off = 16; off > 0; off >>= 1) {
    int os = __shfl_xor_sync(ABV_FULL, bs, off), oi = __shfl_xor_sync(ABV_FULL, bi, off);
    if (os > bs || (os == bs && oi < bi)) { bs = os; bi = oi; }
  }
  if (lane == 0) {
    bool hit = n_t > 0 && bs > seed;
    best_score[qi] = hit ? bs : seed;
    best_idx[qi] = hit ? bi : -1;
  }
}

// ---------------------------------------------------------------------------------------------
// bulk async copy (TMA 1-D: SASS UBLKCP) global -> shared with mbarrier completion
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
  } while (!ok);
}

struct Mm16Params {
  const uint8_t *q; const long long *q_off;   // encoded queries
  const int *q_list; int n_list;              // the queries of this launch (one K bucket)
  const uint8_t *tp_codes; int tp_stride;     // target-pair symbol rows (bytes, stride multiple of 16)
  const int4 *tp_meta;                        // {len1, len2, idx1, idx2}; idx2 = -1: single target
  int n_pairs;
  const int *matrix_ext;                      // (alpha+1)^2, last row/col (the pad symbol) zero
  int a1;                                     // alpha + 1
  int go, ge, seed;
  int *best_score, *best_idx;                 // indexed by query
  int *all_scores; int n_t;                   // optional full score matrix [query*n_t + target]
  unsigned int *q_counter;
  unsigned int *tail_flag;                    // optional: set to 1 by the first CTA that finds no work item left -- the launch of the next
                                              // strip width waits for it on its own stream (cuStreamWaitValue32) and fills this grid's tail
  unsigned long long *best_key;               // per query: packed (score, target) of the best hit so far, see best_key_pack
  const int4 *items; int n_items;             // work items of this launch, built by the host: {query, first target pair, end target pair, -};
                                              // a persistent CTA takes the next one with an atomic counter
  int lv_n, lv0, lv_step;                     // streaming global kernel: level plan (dp_core.cuh, mm16g_levels); lv_n <= 1: every pair is reset explicitly
};

// Per-query best hits are combined across work items with one 64-bit atomicMax: higher score wins, then
// the lower target index (the first maximum in target order, ambivert.py:478-485).  0 = no hit yet.
__host__ __device__ inline unsigned long long best_key_pack(int score, int idx) {
  return ((unsigned long long)((unsigned)score ^ 0x80000000u) << 32) | (unsigned)(0x7fffffff - idx);
}

__global__ void best_key_finalize_kernel(const unsigned long long *__restrict__ keys, const int *__restrict__ q_list, int n, int seed,
                                         int *__restrict__ best_score, int *__restrict__ best_idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int qi = q_list[i];
  const unsigned long long k = keys[qi];
  const int bs = (int)((unsigned)(k >> 32) ^ 0x80000000u), bi = 0x7fffffff - (int)(unsigned)k;
  const bool hit = k != 0ull && bs > seed;            // strict '>' against the seed (ambivert.py:482)
  best_score[qi] = hit ? bs : seed;
  best_idx[qi] = hit ? bi : -1;
}

#ifndef ABV_MM16_WARPS
#define ABV_MM16_WARPS 8
#endif
constexpr int MM16_WARPS = ABV_MM16_WARPS;

template <int K>
struct Mm16Cfg {
  static constexpr int V = (K % 4 == 0) ? 4 : 2;     // words per shared-memory vector load
  static constexpr int P = 32 * K;                   // profile row length in words
  static constexpr int MIN_BLOCKS = K <= 8 ? ABV_MINB8 : (K <= 12 ? 3 : 2);
  static constexpr int KP = StripCfg<K>::KP;
};

template <int K>
__device__ __forceinline__ void mm16_load_S(const uint32_t *lut_row, int lane, uint32_t (&S)[K]) {
  if (Mm16Cfg<K>::V == 4) {
    const uint4 *rp = reinterpret_cast<const uint4 *>(lut_row) + lane;
#pragma unroll
    for (int i = 0; i < K / 4; ++i) {
      uint4 v = rp[i * 32];
      S[4 * i] = v.x; S[4 * i + 1] = v.y; S[4 * i + 2] = v.z; S[4 * i + 3] = v.w;
    }
  } else {
    const uint2 *rp = reinterpret_cast<const uint2 *>(lut_row) + lane;
#pragma unroll
    for (int i = 0; i < K / 2; ++i) {
      uint2 v = rp[i * 32];
      S[2 * i] = v.x; S[2 * i + 1] = v.y;
    }
  }
}

// The streaming kernel addresses its profile row as ONE multiply-add per step: lane_addr (shared-window address of the lane's
// first vector, fixed per CTA) + code * row bytes; the vectors of a row are 32 lanes apart (immediate offsets).
template <int KP>
__device__ __forceinline__ void mm16g_load_S(uint32_t lane_addr, int code, uint32_t (&S)[KP]) {
  constexpr int V = (KP % 4 == 0) ? 4 : 2;
  const uint32_t a = lane_addr + (uint32_t)code * (32u * KP * 4u);
  if (V == 4) {
#pragma unroll
    for (int i = 0; i < KP / 4; ++i)
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(S[4 * i]), "=r"(S[4 * i + 1]), "=r"(S[4 * i + 2]), "=r"(S[4 * i + 3]) : "r"(a + i * 512u));
  } else {
#pragma unroll
    for (int i = 0; i < KP / 2; ++i)
      asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(S[2 * i]), "=r"(S[2 * i + 1]) : "r"(a + i * 256u));
  }
}

template <int K>
__device__ __forceinline__ uint32_t mm16_pick(const uint32_t (&Hp)[K], int kstar) {
  uint32_t v = Hp[0];
#pragma unroll
  for (int k = 1; k < K; ++k) v = (k == kstar) ? Hp[k] : v;
  return v;
}

__device__ __forceinline__ uint32_t warp_max2(uint32_t v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = p16::max2(v, __shfl_xor_sync(ABV_FULL, v, off));
  return v;
}

// dynamic shared memory layout of mm16_kernel (offsets in bytes); host and device agree through this
struct Mm16Smem {
  int lut, qcodes, mext, tbuf, mbar, wres, misc, total;
};
__host__ __device__ inline Mm16Smem mm16_smem_layout(int K, int a1, int tp_stride) {
  Mm16Smem L;
  int off = 0;
  L.lut = off; off += a1 * a1 * 32 * K * 4;
  L.tbuf = off; off += MM16_WARPS * 2 * tp_stride;
  L.mbar = off; off += MM16_WARPS * 2 * 8;
  L.mext = off; off += a1 * a1 * 4;
  L.wres = off; off += MM16_WARPS * 2 * 4;
  L.misc = off; off += 16;
  L.qcodes = off; off += 32 * K;
  L.total = (off + 15) & ~15;
  return L;
}

template <int MODE, int K>
__global__ void __launch_bounds__(MM16_WARPS * 32, Mm16Cfg<K>::MIN_BLOCKS) mm16_kernel(Mm16Params p) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int P = Mm16Cfg<K>::P;
  constexpr int V = Mm16Cfg<K>::V;
  const Mm16Smem L = mm16_smem_layout(K, p.a1, p.tp_stride);
  uint32_t *lut = reinterpret_cast<uint32_t *>(smem + L.lut);
  uint8_t *qcodes = smem + L.qcodes;
  int *mext = reinterpret_cast<int *>(smem + L.mext);
  int *wres = reinterpret_cast<int *>(smem + L.wres);
  int *misc = reinterpret_cast<int *>(smem + L.misc);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t *tbuf0 = smem + L.tbuf + warp * 2 * p.tp_stride;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L.mbar) + warp * 2;

  const int a1 = p.a1, ncodes = a1 * a1;
  const int go = p.go, ge = p.ge;
  const uint32_t go2 = p16::pack(go, go), ge2 = p16::pack(ge, ge);
  const uint32_t NEG2 = p16::pack(p16::NEG16, p16::NEG16);

  for (int i = threadIdx.x; i < ncodes; i += blockDim.x) mext[i] = p.matrix_ext[i];
  if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
  mbar_fence_init();
  uint32_t phase0 = 0, phase1 = 0;
  __syncthreads();

  for (;;) {
    if (threadIdx.x == 0) misc[0] = (int)atomicAdd(p.q_counter, 1u);
    __syncthreads();
    const int item = misc[0];
    if (item >= p.n_items) {
      if (threadIdx.x == 0 && p.tail_flag) atomicExch(p.tail_flag, 1u);
      break;
    }
    const int4 work = p.items[item];
    const int qi = work.x, pair_lo = work.y, pair_hi = work.z;
    const uint8_t *qs = p.q + p.q_off[qi];
    const int lq = (int)(p.q_off[qi + 1] - p.q_off[qi]);
    for (int c = threadIdx.x; c < P; c += blockDim.x) qcodes[c] = c < lq ? qs[c] : (uint8_t)(a1 - 1);
    __syncthreads();
    // per-query profile: lut[code][pos(c)] = (S(q[c], a), S(q[c], b)), code = a*a1 + b
    for (int idx = threadIdx.x; idx < ncodes * P; idx += blockDim.x) {
      const int code = idx / P, pos = idx - code * P;
      const int j = pos % V, it = pos / V, t = it & 31, i = it >> 5;
      const int c = t * K + i * V + j;
      const int qa = qcodes[c] * a1;
      const int a = code / a1, bsym = code - a * a1;
      lut[idx] = p16::pack(mext[qa + a], mext[qa + bsym]);
    }
    __syncthreads();

    const int tstar = (lq - 1) / K, kstar = (lq - 1) - tstar * K;
    const int c0 = lane * K;
    int wbest = INT_MIN, widx = INT_MAX;

    int pi = pair_lo + warp;
    if (pi < pair_hi && lane == 0) bulk_load(tbuf0, p.tp_codes + (long long)pi * p.tp_stride, p.tp_stride, &bar[0]);
    int it = 0;
    for (; pi < pair_hi; pi += MM16_WARPS, ++it) {
      const int cur = it & 1;
      const int pn = pi + MM16_WARPS;
      __syncwarp();
      if (pn < pair_hi && lane == 0)
        bulk_load(tbuf0 + (cur ^ 1) * p.tp_stride, p.tp_codes + (long long)pn * p.tp_stride, p.tp_stride, &bar[cur ^ 1]);
      const int4 meta = p.tp_meta[pi];
      if (cur == 0) { mbar_wait(&bar[0], phase0); phase0 ^= 1; }
      else { mbar_wait(&bar[1], phase1); phase1 ^= 1; }
      const uint8_t *tb = tbuf0 + cur * p.tp_stride;
      const int L1m1 = meta.x - 1, L2m1 = meta.y - 1;
      const int lmax = max(meta.x, meta.y);

      uint32_t Hp[K], F[K], S[K];
      uint32_t hleft_prev, h_out = 0, e_out = 0, best2 = 0;
      int cap1 = 0, cap2 = 0;
      int sc1, sc2;

      if (MODE == GLOBAL) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const int v = go + ge * (c0 + k);            // H[c][-1]  (global_align.c:148)
          Hp[k] = p16::pack(v, v);
          F[k] = p16::pack(v + go, v + go);            // F entering row 0 (global_align.c:146)
        }
        { const int v = c0 ? go + ge * (c0 - 1) : 0; hleft_prev = p16::pack(v, v); }
        int vcol = go;                                 // H[-1][r] of the virtual column, r = s on lane 0
        const int nsteps = lmax + tstar;
        for (int s = 0; s < nsteps; ++s) {
          uint32_t h_sh = __shfl_up_sync(ABV_FULL, h_out, 1);
          uint32_t e_sh = __shfl_up_sync(ABV_FULL, e_out, 1);
          if (lane == 0) { h_sh = p16::pack(vcol, vcol); e_sh = p16::pack(vcol + go, vcol + go); vcol += ge; }
          const int r = s - lane;
          if (r >= 0 && r < lmax && lane <= tstar) {
            mm16_load_S<K>(lut + (int)tb[r] * P, lane, S);
            uint32_t e = e_sh;
            mm16_row<K, false, false>(hleft_prev, e, Hp, F, S, go2, ge2, best2);
            hleft_prev = h_sh; e_out = e; h_out = Hp[K - 1];
            if (lane == tstar && (r == L1m1 || r == L2m1)) {
              const uint32_t hs = mm16_pick<K>(Hp, kstar);
              if (r == L1m1) cap1 = p16::lo(hs);
              if (r == L2m1) cap2 = p16::hi(hs);
            }
          }
        }
        sc1 = __shfl_sync(ABV_FULL, cap1, tstar);
        sc2 = __shfl_sync(ABV_FULL, cap2, tstar);
      } else if (MODE == LOCAL) {
#pragma unroll
        for (int k = 0; k < K; ++k) { Hp[k] = 0; F[k] = go2; }   // local_align.c:123-135
        hleft_prev = 0;
        const int nsteps = lmax + tstar;
        for (int s = 0; s < nsteps; ++s) {
          uint32_t h_sh = __shfl_up_sync(ABV_FULL, h_out, 1);
          uint32_t e_sh = __shfl_up_sync(ABV_FULL, e_out, 1);
          if (lane == 0) { h_sh = 0; e_sh = go2; }
          const int r = s - lane;
          if (r >= 0 && r < lmax && lane <= tstar) {
            mm16_load_S<K>(lut + (int)tb[r] * P, lane, S);
            uint32_t e = e_sh;
            mm16_row<K, true, true>(hleft_prev, e, Hp, F, S, go2, ge2, best2);
            hleft_prev = h_sh; e_out = e; h_out = Hp[K - 1];
          }
        }
        best2 = warp_max2(best2);
        sc1 = p16::lo(best2); sc2 = p16::hi(best2);
      } else if (MODE == OVERLAP) {
        // row 0 has no dependencies: H[c][0] = S, F = H + go (align.c:135-138); see dp_core.cuh, mm16_overlap_cands
        mm16_load_S<K>(lut + (int)tb[0] * P, lane, S);
#pragma unroll
        for (int k = 0; k < K; ++k) { Hp[k] = S[k]; F[k] = p16::add(S[k], go2); }
        hleft_prev = __shfl_up_sync(ABV_FULL, Hp[K - 1], 1);
        if (lane == 0) hleft_prev = 0;                       // the first column is forced: diagonal 0, no gap states
        best2 = NEG2;
        if (lane <= tstar)
          mm16_overlap_cands<K>(Hp, c0, lq, lane == tstar ? kstar : -1, 0xffffffffu,
                                (L1m1 == 0 ? 0x0000ffffu : 0u) | (L2m1 == 0 ? 0xffff0000u : 0u), best2);
        const int nsteps = lmax - 1 + tstar;
        for (int s = 0; s < nsteps; ++s) {
          uint32_t h_sh = __shfl_up_sync(ABV_FULL, h_out, 1);
          uint32_t e_sh = __shfl_up_sync(ABV_FULL, e_out, 1);
          if (lane == 0) { h_sh = 0; e_sh = NEG2; }
          const int r = 1 + s - lane;
          if (r >= 1 && r < lmax && lane <= tstar) {
            mm16_load_S<K>(lut + (int)tb[r] * P, lane, S);
            if (lane == 0) F[0] = NEG2;
            uint32_t e = e_sh, dummy = 0;
            mm16_row<K, false, false>(hleft_prev, e, Hp, F, S, go2, ge2, dummy);
            hleft_prev = h_sh; e_out = e; h_out = Hp[K - 1];
            const bool l1 = (r == L1m1), l2 = (r == L2m1);
            if (lane == tstar || l1 || l2)
              mm16_overlap_cands<K>(Hp, c0, lq, lane == tstar ? kstar : -1, (r <= L1m1 ? 0x0000ffffu : 0u) | (r <= L2m1 ? 0xffff0000u : 0u),
                                    (l1 ? 0x0000ffffu : 0u) | (l2 ? 0xffff0000u : 0u), best2);
          }
        }
        best2 = warp_max2(best2);
        sc1 = p16::lo(best2); sc2 = p16::hi(best2);
      } else {  // GLOCAL
        // row 0 has no dependencies: H[c][0] = S(q[c], t[0]), F = H + go (glocal_align.c:128-136)
        mm16_load_S<K>(lut + (int)tb[0] * P, lane, S);
#pragma unroll
        for (int k = 0; k < K; ++k) { Hp[k] = S[k]; F[k] = p16::add(S[k], go2); }
        hleft_prev = __shfl_up_sync(ABV_FULL, Hp[K - 1], 1);
        if (lane == 0) hleft_prev = NEG2;
        best2 = NEG2;
        const int nsteps = lmax - 1 + tstar;
        for (int s = 0; s < nsteps; ++s) {
          uint32_t h_sh = __shfl_up_sync(ABV_FULL, h_out, 1);
          uint32_t e_sh = __shfl_up_sync(ABV_FULL, e_out, 1);
          if (lane == 0) { h_sh = NEG2; e_sh = NEG2; }
          const int r = 1 + s - lane;
          if (r >= 1 && r < lmax && lane <= tstar) {
            mm16_load_S<K>(lut + (int)tb[r] * P, lane, S);
            uint32_t e = e_sh;
            const bool l1 = (r == L1m1), l2 = (r == L2m1);
            if (l1 || l2) {
              const uint32_t mask = (l1 ? 0x0000ffffu : 0u) | (l2 ? 0xffff0000u : 0u);
              mm16_row_glocal_last<K>(hleft_prev, e, Hp, F, S, go2, ge2, mask, c0, lq, best2);
            } else {
              uint32_t dummy = 0;
              mm16_row<K, false, false>(hleft_prev, e, Hp, F, S, go2, ge2, dummy);
            }
            hleft_prev = h_sh; e_out = e; h_out = Hp[K - 1];
          }
        }
        best2 = warp_max2(best2);
        sc1 = max(0, p16::lo(best2)); sc2 = max(0, p16::hi(best2));   // max starts at 0 (glocal_align.c:115)
      }

      // fold this pair's two scores into the warp's running best: higher score, then lower index
      if (sc1 > wbest || (sc1 == wbest && meta.z < widx)) { wbest = sc1; widx = meta.z; }
      if (meta.w >= 0 && (sc2 > wbest || (sc2 == wbest && meta.w < widx))) { wbest = sc2; widx = meta.w; }
      if (p.all_scores && lane == 0) {
        p.all_scores[(long long)qi * p.n_t + meta.z] = sc1;
        if (meta.w >= 0) p.all_scores[(long long)qi * p.n_t + meta.w] = sc2;
      }
    }

    if (lane == 0) { wres[2 * warp] = wbest; wres[2 * warp + 1] = widx; }
    __syncthreads();
    if (threadIdx.x == 0) {
      int bs = INT_MIN, bi = INT_MAX;
      for (int w = 0; w < MM16_WARPS; ++w) {
        const int s = wres[2 * w], i = wres[2 * w + 1];
        if (i != INT_MAX && (s > bs || (s == bs && i < bi))) { bs = s; bi = i; }
      }
      if (bi != INT_MAX) atomicMax(p.best_key + qi, best_key_pack(bs, bi));   // seed rule applied by best_key_finalize_kernel
    }
    // the next iteration's first __syncthreads orders these reads before wres/lut are rewritten
  }
}

// ---------------------------------------------------------------------------------------------
// mm16g_kernel: the GLOBAL many-to-many scorer, second generation.  Same data layout and arithmetic
// idea as mm16_kernel, plus:
//   * target pairs STREAM through the lane pipeline back to back: lane t switches to the next pair
//     t steps after lane 0, so there is no ramp-up/ramp-down per pair and all 32 lanes work in every
//     step (mm16_kernel: 25 of 32 on 240-column queries);
//   * biased operands (dp_core.cuh, namespace b16): the two adds of a cell are plain 32-bit adds that
//     can issue on the FMA pipe; the half-rate ALU/DPX pipe only sees max3 + 2 add-max per cell pair;
//   * the hot loop has no per-step predicates: pair switches happen in the first 32 steps of a pair
//     (one lane per step), corner captures at warp-uniform steps outside the inner loop;
//   * three row buffers per warp; the copy for pair i+2 is issued once every lane has left pair i-1.
struct Mm16gSmem {
  int lut, tbuf, mbar, mext, wres, misc, zero, slots, qcodes, evs, total;
};
// glocal, K <= 9: the diagonal sums of a lane's last target row are parked in shared memory at the step the lane
// passes that row and folded once per pair at the capture ("deferred" last rows); wider strips keep the in-loop fold
__host__ __device__ constexpr bool mm16g_defer(bool glocal, int K) { return glocal && K <= 9; }
__host__ __device__ inline Mm16gSmem mm16g_smem_layout(int K, int a1, int tp_stride, bool defer) {
  Mm16gSmem L;
  int off = 0;
  const int KP = (K % 2) ? ((K + 3) & ~3) : K;             // StripCfg<K>::KP
  L.lut = off; off += a1 * a1 * 32 * KP * 4;
  L.tbuf = off; off += MM16_WARPS * 3 * tp_stride;
  L.mbar = off; off += MM16_WARPS * 3 * 8;
  L.mext = off; off += a1 * a1 * 4;
  L.wres = off; off += MM16_WARPS * 2 * 4;
  L.misc = off; off += 16;
  L.zero = off; off += 64;
  L.slots = off; off += MM16_WARPS * 4 * 4;
  L.qcodes = off; off += 32 * KP;
  off = (off + 15) & ~15;
  L.evs = off; off += defer ? MM16_WARPS * 2 * 32 * KP * 4 : 0;     // per warp: [half][vector][lane] words
  L.total = (off + 15) & ~15;
  return L;
}

// GOPC: gap_open - gap_extend as a COMPILE-TIME constant (<= 0), or MM16G_GOP_RUNTIME.  The specialised instantiation (the
// pipeline's scoring, -7 - -1 = -6: every call site of the reference, ambivert.py:86-89, :127-130, :170-173) carries the
// addend of the two add-max instructions of a cell as an immediate, so they read two registers instead of three (fewer
// register-bank conflicts on the half-rate pipe that bounds the kernel) and one register is freed.
constexpr int MM16G_GOP_RUNTIME = 1;
constexpr int MM16G_GOP_DNA = -6;
// BORDER: lane 31 emits the constants of column -1 (dp_core.cuh, mm16g_border) and the strip width serves 31K columns; the
// queries of (31K, 32K] columns take the instantiation that keeps all 32 lanes and selects the constants into lane 0.
template <int MODE, int K, int GOPC, bool BORDER>
__global__ void __launch_bounds__(MM16_WARPS * 32, Mm16Cfg<K>::MIN_BLOCKS) mm16g_kernel(Mm16Params p) {
  static_assert(MODE == GLOBAL || MODE == GLOCAL, "the streaming kernel implements global and glocal");
  constexpr bool GL = (MODE == GLOCAL);
  static_assert(!BORDER || mm16g_border(K, GL), "no border lane for this strip width / mode");
  constexpr bool DEFER = mm16g_defer(GL, K);
  constexpr int BL = MM16G_BORDER_LANE;
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int KP = StripCfg<K>::KP;
  constexpr int P = StripCfg<K>::P;
  constexpr int V = StripCfg<K>::V;
  constexpr int W = MM16_WARPS;
  const Mm16gSmem L = mm16g_smem_layout(K, p.a1, p.tp_stride, DEFER);
  uint32_t *lut = reinterpret_cast<uint32_t *>(smem + L.lut);
  uint8_t *qcodes = smem + L.qcodes;
  int *mext = reinterpret_cast<int *>(smem + L.mext);
  int *wres = reinterpret_cast<int *>(smem + L.wres);
  int *misc = reinterpret_cast<int *>(smem + L.misc);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tbuf0 = L.tbuf + warp * 3 * p.tp_stride;
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + L.mbar) + warp * 3;
  int *slots = reinterpret_cast<int *>(smem + L.slots) + warp * 4;       // glocal: [pair parity][half] running last-row maxima
  const int a1 = p.a1, ncodes = a1 * a1;
  const int go = p.go, ge = p.ge;
  const uint32_t gop2 = GOPC <= 0 ? p16::pack(GOPC, GOPC) : p16::pack(go - ge, go - ge);
  const uint32_t gopw = b16::addend(go - ge, go - ge), gew = b16::addend(ge, ge);
  const uint32_t NEGb = b16::biased(p16::NEG16, p16::NEG16);
  // values handed to column 0 by the virtual column -1 (drift-normalised, see mm16d_row):
  //   global: H~ = go + g, E~ = 2go (global_align.c:136-140); glocal: minus infinity (glocal_align.c:119-123)
  // (global: the constants of level 0; a pair on level v holds every value v higher, see mm16g_levels)
  const uint32_t Hc0 = GL ? NEGb : b16::biased(go + ge, go + ge);
  const uint32_t Ec0 = GL ? NEGb : b16::biased(2 * go, 2 * go);
  const uint32_t H000 = b16::biased(2 * ge, 2 * ge);      // global: H~[-1][-1]
  const int lv_n = GL ? 1 : max(p.lv_n, 1);
  const uint32_t NOEGAP = p16::pack(MM16G_NO_EGAP, MM16G_NO_EGAP);
  const uint32_t negg = p16::pack(-ge, -ge);             // glocal: step of the last-row maximum chain

  for (int i = threadIdx.x; i < ncodes; i += blockDim.x) mext[i] = p.matrix_ext[i];
  if (threadIdx.x < 16) reinterpret_cast<uint32_t *>(smem + L.zero)[threadIdx.x] = 0u;
  if (lane < 4) slots[lane] = INT_MIN;
  if (lane == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(&bar[2], 1); }
  mbar_fence_init();
  uint32_t phbits = 0;
  __syncthreads();

  for (;;) {
    if (threadIdx.x == 0) { misc[0] = (int)atomicAdd(p.q_counter, 1u); misc[1] = 0; }
    __syncthreads();
    const int item = misc[0];
    if (item >= p.n_items) {
      if (threadIdx.x == 0 && p.tail_flag) atomicExch(p.tail_flag, 1u);
      break;
    }
    const int4 work = p.items[item];
    const int qi = work.x, pair_lo = work.y, pair_hi = work.z;
    const uint8_t *qs = p.q + p.q_off[qi];
    const int lq = (int)(p.q_off[qi + 1] - p.q_off[qi]);
    for (int c = threadIdx.x; c < P; c += blockDim.x) qcodes[c] = c < lq ? qs[c] : (uint8_t)(a1 - 1);
    __syncthreads();
    for (int idx = threadIdx.x; idx < ncodes * P; idx += blockDim.x) {
      const int code = idx / P, pos = idx - code * P;
      const int j = pos % V, it = pos / V, t = it & 31, i = it >> 5;
      const int k = i * V + j;                                // k >= K: padding word of an odd strip, never used
      const int qa = (k < K ? (int)qcodes[t * K + k] : a1 - 1) * a1;
      const int a = code / a1, bsym = code - a * a1;
      lut[idx] = (BORDER && t == BL) ? 0u : b16::addend(mext[qa + a] - 2 * ge, mext[qa + bsym] - 2 * ge);
    }
    __syncthreads();

    const int tstar = (lq - 1) / K, kstar = (lq - 1) - tstar * K;      // BORDER: tstar <= 30 (the host's bucket capacity, mm16g_capacity)
    const int c0 = lane * K;
    // glocal: H~[c0][-1] of a starting pair (dp_core.cuh, MM16G_NO_EGAP).  The border lane takes a value from which its switch
    // (Hp[k] = G0 + k|g|, hleft = G0 + g) lands at or below -inf: one row later it is back on its fixed point
    const uint32_t G0 = (BORDER && lane == BL) ? NEGb + (uint32_t)(K - 1) * gew : b16::biased(ge * (1 - c0), ge * (1 - c0));
    // The warps of the CTA share the item's slice [pair_lo, pair_hi) of the target pairs.  Each warp takes its next pair from a
    // counter in shared memory at the moment it issues the pair's copy (two pairs ahead), so that all eight warps run out of pairs
    // within one pair's duration of each other: handed out strided, the warps (two per SM sub-partition, progressing at different
    // rates) reached the item-end barrier spread over 5.7 % of the kernel's warp time (ncu, profiles/r02_*).
    auto take_pair = [&](int strided) {                            // warp-uniform; >= pair_hi: no pair left
      if (!kDynPairs) return strided;
      int v = 0;
      if (lane == 0) v = atomicAdd(&misc[1], 1);
      return pair_lo + __shfl_sync(ABV_FULL, v, 0);
    };
    int p_cur = take_pair(pair_lo + warp);
    int p_next = p_cur < pair_hi ? take_pair(pair_lo + warp + W) : pair_hi;
    int wbest = INT_MIN, widx = INT_MAX;

    if (p_cur < pair_hi) {
      if (lane == 0) {
        bulk_load(smem + tbuf0, p.tp_codes + (long long)p_cur * p.tp_stride, p.tp_stride, &bar[0]);
        if (p_next < pair_hi) bulk_load(smem + tbuf0 + p.tp_stride, p.tp_codes + (long long)p_next * p.tp_stride, p.tp_stride, &bar[1]);
      }
      uint32_t Hp[K], F[K], S[KP];
#pragma unroll
      for (int k = 0; k < K; ++k) { Hp[k] = b16::biased(0, 0); F[k] = b16::biased(0, 0); }
      uint32_t hleft = b16::biased(0, 0), h_out = b16::biased(0, 0), e_out = b16::biased(0, 0);
      int tb_off = L.zero + 32 - lane;           // before its first pair a lane sweeps zero (valid) symbols
      int prev_rows = 0;
      int pend_s1 = -1, pend_s2 = -1, pend_i1 = -1, pend_i2 = -1, pend_d1 = 0, pend_d2 = 0;
      int4 meta_next = p.tp_meta[p_cur];
      // glocal: last-row events of the previous pair that are still to come (they are spread over the
      // lanes like the pair switches: lane t reaches row r at step r + t of the pair's own time line)
      const int NOEV = 1 << 20;
      int pend_e1 = NOEV, pend_e2 = NOEV, pend_r1 = 0, pend_r2 = 0;
      uint32_t lm0 = NEGb, lm1 = NEGb;       // glocal: this lane's last-row maxima of the pairs of even / odd parity (both halves)

#if defined(ABV_DUMMY_FMA) || defined(ABV_DUMMY_ALU)
      uint32_t dummy[4] = {(uint32_t)lane, (uint32_t)warp, 7u, 11u};
#endif
      uint32_t Hc = Hc0, Ec = Ec0, H00 = H000;       // boundary constants of the pair lane 0 is on (its level added)
      int lv_i = 0;                                  // position of the current pair on the staircase of levels
      const uint32_t lut_lane = smem_u32(lut) + (uint32_t)lane * (V * 4u);      // this lane's first vector of profile row 0
      const int src_lane = lane == 0 ? BL : (lane == BL ? BL : lane - 1);   // BORDER: who a lane receives its left border from
      auto exchange = [&](uint32_t &h_sh, uint32_t &e_sh) {
        if (BORDER) {
          h_sh = __shfl_sync(ABV_FULL, h_out, src_lane);
          e_sh = __shfl_sync(ABV_FULL, e_out, src_lane);
        } else {
          h_sh = __shfl_up_sync(ABV_FULL, h_out, 1);
          e_sh = __shfl_up_sync(ABV_FULL, e_out, 1);
          if (lane == 0) { h_sh = Hc; e_sh = Ec; }
        }
      };
      // the border lane on the fixed point of the constants (hc, ec) (dp_core.cuh, mm16g_border)
      auto border_set = [&](uint32_t hc, uint32_t ec) {
        if (BORDER && lane == BL) {
#pragma unroll
          for (int k = 0; k < K; ++k) { Hp[k] = hc; F[k] = ec; }
          hleft = hc; h_out = hc; e_out = ec;
        }
      };
      if (GL) border_set(NEGb, NEGb);                // glocal: column -1 is minus infinity for every pair
      auto row_e = [&](int s, uint32_t h_sh, uint32_t e_sh, uint32_t gope) {   // one row of this lane; gope: addend of the row-gap chain
        const int code = smem[tb_off + s];
#if ABV_LDS_PTX
        mm16g_load_S<KP>(lut_lane, code, S);
#else
        mm16_load_S<KP>(lut + code * P, lane, S);
#endif
        uint32_t e = e_sh;
        mm16d_row<K, KP>(hleft, e, Hp, F, S, gop2, gope);
        hleft = h_sh; e_out = e; h_out = Hp[K - 1];
      };
      auto row = [&](int s, uint32_t h_sh, uint32_t e_sh) { row_e(s, h_sh, e_sh, gop2); };   // one ordinary row
      auto body = [&](int s) {
        uint32_t h_sh, e_sh;
        exchange(h_sh, e_sh);
        row(s, h_sh, e_sh);
#ifdef ABV_DUMMY_FMA      // sensitivity experiment: extra FMA-pipe instructions per step (tools/sweep_builds.sh)
#pragma unroll
        for (int d = 0; d < ABV_DUMMY_FMA; ++d) asm volatile("mad.lo.u32 %0, %0, 3, 1;" : "+r"(dummy[d & 3]));
#endif
#ifdef ABV_DUMMY_ALU      // ... and extra ALU-pipe instructions per step
#pragma unroll
        for (int d = 0; d < ABV_DUMMY_ALU; ++d) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(dummy[d & 3]) : "r"(dummy[(d + 1) & 3]), "r"(dummy[(d + 2) & 3]));
#endif
      };
      // glocal, the lane is on the last target row r of a half: the forced diagonal (glocal_align.c:172-183)
      // is what S[] holds after the row's first pass; every real query column competes for the end
      auto last_row = [&](bool h1, bool h2, int r, int parity) {
        if (lane > tstar) return;                                           // only padding columns
        // max_k (S[k] + g*k) as the chain n_k = max(n_{k-1} - g, S[k]) (one add-max per column), then + g*k_last
        uint32_t m = NEGb;
        const uint32_t d0 = b16::addend(ge * (c0 + r), ge * (c0 + r));          // undo the drift: + g*(c + r)
        if (lane < tstar) {
#pragma unroll
          for (int k = 0; k < K; ++k) m = p16::addmax(m, negg, S[k]);
          m += d0 + (uint32_t)(K - 1) * gew;
        } else {
#pragma unroll
          for (int k = 0; k < K; ++k)
            if (k <= kstar) m = p16::addmax(m, negg, S[k]);
          m += d0 + (uint32_t)kstar * gew;
        }
        // the lane's own maximum of this (pair, half): one event per lane, folded over the lanes at the capture
        const uint32_t mm = (h1 ? (m & 0xffffu) : (NEGb & 0xffffu)) | (h2 ? (m & 0xffff0000u) : (NEGb & 0xffff0000u));
        if (parity) lm1 = p16::max2(lm1, mm); else lm0 = p16::max2(lm0, mm);
      };
      // lane t is on the last row of half h of a pair at step ev_h + t of that pair's time line
      auto last_rows_at = [&](int s, int e1, int e2, int r1, int r2, int parity) {
        const bool h1 = (lane == s - e1), h2 = (lane == s - e2);
        if (h1 || h2) last_row(h1, h2, h1 ? r1 : r2, parity);
      };
      auto capture = [&](int half, int idx, int drift, int parity) {
        if (GL) {
          const uint32_t v = warp_max2(parity ? lm1 : lm0);                // every lane has passed this half's last row
          const int sc = max(0, half ? b16::hi(v) : b16::lo(v));           // max starts at 0 (glocal_align.c:115)
          if (sc > wbest || (sc == wbest && idx < widx)) { wbest = sc; widx = idx; }
          if (p.all_scores && lane == 0) p.all_scores[(long long)qi * p.n_t + idx] = sc;
          const uint32_t keep = half ? 0x0000ffffu : 0xffff0000u;           // this half starts over for the pair after next
          if (parity) lm1 = (lm1 & keep) | (NEGb & ~keep); else lm0 = (lm0 & keep) | (NEGb & ~keep);
        } else if (lane == tstar) {
          const uint32_t hs = mm16_pick<K>(Hp, kstar);
          const int sc = (half ? b16::hi(hs) : b16::lo(hs)) + drift;     // undo the drift: + g*(c + r) of the corner
          if (sc > wbest || (sc == wbest && idx < widx)) { wbest = sc; widx = idx; }
          if (p.all_scores) p.all_scores[(long long)qi * p.n_t + idx] = sc;
        }
      };

      for (int i = 0;; ++i) {
        const bool real = p_cur < pair_hi;                    // the iteration after the last pair drains the lane pipeline
        int4 meta = make_int4(0, 0, -1, -1);
        int rows = MM16G_MIN_ROWS, bufoff = L.zero + 32;
        if (real) {
          meta = meta_next;
          if (p_next < pair_hi) meta_next = p.tp_meta[p_next];
          rows = mm16g_rows(meta.x, meta.y);
          const int b = i % 3;
          bufoff = tbuf0 + b * p.tp_stride;
          mbar_wait(&bar[b], (phbits >> b) & 1u);
          phbits ^= 1u << b;
        }
        // warp-uniform steps at which lane tstar has just finished the last row of half 1 / half 2
        const int cs1 = real ? meta.x - 1 + tstar : -1;
        const int cs2 = (real && meta.w >= 0) ? meta.y - 1 + tstar : -1;
        // glocal: lane t is on the last row of half h at step ev_h + t
        const int ev1 = (GL && real) ? meta.x - 1 : NOEV;
        const int ev2 = (GL && real && meta.w >= 0) ? meta.y - 1 : NOEV;
        tb_off += prev_rows;
        // global: this pair's level; pair 0 of the item, every lv_n-th pair and the draining dummy pair are reset explicitly
        lv_i = (i == 0 || !real || lv_i + 1 >= lv_n) ? 0 : lv_i + 1;
        const bool inj = !GL && lv_i != 0;
        const bool inj_next = !GL && real && p_next < pair_hi && lv_i + 1 < lv_n;
        const int lvl = GL ? 0 : p.lv0 + lv_i * p.lv_step;
        if (!GL) {
          const uint32_t lw = b16::addend(lvl, lvl);
          Hc = Hc0 + lw; Ec = Ec0 + lw; H00 = H000 + lw;
          if (!inj) border_set(Hc, Ec);              // an explicitly reset pair: lane 0 reads this level's constants from its step 0 on
        }
        const int dr1 = ge * (lq - 1 + meta.x - 1) - lvl, dr2 = ge * (lq - 1 + meta.y - 1) - lvl;   // undo drift and level at the corner
        const int par = i & 1;
        int p_nn = pair_hi;
        auto issue_next_copy = [&]() {       // every lane has left pair i-1: its buffer can take pair i+2
          __syncwarp();
          if (p_next < pair_hi) p_nn = take_pair(p_next + W);
          if (lane == 0 && p_nn < pair_hi) {
            const int b2 = (i + 2) % 3;
            bulk_load(smem + tbuf0 + b2 * p.tp_stride, p.tp_codes + (long long)p_nn * p.tp_stride, p.tp_stride, &bar[b2]);
          }
        };
        if (DEFER) {
          // glocal with deferred last rows.  The last target row of a half is a forced diagonal (glocal_align.c:172-183):
          // its cells are what S[] holds after the row's first pass.  Lane t passes that row at step ev + t, a different step
          // for every lane, so the lane only parks S[] in shared memory there (predicated vector stores, no ALU work) and the
          // warp folds all parked rows once, at the warp-uniform capture step.  A lane overwrites its slot when it passes
          // the NEXT pair's last row, at step >= 31 + t of that pair, while the capture of this pair falls on one of its
          // first 31 steps: no target shorter than 32 is allowed on this path (align_abi.cu, job_prepare).
          uint32_t *evw = reinterpret_cast<uint32_t *>(smem + L.evs) + warp * (2 * 32 * KP);
          auto park = [&](int half) {
            uint32_t *dst = evw + half * (32 * KP);
            if (V == 4) {
#pragma unroll
              for (int v = 0; v < KP / 4; ++v) reinterpret_cast<uint4 *>(dst)[v * 32 + lane] = make_uint4(S[4 * v], S[4 * v + 1], S[4 * v + 2], S[4 * v + 3]);
            } else {
#pragma unroll
              for (int v = 0; v < KP / 2; ++v) reinterpret_cast<uint2 *>(dst)[v * 32 + lane] = make_uint2(S[2 * v], S[2 * v + 1]);
            }
          };
          auto park_at = [&](int s, int e1, int e2) {
            if (lane == s - e1) park(0);
            if (lane == s - e2) park(1);
          };
          // max over the real query columns c of (S~[c] + g*(c + r)), r = the half's last row: per lane the chain
          // n_k = max(n_{k-1} - g, S~[k]) (one add-max per column); the drift is undone in two non-negative steps (no borrow
          // between the halves whatever the other half holds): + |g|*(32K - c) per lane before the warp maximum, g*(32K + r) after
          auto fold = [&](int half, int idx, int r) {
            uint32_t m = NEGb;
            if (lane <= tstar) {
              uint32_t T[KP];
              const uint32_t *src = evw + half * (32 * KP);
              if (V == 4) {
#pragma unroll
                for (int v = 0; v < KP / 4; ++v) { const uint4 x = reinterpret_cast<const uint4 *>(src)[v * 32 + lane]; T[4 * v] = x.x; T[4 * v + 1] = x.y; T[4 * v + 2] = x.z; T[4 * v + 3] = x.w; }
              } else {
#pragma unroll
                for (int v = 0; v < KP / 2; ++v) { const uint2 x = reinterpret_cast<const uint2 *>(src)[v * 32 + lane]; T[2 * v] = x.x; T[2 * v + 1] = x.y; }
              }
              const int klast = lane < tstar ? K - 1 : kstar;
#pragma unroll
              for (int k = 0; k < K; ++k)
                if (k <= klast) m = p16::addmax(m, negg, T[k]);
              const int up = -ge * (P - c0 - klast);
              m += p16::pack(up, up);
            }
            const uint32_t v = warp_max2(m);
            const int sc = max(0, (half ? b16::hi(v) : b16::lo(v)) + ge * (P + r));     // max starts at 0 (glocal_align.c:115)
            if (sc > wbest || (sc == wbest && idx < widx)) { wbest = sc; widx = idx; }
            if (p.all_scores && lane == 0) p.all_scores[(long long)qi * p.n_t + idx] = sc;
          };
          auto captures = [&](int done) {
            if (done < MM16G_MIN_ROWS) {
              if (done == pend_s1) fold(0, pend_i1, pend_r1);
              if (done == pend_s2) fold(1, pend_i2, pend_r2);
            }
            if (done == cs1) fold(0, meta.z, meta.x - 1);
            if (done == cs2) fold(1, meta.w, meta.y - 1);
          };
          const bool early = min(ev1, ev2) < MM16G_MIN_ROWS;         // this pair's own last rows start inside the switching steps
          const int s_chk = min(rows, max(MM16G_MIN_ROWS, min(ev1, ev2)));      // from here on lanes reach this pair's last rows
          int s = 0;
          while (s < rows) {
            const bool switching = s < MM16G_MIN_ROWS;
            int stop = switching ? MM16G_MIN_ROWS : (s < s_chk ? s_chk : rows);
            if (switching) {
              if (pend_s1 >= s && pend_s1 < stop) stop = pend_s1 + 1;
              if (pend_s2 >= s && pend_s2 < stop) stop = pend_s2 + 1;
            }
            if (cs1 >= s && cs1 < stop) stop = cs1 + 1;
            if (cs2 >= s && cs2 < stop) stop = cs2 + 1;
            if (switching) {
              // lane s switches to this pair at step s and computes its row 0 through the common row (dp_core.cuh,
              // MM16G_NO_EGAP); the lanes still on the previous pair park its last rows
#pragma unroll kGlSwitchUnroll
              for (; s < stop; ++s) {
                uint32_t h_sh, e_sh, gope = gop2;
                exchange(h_sh, e_sh);
                if (lane == s) {
#pragma unroll
                  for (int k = 0; k < K; ++k) { Hp[k] = G0 - (uint32_t)k * gew; F[k] = NEGb; }
                  hleft = G0 + gew;
                  tb_off = bufoff - lane;
                  e_sh = NEGb; gope = NOEGAP;
                }
                row_e(s, h_sh, e_sh, gope);
                park_at(s, pend_e1, pend_e2);
                if (early) park_at(s, ev1, ev2);
              }
            } else if (s < s_chk) {
#pragma unroll kHotUnroll
              for (; s < stop; ++s) body(s);                    // no lane is on a last row yet
            } else {
              for (; s < stop; ++s) {
                body(s);
                park_at(s, ev1, ev2);
              }
            }
            captures(s - 1);
            if (s == MM16G_MIN_ROWS) issue_next_copy();
          }
        } else if (GL) {
          // undo of the drift on the previous pair's last rows, per half: + g*(c0 + r)
          const uint32_t dpend = b16::addend(ge * (c0 + K - 1 + pend_r1), ge * (c0 + K - 1 + pend_r2));
          const bool early = min(ev1, ev2) < 2 * MM16G_MIN_ROWS;     // a short target: its last rows start inside the switching steps
          // Segments end at the warp-uniform capture steps, so the loops carry no capture tests and unroll (as in global).
          // Steps 0..31: lane s switches to this pair at step s and computes its row 0 through the common row
          // (dp_core.cuh, MM16G_NO_EGAP); the lanes still on the previous pair fold its last rows.
          int s = 0;
          const int s_chk = min(rows, max(MM16G_MIN_ROWS, min(ev1, ev2)));      // from here on lanes reach this pair's last rows
          while (s < rows) {
            const bool switching = s < MM16G_MIN_ROWS;
            int stop = switching ? MM16G_MIN_ROWS : (s < s_chk ? s_chk : rows);
            if (switching) {
              if (pend_s1 >= s && pend_s1 < stop) stop = pend_s1 + 1;
              if (pend_s2 >= s && pend_s2 < stop) stop = pend_s2 + 1;
            }
            if (cs1 >= s && cs1 < stop) stop = cs1 + 1;
            if (cs2 >= s && cs2 < stop) stop = cs2 + 1;
            if (switching) {
#pragma unroll 1
              for (; s < stop; ++s) {
                const bool sw = lane == s;
                if (sw) {
#pragma unroll
                  for (int k = 0; k < K; ++k) { Hp[k] = G0 - (uint32_t)k * gew; F[k] = NEGb; }
                  hleft = G0 + gew;
                  tb_off = bufoff - lane;
                }
                uint32_t h_sh, e_sh;
                exchange(h_sh, e_sh);
                if (sw) e_sh = NEGb;
                row_e(s, h_sh, e_sh, sw ? NOEGAP : gop2);
                // last rows of the PREVIOUS pair: lane t is on them at steps pend_e1 + t / pend_e2 + t, all inside these 32
                // steps.  Every lane that owns only real columns folds the forced diagonal of the row it is on
                // (glocal_align.c:172-183: what S[] holds after the row's first pass) and keeps it if the row is a last
                // row; the one lane with padding columns takes the masked path.
                const bool h1 = lane == s - pend_e1, h2 = lane == s - pend_e2;
                if (lane < tstar) {
                  uint32_t m = NEGb;
#pragma unroll
                  for (int k = 0; k < K; ++k) m = p16::addmax(m, negg, S[k]);      // n_k = max(n_{k-1} - g, S[k])
                  m += dpend;                                                         // + g*(c0 + K-1 + r) per half
                  const uint32_t mm = (h1 ? (m & 0xffffu) : (NEGb & 0xffffu)) | (h2 ? (m & 0xffff0000u) : (NEGb & 0xffff0000u));
                  if (par) lm0 = p16::max2(lm0, mm); else lm1 = p16::max2(lm1, mm);       // parity of the previous pair
                } else if (h1 || h2) {
                  last_row(h1, h2, h1 ? pend_r1 : pend_r2, par ^ 1);
                }
                if (early) last_rows_at(s, ev1, ev2, meta.x - 1, meta.y - 1, par);
              }
            } else if (s < s_chk) {
#pragma unroll kHotUnroll
              for (; s < stop; ++s) body(s);                    // no lane is on a last row yet
            } else {
              for (; s < stop; ++s) {
                body(s);
                last_rows_at(s, ev1, ev2, meta.x - 1, meta.y - 1, par);
              }
            }
            const int done = s - 1;
            if (done < MM16G_MIN_ROWS) {
              if (done == pend_s1) capture(0, pend_i1, pend_d1, par ^ 1);
              if (done == pend_s2) capture(1, pend_i2, pend_d2, par ^ 1);
            }
            if (done == cs1) capture(0, meta.z, dr1, par);
            if (done == cs2) capture(1, meta.w, dr2, par);
            if (s == MM16G_MIN_ROWS) issue_next_copy();
          }
        } else {
          // The pair's steps run in segments that end where a corner cell has to be captured (a warp-uniform
          // step), so that the loops themselves carry no capture tests and unroll; in the first 32 steps lane
          // s switches to this pair at step s: by an explicit reset of its state to the global boundary
          // (global_align.c:136-148), or, on an injected pair, by moving its row pointer only: the separator row
          // it has just swept left the boundary state of this pair's level in its registers (mm16g_levels).
          auto captures = [&](int done) {
            if (done < MM16G_MIN_ROWS) {
              if (done == pend_s1) capture(0, pend_i1, pend_d1, par ^ 1);
              if (done == pend_s2) capture(1, pend_i2, pend_d2, par ^ 1);
            }
            if (done == cs1) capture(0, meta.z, dr1, par);
            if (done == cs2) capture(1, meta.w, dr2, par);
          };
          int s = 0;
          while (s < rows) {
            const bool switching = s < MM16G_MIN_ROWS;
            int stop = switching ? MM16G_MIN_ROWS : rows;
            if (switching) {
              if (pend_s1 >= s && pend_s1 < stop) stop = pend_s1 + 1;
              if (pend_s2 >= s && pend_s2 < stop) stop = pend_s2 + 1;
            }
            if (cs1 >= s && cs1 < stop) stop = cs1 + 1;
            if (cs2 >= s && cs2 < stop) stop = cs2 + 1;
            if (switching && inj) {
#pragma unroll kInjUnroll
              for (; s < stop; ++s) {
                if (lane == s) tb_off = bufoff - lane;
                body(s);
              }
            } else if (switching) {
#pragma unroll kSwitchUnroll
              for (; s < stop; ++s) {
                if (lane == s) {
#pragma unroll
                  for (int k = 0; k < K; ++k) { Hp[k] = Hc; F[k] = Ec; }
                  hleft = lane ? Hc : H00;
                  tb_off = bufoff - lane;
                }
                body(s);
              }
            } else {
#pragma unroll kHotUnroll
              for (; s < stop; ++s) body(s);
            }
            captures(s - 1);
            if (s == MM16G_MIN_ROWS) issue_next_copy();
          }
          if (inj_next) {
            // the separator row (the pad row after the pair's last row): lane 0 takes E~ = Hc of the next level from the
            // left, and as its diagonal for the next pair's row 0 that level's H~[-1][-1]
            const uint32_t up = b16::addend(p.lv_step, p.lv_step);
            uint32_t h_sh, e_sh;
            if (BORDER) {
              if (lane == BL) { h_out = H00 + up; e_out = Hc + up; }
              exchange(h_sh, e_sh);
            } else {
              h_sh = __shfl_up_sync(ABV_FULL, h_out, 1);
              e_sh = __shfl_up_sync(ABV_FULL, e_out, 1);
              if (lane == 0) { h_sh = H00 + up; e_sh = Hc + up; }
            }
            row(s, h_sh, e_sh);
            border_set(Hc + up, Ec + up);              // from the next pair's step 0 on lane 0 reads the next level's constants
            captures(s);
            rows += 1;
          }
        }
        pend_s1 = cs1 >= rows ? cs1 - rows : -1; pend_i1 = meta.z; pend_d1 = dr1;
        pend_s2 = cs2 >= rows ? cs2 - rows : -1; pend_i2 = meta.w; pend_d2 = dr2;
        pend_e1 = ev1 == NOEV ? NOEV : ev1 - rows; pend_r1 = meta.x - 1;
        pend_e2 = ev2 == NOEV ? NOEV : ev2 - rows; pend_r2 = meta.y - 1;
        prev_rows = rows;
        if (!real) break;
        p_cur = p_next; p_next = p_nn;
      }
#if defined(ABV_DUMMY_FMA) || defined(ABV_DUMMY_ALU)
      if (dummy[0] + dummy[1] + dummy[2] + dummy[3] == 0x12345u) wbest = 0;      // keeps the experiment's chains alive
#endif
    }

    if (lane == (GL ? 0 : tstar)) { wres[2 * warp] = wbest; wres[2 * warp + 1] = widx; }
    __syncthreads();
    if (threadIdx.x == 0) {
      int bs = INT_MIN, bi = INT_MAX;
      for (int w = 0; w < W; ++w) {
        const int s = wres[2 * w], i = wres[2 * w + 1];
        if (i != INT_MAX && (s > bs || (s == bs && i < bi))) { bs = s; bi = i; }
      }
      if (bi != INT_MAX) atomicMax(p.best_key + qi, best_key_pack(bs, bi));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Integer-pipe microbenchmark: register-only dependent chains, UNROLL independent chains per thread.
template <int WHICH>
__global__ void int_peak_kernel(uint32_t *out, int iters, uint32_t seed) {
  constexpr int CH = 8;
  uint32_t a[CH], b[CH], c[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    a[i] = seed + threadIdx.x * 7 + i;
    b[i] = seed * 3 + i * 5 + (threadIdx.x & 3);
    c[i] = seed ^ (i * 0x10001u);
  }
  uint32_t gopr = 0xfffafffau;
  asm volatile("" : "+r"(gopr));      // opaque: stays in a register
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (WHICH == 0) { a[i] = a[i] + b[i]; b[i] = b[i] + c[i]; c[i] = c[i] + a[i]; a[i] += 3; }                   // 4 IADD
      else if (WHICH == 1) { a[i] = __vimax3_s16x2(a[i], b[i], c[i]); b[i] = __vimax3_s16x2(b[i], c[i], a[i]);
                             c[i] = __vimax3_s16x2(c[i], a[i], b[i] ^ 1); a[i] = __vimax3_s16x2(a[i], b[i], 0x00010001u); }
      else if (WHICH == 2) { a[i] = __viaddmax_s16x2(a[i], b[i], c[i]); b[i] = __viaddmax_s16x2(b[i], c[i], a[i]);
                             c[i] = __viaddmax_s16x2(c[i], a[i], b[i]); a[i] = __viaddmax_s16x2(a[i], 0x00010001u, b[i]); }
      else if (WHICH == 3) { a[i] = __vadd2(a[i], b[i]); b[i] = __vadd2(b[i], c[i]); c[i] = __vadd2(c[i], a[i]); a[i] = __vadd2(a[i], 0x00030003u); }
      else if (WHICH == 4) {   // the packed DP cell: 2 add + max3 + 2 addmax   (5 ops)
        uint32_t dd = __vadd2(a[i], c[i]);
        uint32_t h = __vimax3_s16x2(dd, b[i], c[i]);
        uint32_t hg = __vadd2(h, 0xfff9fff9u);
        b[i] = __viaddmax_s16x2(b[i], 0xffffffffu, hg);
        c[i] = __viaddmax_s16x2(c[i], 0xffffffffu, hg);
        a[i] = h;
      } else if (WHICH == 5) { // the 32-bit DP cell: 2 add + max3 + 2 addmax   (5 ops)
        int dd = (int)a[i] + (int)c[i];
        int h = __vimax3_s32(dd, (int)b[i], (int)c[i]);
        int hg = h - 7;
        b[i] = (uint32_t)__viaddmax_s32((int)b[i], -1, hg);
        c[i] = (uint32_t)__viaddmax_s32((int)c[i], -1, hg);
        a[i] = (uint32_t)h;
      } else if (WHICH == 6) { a[i] = (uint32_t)__vimax3_s32((int)a[i], (int)b[i], (int)c[i]); b[i] = (uint32_t)__vimax3_s32((int)b[i], (int)c[i], (int)a[i]);
                               c[i] = (uint32_t)__vimax3_s32((int)c[i], (int)a[i], (int)(b[i] ^ 1)); a[i] = (uint32_t)__vimax3_s32((int)a[i], (int)b[i], 65537); }
      else if (WHICH == 8) {   // the drift-normalised packed cell of mm16g_kernel: add + max3 + 2 add-max (4 ops)
        uint32_t dd = a[i] + c[i];
        uint32_t h = __vimax3_s16x2(dd, b[i], c[i]);
        b[i] = __viaddmax_s16x2(h, 0xfffafffau, b[i]);
        c[i] = __viaddmax_s16x2(h, 0xfffafffau, c[i]);
        a[i] = h;
      } else if (WHICH == 9) {   // the same cell with the gap addend in a REGISTER, as the kernel has it (WHICH 8: an immediate)
        uint32_t dd = a[i] + c[i];
        uint32_t h = __vimax3_s16x2(dd, b[i], c[i]);
        b[i] = __viaddmax_s16x2(h, gopr, b[i]);
        c[i] = __viaddmax_s16x2(h, gopr, c[i]);
        a[i] = h;
      } else if (WHICH == 10) {  // add-max with three register sources
        a[i] = __viaddmax_s16x2(a[i], gopr, c[i]); b[i] = __viaddmax_s16x2(b[i], gopr, a[i]);
        c[i] = __viaddmax_s16x2(c[i], gopr, b[i]); a[i] = __viaddmax_s16x2(a[i], gopr, b[i]);
      } else if (WHICH == 11) {  // two-input packed max
        a[i] = __vmaxs2(a[i], b[i]); b[i] = __vmaxs2(b[i], c[i]); c[i] = __vmaxs2(c[i], a[i] ^ 1); a[i] = __vmaxs2(a[i], 0x00010001u);
      } else {                 // packed cell with plain 32-bit adds (biased-operand variant): 5 ops
        uint32_t dd = a[i] + c[i];
        uint32_t h = __vimax3_s16x2(dd, b[i], c[i]);
        uint32_t hg = h + 0xfff8fff9u;
        b[i] = __viaddmax_s16x2(b[i], 0xffffffffu, hg);
        c[i] = __viaddmax_s16x2(c[i], 0xffffffffu, hg);
        a[i] = h;
      }
    }
  }
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) r ^= a[i] ^ b[i] ^ c[i];
  if (r == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = r;   // keeps the chains live
}
constexpr int INT_PEAK_OPS[12] = {4, 4, 4, 4, 5, 5, 4, 5, 4, 4, 4, 4};   // lane-ops per chain per iteration
constexpr int INT_PEAK_CHAINS = 8;

}  // namespace abv
