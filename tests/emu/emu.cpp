// tests/emu/emu.cpp -- TEST INFRASTRUCTURE ONLY.
// Drives the device cell logic of ambivert_b200/csrc/dp_core.cuh (the very functions the sm_100a
// kernels call) lane by lane on the CPU, mirroring the warp loops of kernels.cuh, so that the
// recurrences, tie rules, direction nibbles, traceback walk and shared-memory layouts can be
// checked against the oracle in the CPU test suite.  Never linked into the product library.
#include <algorithm>
#include <cstring>
#include <vector>

#include "../../ambivert_b200/csrc/dp_core.cuh"

using namespace abv;

template <int MODE>
static int emu_wf32_t(const unsigned char *sa, int la, const unsigned char *sb, int lb, int alpha, const int *M,
                      int go, int ge, int *score_out, FragOut *frags, int cap) {
  const int nblk = (la + 31) >> 5, steps8 = dir_steps8(lb);
  std::vector<uint32_t> dir((size_t)dir_words(la, lb), 0xdeadbeefu);
  std::vector<int> borderH(lb), borderE(lb);
  EndCell wbest = wf32_initial_candidate<MODE>();
  int corner = 0;
  for (int b = 0; b < nblk; ++b) {
    int F[32] = {0}, hleft_prev[32] = {0}, h_out[32] = {0}, e_out[32] = {0};
    uint32_t acc[32] = {0};
    EndCell cand[32];
    for (int t = 0; t < 32; ++t) cand[t] = wf32_initial_candidate<MODE>();
    const int lanes_used = std::min(32, la - (b << 5));
    const int nsteps = lb + lanes_used - 1;
    for (int s = 0; s < nsteps; ++s) {
      int h_sh[32], e_sh[32];
      for (int t = 0; t < 32; ++t) { h_sh[t] = t ? h_out[t - 1] : h_out[0]; e_sh[t] = t ? e_out[t - 1] : e_out[0]; }
      std::vector<std::pair<int, std::pair<int, int>>> border_writes;
      for (int t = 0; t < 32; ++t) {
        const int c = (b << 5) + t, r = s - t;
        const bool colvalid = c < la;
        if (t == 0 && s < lb) {
          if (b == 0) wf32_virtual_column<MODE>(s, go, ge, h_sh[0], e_sh[0]);
          else { h_sh[0] = borderH[s]; e_sh[0] = borderE[s]; }
        }
        unsigned nib = 0;
        if (r >= 0 && r < lb && colvalid) {
          const int sub = M[(int)sa[c] * alpha + (int)sb[r]];
          int h, eo;
          nib = wf32_cell<MODE>(c, r, la, lb, sub, hleft_prev[t], e_sh[t], F[t], go, ge, h, eo, cand[t]);
          hleft_prev[t] = h_sh[t];
          h_out[t] = h; e_out[t] = eo;
          if (MODE == GLOBAL && c == la - 1 && r == lb - 1) corner = h;
          if (t == 31 && b + 1 < nblk) border_writes.push_back({r, {h, eo}});
        }
        acc[t] |= nib << ((s & 7) * 4);
        if ((s & 7) == 7 || s == nsteps - 1) { dir[((long long)b * steps8 + (s >> 3)) * 32 + t] = acc[t]; acc[t] = 0; }
      }
      for (auto &w : border_writes) { borderH[w.first] = w.second.first; borderE[w.first] = w.second.second; }
    }
    EndCell best = cand[0];
    // same xor-butterfly order as the kernel
    EndCell cur[32];
    for (int t = 0; t < 32; ++t) cur[t] = cand[t];
    for (int off = 16; off > 0; off >>= 1) {
      EndCell nxt[32];
      for (int t = 0; t < 32; ++t) { EndCell o = cur[t ^ off]; nxt[t] = end_better(o, cur[t]) ? o : cur[t]; }
      for (int t = 0; t < 32; ++t) cur[t] = nxt[t];
    }
    best = cur[0];
    if (best.col >= 0 && best.score >= wbest.score) wbest = best;
  }
  int score, ec, er;
  wf32_finish<MODE>(wbest, corner, la, lb, score, ec, er);
  *score_out = score;
  DirReader rd; rd.w = dir.data(); rd.steps8 = steps8;
  std::vector<FragOut> tmp(la + lb + 2);
  int n = walk_back<MODE>(rd, ec, er, tmp.data(), (int)tmp.size());
  for (int i = 0; i < n && i < cap; ++i) frags[i] = tmp[n - 1 - i];
  return n;
}

// ---- the packed many-to-many kernel, one (query, target pair) ---------------------------------
template <int K>
static void load_S(const uint32_t *lut_row, int lane, uint32_t (&S)[K]) {
  constexpr int V = (K % 4 == 0) ? 4 : 2;
  for (int i = 0; i < K / V; ++i)
    for (int j = 0; j < V; ++j) S[i * V + j] = lut_row[(i * 32 + lane) * V + j];
}

template <int MODE, int K>
static void emu_mm16_t(const unsigned char *q, int lq, const unsigned char *t1, int l1, const unsigned char *t2, int l2,
                       int alpha, const int *M, int go, int ge, int *sc1_out, int *sc2_out) {
  constexpr int V = (K % 4 == 0) ? 4 : 2, P = 32 * K;
  const int a1 = alpha + 1, ncodes = a1 * a1;
  std::vector<int> mext((size_t)a1 * a1, 0);
  for (int x = 0; x < alpha; ++x) for (int y = 0; y < alpha; ++y) mext[x * a1 + y] = M[x * alpha + y];
  std::vector<unsigned char> qcodes(P);
  for (int c = 0; c < P; ++c) qcodes[c] = c < lq ? q[c] : (unsigned char)(a1 - 1);
  std::vector<uint32_t> lut((size_t)ncodes * P);
  for (int idx = 0; idx < ncodes * P; ++idx) {
    const int code = idx / P, pos = idx - code * P;
    const int j = pos % V, it = pos / V, t = it & 31, i = it >> 5;
    const int c = t * K + i * V + j;
    const int qa = qcodes[c] * a1;
    const int a = code / a1, b = code - a * a1;
    lut[idx] = p16::pack(mext[qa + a], mext[qa + b]);
  }
  const int lmax = std::max(l1, l2);
  std::vector<unsigned char> tb(lmax + 16);
  for (int r = 0; r < (int)tb.size(); ++r) {
    const int a = r < l1 ? t1[r] : alpha, b = r < l2 ? t2[r] : alpha;
    tb[r] = (unsigned char)(a * a1 + b);
  }
  const uint32_t go2 = p16::pack(go, go), ge2 = p16::pack(ge, ge), NEG2 = p16::pack(p16::NEG16, p16::NEG16);
  const int tstar = (lq - 1) / K, kstar = (lq - 1) - tstar * K;
  const int L1m1 = l1 - 1, L2m1 = l2 - 1;
  uint32_t Hp[32][K], F[32][K], S[K];
  uint32_t hleft_prev[32], h_out[32] = {0}, e_out[32] = {0}, best2[32] = {0};
  int cap1 = 0, cap2 = 0, sc1 = 0, sc2 = 0;
  if (MODE == GLOBAL) {
    for (int t = 0; t < 32; ++t) {
      const int c0 = t * K;
      for (int k = 0; k < K; ++k) { const int v = go + ge * (c0 + k); Hp[t][k] = p16::pack(v, v); F[t][k] = p16::pack(v + go, v + go); }
      const int v = c0 ? go + ge * (c0 - 1) : 0; hleft_prev[t] = p16::pack(v, v);
    }
    int vcol = go;
    const int nsteps = lmax + tstar;
    for (int s = 0; s < nsteps; ++s) {
      uint32_t h_sh[32], e_sh[32];
      for (int t = 0; t < 32; ++t) { h_sh[t] = t ? h_out[t - 1] : h_out[0]; e_sh[t] = t ? e_out[t - 1] : e_out[0]; }
      h_sh[0] = p16::pack(vcol, vcol); e_sh[0] = p16::pack(vcol + go, vcol + go); vcol += ge;
      for (int t = 0; t < 32; ++t) {
        const int r = s - t;
        if (r >= 0 && r < lmax && t <= tstar) {
          load_S<K>(&lut[(size_t)tb[r] * P], t, S);
          uint32_t e = e_sh[t], dummy = 0;
          mm16_row<K, false, false>(hleft_prev[t], e, Hp[t], F[t], S, go2, ge2, dummy);
          hleft_prev[t] = h_sh[t]; e_out[t] = e; h_out[t] = Hp[t][K - 1];
          if (t == tstar && (r == L1m1 || r == L2m1)) {
            const uint32_t hs = Hp[t][kstar];
            if (r == L1m1) cap1 = p16::lo(hs);
            if (r == L2m1) cap2 = p16::hi(hs);
          }
        }
      }
    }
    sc1 = cap1; sc2 = cap2;
  } else if (MODE == LOCAL) {
    for (int t = 0; t < 32; ++t) { for (int k = 0; k < K; ++k) { Hp[t][k] = 0; F[t][k] = go2; } hleft_prev[t] = 0; }
    const int nsteps = lmax + tstar;
    for (int s = 0; s < nsteps; ++s) {
      uint32_t h_sh[32], e_sh[32];
      for (int t = 0; t < 32; ++t) { h_sh[t] = t ? h_out[t - 1] : h_out[0]; e_sh[t] = t ? e_out[t - 1] : e_out[0]; }
      h_sh[0] = 0; e_sh[0] = go2;
      for (int t = 0; t < 32; ++t) {
        const int r = s - t;
        if (r >= 0 && r < lmax && t <= tstar) {
          load_S<K>(&lut[(size_t)tb[r] * P], t, S);
          uint32_t e = e_sh[t];
          mm16_row<K, true, true>(hleft_prev[t], e, Hp[t], F[t], S, go2, ge2, best2[t]);
          hleft_prev[t] = h_sh[t]; e_out[t] = e; h_out[t] = Hp[t][K - 1];
        }
      }
    }
    uint32_t b = 0;
    for (int t = 0; t < 32; ++t) b = p16::max2(b, best2[t]);
    sc1 = p16::lo(b); sc2 = p16::hi(b);
  } else if (MODE == OVERLAP) {      // mm16_kernel<OVERLAP>: forced first row and first column, candidates on the last column / last rows
    for (int t = 0; t < 32; ++t) {
      load_S<K>(&lut[(size_t)tb[0] * P], t, S);
      for (int k = 0; k < K; ++k) { Hp[t][k] = S[k]; F[t][k] = p16::add(S[k], go2); }
      best2[t] = NEG2;
    }
    for (int t = 31; t > 0; --t) hleft_prev[t] = Hp[t - 1][K - 1];
    hleft_prev[0] = 0;
    for (int t = 0; t <= tstar; ++t)
      mm16_overlap_cands<K>(Hp[t], t * K, lq, t == tstar ? kstar : -1, 0xffffffffu, (L1m1 == 0 ? 0x0000ffffu : 0u) | (L2m1 == 0 ? 0xffff0000u : 0u), best2[t]);
    const int nsteps = lmax - 1 + tstar;
    for (int s = 0; s < nsteps; ++s) {
      uint32_t h_sh[32], e_sh[32];
      for (int t = 0; t < 32; ++t) { h_sh[t] = t ? h_out[t - 1] : h_out[0]; e_sh[t] = t ? e_out[t - 1] : e_out[0]; }
      h_sh[0] = 0; e_sh[0] = NEG2;
      for (int t = 0; t < 32; ++t) {
        const int r = 1 + s - t;
        if (r >= 1 && r < lmax && t <= tstar) {
          load_S<K>(&lut[(size_t)tb[r] * P], t, S);
          if (t == 0) F[t][0] = NEG2;
          uint32_t e = e_sh[t], dummy = 0;
          mm16_row<K, false, false>(hleft_prev[t], e, Hp[t], F[t], S, go2, ge2, dummy);
          hleft_prev[t] = h_sh[t]; e_out[t] = e; h_out[t] = Hp[t][K - 1];
          const bool a = (r == L1m1), b = (r == L2m1);
          if (t == tstar || a || b)
            mm16_overlap_cands<K>(Hp[t], t * K, lq, t == tstar ? kstar : -1, (r <= L1m1 ? 0x0000ffffu : 0u) | (r <= L2m1 ? 0xffff0000u : 0u),
                                  (a ? 0x0000ffffu : 0u) | (b ? 0xffff0000u : 0u), best2[t]);
        }
      }
    }
    uint32_t b = NEG2;
    for (int t = 0; t < 32; ++t) b = p16::max2(b, best2[t]);
    sc1 = p16::lo(b); sc2 = p16::hi(b);
  } else {
    for (int t = 0; t < 32; ++t) {
      load_S<K>(&lut[(size_t)tb[0] * P], t, S);
      for (int k = 0; k < K; ++k) { Hp[t][k] = S[k]; F[t][k] = p16::add(S[k], go2); }
      best2[t] = NEG2;
    }
    for (int t = 31; t > 0; --t) hleft_prev[t] = Hp[t - 1][K - 1];
    hleft_prev[0] = NEG2;
    const int nsteps = lmax - 1 + tstar;
    for (int s = 0; s < nsteps; ++s) {
      uint32_t h_sh[32], e_sh[32];
      for (int t = 0; t < 32; ++t) { h_sh[t] = t ? h_out[t - 1] : h_out[0]; e_sh[t] = t ? e_out[t - 1] : e_out[0]; }
      h_sh[0] = NEG2; e_sh[0] = NEG2;
      for (int t = 0; t < 32; ++t) {
        const int r = 1 + s - t;
        if (r >= 1 && r < lmax && t <= tstar) {
          load_S<K>(&lut[(size_t)tb[r] * P], t, S);
          uint32_t e = e_sh[t];
          const bool a = (r == L1m1), b = (r == L2m1);
          if (a || b) {
            const uint32_t mask = (a ? 0x0000ffffu : 0u) | (b ? 0xffff0000u : 0u);
            mm16_row_glocal_last<K>(hleft_prev[t], e, Hp[t], F[t], S, go2, ge2, mask, t * K, lq, best2[t]);
          } else {
            uint32_t dummy = 0;
            mm16_row<K, false, false>(hleft_prev[t], e, Hp[t], F[t], S, go2, ge2, dummy);
          }
          hleft_prev[t] = h_sh[t]; e_out[t] = e; h_out[t] = Hp[t][K - 1];
        }
      }
    }
    uint32_t b = NEG2;
    for (int t = 0; t < 32; ++t) b = p16::max2(b, best2[t]);
    sc1 = std::max(0, p16::lo(b)); sc2 = std::max(0, p16::hi(b));
  }
  *sc1_out = sc1; *sc2_out = sc2;
}

template <int MODE>
static int emu_mm16_k(int K, const unsigned char *q, int lq, const unsigned char *t1, int l1, const unsigned char *t2, int l2,
                      int alpha, const int *M, int go, int ge, int *s1, int *s2) {
  switch (K) {
    case 2: emu_mm16_t<MODE, 2>(q, lq, t1, l1, t2, l2, alpha, M, go, ge, s1, s2); return 0;
    case 4: emu_mm16_t<MODE, 4>(q, lq, t1, l1, t2, l2, alpha, M, go, ge, s1, s2); return 0;
    case 6: emu_mm16_t<MODE, 6>(q, lq, t1, l1, t2, l2, alpha, M, go, ge, s1, s2); return 0;
    case 8: emu_mm16_t<MODE, 8>(q, lq, t1, l1, t2, l2, alpha, M, go, ge, s1, s2); return 0;
    case 10: emu_mm16_t<MODE, 10>(q, lq, t1, l1, t2, l2, alpha, M, go, ge, s1, s2); return 0;
    case 12: emu_mm16_t<MODE, 12>(q, lq, t1, l1, t2, l2, alpha, M, go, ge, s1, s2); return 0;
    case 16: emu_mm16_t<MODE, 16>(q, lq, t1, l1, t2, l2, alpha, M, go, ge, s1, s2); return 0;
  }
  return -1;
}

extern "C" {

int emu_wf32(int mode, const unsigned char *sa, int la, const unsigned char *sb, int lb, int alpha, const int *M,
             int go, int ge, int *score, int *frags /* 4 ints each */, int cap) {
  FragOut *f = reinterpret_cast<FragOut *>(frags);
  switch (mode) {
    case OVERLAP: return emu_wf32_t<OVERLAP>(sa, la, sb, lb, alpha, M, go, ge, score, f, cap);
    case LOCAL: return emu_wf32_t<LOCAL>(sa, la, sb, lb, alpha, M, go, ge, score, f, cap);
    case GLOBAL: return emu_wf32_t<GLOBAL>(sa, la, sb, lb, alpha, M, go, ge, score, f, cap);
    case GLOCAL: return emu_wf32_t<GLOCAL>(sa, la, sb, lb, alpha, M, go, ge, score, f, cap);
  }
  return -1;
}

int emu_mm16(int mode, int K, const unsigned char *q, int lq, const unsigned char *t1, int l1,
             const unsigned char *t2, int l2, int alpha, const int *M, int go, int ge, int *s1, int *s2) {
  if (32 * K < lq) return -1;
  switch (mode) {
    case LOCAL: return emu_mm16_k<LOCAL>(K, q, lq, t1, l1, t2, l2, alpha, M, go, ge, s1, s2);
    case GLOBAL: return emu_mm16_k<GLOBAL>(K, q, lq, t1, l1, t2, l2, alpha, M, go, ge, s1, s2);
    case GLOCAL: return emu_mm16_k<GLOCAL>(K, q, lq, t1, l1, t2, l2, alpha, M, go, ge, s1, s2);
    case OVERLAP: return (lq < 2 || go > ge) ? -1 : emu_mm16_k<OVERLAP>(K, q, lq, t1, l1, t2, l2, alpha, M, go, ge, s1, s2);   // as align_abi.cu: exact for go <= ge
  }
  return -1;
}
}

static int g_level_cap = 1 << 20;   // test hook: shorter level staircases, so that few pairs already wrap around
extern "C" void emu_set_level_cap(int cap) { g_level_cap = cap < 1 ? 1 : cap; }

// ---- the streaming kernel (mm16g_kernel<GLOBAL|GLOCAL>): ONE warp sweeping a list of target pairs ----
// Mirrors the control flow of the kernel: three row buffers, 32 switch steps per pair, pending corner
// captures, the zero-symbol prefix before the first pair and the dummy pair that drains the pipeline;
// for glocal the dependency-free row 0 at the switch step and the last-row maxima collected per lane.
template <bool GL, int K, bool BORDER>
static void emu_mm16g_t(const unsigned char *q, int lq, int n_pairs, const unsigned char *const *t1s, const int *l1s,
                        const unsigned char *const *t2s, const int *l2s, int alpha, const int *M, int go, int ge,
                        int *sc1_out, int *sc2_out) {
  constexpr int KP = StripCfg<K>::KP, V = StripCfg<K>::V, P = StripCfg<K>::P;
  static_assert(!BORDER || mm16g_border(K, GL), "no border lane for this strip width / mode");   // BORDER: lane 31 emits the constants of column -1
  constexpr int BL = MM16G_BORDER_LANE;
  const int a1 = alpha + 1, ncodes = a1 * a1;
  std::vector<int> mext((size_t)a1 * a1, 0);
  for (int x = 0; x < alpha; ++x) for (int y = 0; y < alpha; ++y) mext[x * a1 + y] = M[x * alpha + y];
  std::vector<unsigned char> qcodes(P);
  for (int c = 0; c < P; ++c) qcodes[c] = c < lq ? q[c] : (unsigned char)(a1 - 1);
  std::vector<uint32_t> lut((size_t)ncodes * P);
  for (int idx = 0; idx < ncodes * P; ++idx) {
    const int code = idx / P, pos = idx - code * P;
    const int j = pos % V, it = pos / V, t = it & 31, i = it >> 5;
    const int k = i * V + j;
    const int qa = (k < K ? (int)qcodes[t * K + k] : a1 - 1) * a1;
    const int a = code / a1, b = code - a * a1;
    lut[idx] = (BORDER && t == BL) ? 0u : b16::addend(mext[qa + a] - 2 * ge, mext[qa + b] - 2 * ge);
  }
  int max_lt = 1;
  for (int i = 0; i < n_pairs; ++i) max_lt = std::max(max_lt, std::max(l1s[i], l2s[i]));
  const int stride = (std::max(max_lt, MM16G_MIN_ROWS) + 1 + 15) & ~15;     // as align_abi.cu job_prepare: room for the separator row
  std::vector<unsigned char> sm(64 + 3 * stride, 0);
  const int zero = 0, tbuf0 = 64;
  auto fill = [&](int buf, int pi) {
    for (int r = 0; r < stride; ++r) {
      const int a = r < l1s[pi] ? t1s[pi][r] : alpha, b = r < l2s[pi] ? t2s[pi][r] : alpha;
      sm[tbuf0 + buf * stride + r] = (unsigned char)(a * a1 + b);
    }
  };
  const uint32_t gop2 = p16::pack(go - ge, go - ge), gopw = b16::addend(go - ge, go - ge), gew = b16::addend(ge, ge);
  const uint32_t NEGb = b16::biased(p16::NEG16, p16::NEG16);
  const uint32_t Hc0 = GL ? NEGb : b16::biased(go + ge, go + ge), Ec0 = GL ? NEGb : b16::biased(2 * go, 2 * go);
  const uint32_t H000 = b16::biased(2 * ge, 2 * ge);
  uint32_t Hc = Hc0, Ec = Ec0, H00 = H000;       // boundary constants of the current pair's level (global, mm16g_levels)
  int maxabs = std::max(std::abs(go), std::abs(ge)), maxpos = 0;
  for (int x = 0; x < alpha * alpha; ++x) { maxabs = std::max(maxabs, std::abs(M[x])); maxpos = std::max(maxpos, M[x]); }
  Mm16gLevels lv = GL ? Mm16gLevels{1, 0, 0} : mm16g_levels(P, stride, go, ge, std::max(maxabs, 1), maxpos);
  lv.n = std::min(lv.n, g_level_cap);
  bool explicit_reset = true, inject = false;    // how lanes enter the current pair; lane 0 on a separator row
  const int tstar = (lq - 1) / K, kstar = (lq - 1) - tstar * K;
  uint32_t Hp[32][K], F[32][K], S[32][KP], hleft[32], h_out[32], e_out[32];
  int tb_off[32], buf_lo[32], buf_hi[32];
  for (int t = 0; t < 32; ++t) { buf_lo[t] = zero; buf_hi[t] = zero + 64; }
  int slots[4] = {INT_MIN, INT_MIN, INT_MIN, INT_MIN};
  for (int t = 0; t < 32; ++t) {
    for (int k = 0; k < K; ++k) { Hp[t][k] = b16::biased(0, 0); F[t][k] = b16::biased(0, 0); }
    hleft[t] = h_out[t] = e_out[t] = b16::biased(0, 0);
    tb_off[t] = zero + 32 - t;
  }
  const int n_i = n_pairs;
  if (n_i > 0) fill(0, 0);
  if (n_i > 1) fill(1, 1);
  int prev_rows = 0, pend_s1 = -1, pend_s2 = -1, pend_i1 = -1, pend_i2 = -1;
  const int NOEV = 1 << 20;
  int pend_e1 = NOEV, pend_e2 = NOEV, pend_r1 = 0, pend_r2 = 0;
  // glocal, deferred last rows (kernels.cuh): a lane parks the diagonal sums of the last target row of a half at the step it
  // passes that row; the warp folds the parked rows at the capture step.  One slot per (half, lane): a capture must come before
  // the lane's next event overwrites it (targets of at least 32 rows, align_abi.cu job_prepare)
  std::vector<uint32_t> evs((size_t)2 * 32 * KP, 0);
  const uint32_t negg = p16::pack(-ge, -ge);
  auto last_row = [&](int t, int half, int, int) {
    for (int k = 0; k < KP; ++k) evs[((size_t)half * 32 + t) * KP + k] = S[t][k];
  };
  auto fold = [&](int half, int r) {
    uint32_t best = NEGb;
    for (int t = 0; t <= tstar; ++t) {
      const int klast = t < tstar ? K - 1 : kstar;
      uint32_t m = NEGb;
      for (int k = 0; k <= klast; ++k) m = p16::addmax(m, negg, evs[((size_t)half * 32 + t) * KP + k]);
      const int up = -ge * (P - t * K - klast);
      m += p16::pack(up, up);
      best = p16::max2(best, m);
    }
    return std::max(0, (half ? b16::hi(best) : b16::lo(best)) + ge * (P + r));
  };
  // one step of the warp; `sw` = lane that switches pair in this step (-1: none), with its new buffer;
  // ev*/pe*: glocal last-row events of the current / previous pair (lane t fires at step ev + t)
  auto step = [&](int s, int sw, int bufoff, int ev1, int r1, int ev2, int r2, bool with_pending, int par) {
    if (sw >= 0) {
      const int t = sw;
      tb_off[t] = bufoff - t;
      buf_lo[t] = bufoff >= tbuf0 ? bufoff : zero; buf_hi[t] = bufoff >= tbuf0 ? bufoff + stride : zero + 64;
      if (!GL && explicit_reset) {
        for (int k = 0; k < K; ++k) { Hp[t][k] = Hc; F[t][k] = Ec; }
        hleft[t] = t ? Hc : H00;
      }
    }
    uint32_t h_sh[32], e_sh[32];
    if (BORDER) {      // shuffle by index: lane 0 reads the border lane, the border lane reads itself
      if (inject) { const uint32_t up = b16::addend(lv.step, lv.step); h_out[BL] = H00 + up; e_out[BL] = Hc + up; }
      for (int t = 0; t < 32; ++t) { const int src = t == 0 ? BL : (t == BL ? BL : t - 1); h_sh[t] = h_out[src]; e_sh[t] = e_out[src]; }
    } else {
      for (int t = 0; t < 32; ++t) { h_sh[t] = t ? h_out[t - 1] : h_out[0]; e_sh[t] = t ? e_out[t - 1] : e_out[0]; }
      h_sh[0] = Hc; e_sh[0] = Ec;
      if (inject) { const uint32_t up = b16::addend(lv.step, lv.step); h_sh[0] = H00 + up; e_sh[0] = Hc + up; }
    }
    for (int t = 0; t < 32; ++t) {
      const int at = tb_off[t] + s;
      if (at < buf_lo[t] || at >= buf_hi[t]) throw 2;      // a row read outside the buffer of the pair the lane is on
      const int code = sm.at(at);
      if (code >= ncodes) throw 1;
      load_S<KP>(&lut[(size_t)code * P], t, S[t]);
      uint32_t gope = gop2;
      if (GL && t == sw) {      // row 0 of a starting pair through the common row (dp_core.cuh, MM16G_NO_EGAP), as the kernel does
        const int c0 = t * K;
        const uint32_t G0 = (BORDER && t == BL) ? NEGb + (uint32_t)(K - 1) * gew : b16::biased(ge * (1 - c0), ge * (1 - c0));
        for (int k = 0; k < K; ++k) { Hp[t][k] = G0 - (uint32_t)k * gew; F[t][k] = NEGb; }
        hleft[t] = G0 + gew;
        e_sh[t] = NEGb;
        gope = p16::pack(MM16G_NO_EGAP, MM16G_NO_EGAP);
      }
      uint32_t e = e_sh[t];
      mm16d_row<K, KP>(hleft[t], e, Hp[t], F[t], S[t], gop2, gope);
      hleft[t] = h_sh[t]; e_out[t] = e; h_out[t] = Hp[t][K - 1];
      if (GL) {
        if (with_pending && t == s - pend_e1) last_row(t, 0, pend_r1, par ^ 1);
        if (with_pending && t == s - pend_e2) last_row(t, 1, pend_r2, par ^ 1);
        if (t == s - ev1) last_row(t, 0, r1, par);
        if (t == s - ev2) last_row(t, 1, r2, par);
      }
    }
  };
  int lvl = 0, pend_lvl = 0, lv_i = 0;
  auto border_set = [&](uint32_t hc, uint32_t ec) {      // the border lane on the fixed point of (hc, ec), as the kernel does
    if (!BORDER) return;
    for (int k = 0; k < K; ++k) { Hp[BL][k] = hc; F[BL][k] = ec; }
    hleft[BL] = hc; h_out[BL] = hc; e_out[BL] = ec;
  };
  if (GL) border_set(NEGb, NEGb);
  auto capture = [&](int half, int idx, int par, int level) {
    if (GL) {
      (half ? sc2_out : sc1_out)[idx] = fold(half, (half ? l2s[idx] : l1s[idx]) - 1);
    } else {
      const uint32_t hs = Hp[tstar][kstar];
      const int drift = ge * (lq - 1 + (half ? l2s[idx] : l1s[idx]) - 1);
      (half ? sc2_out : sc1_out)[idx] = (half ? b16::hi(hs) : b16::lo(hs)) + drift - level;
    }
  };
  for (int i = 0; i <= n_i; ++i) {
    const bool real = i < n_i;
    int l1 = 0, l2 = 0, has2 = 0, rows = MM16G_MIN_ROWS, bufoff = zero + 32;
    if (real) { l1 = l1s[i]; l2 = l2s[i]; has2 = l2 > 0; rows = mm16g_rows(l1, l2); bufoff = tbuf0 + (i % 3) * stride; }
    const int cs1 = real ? l1 - 1 + tstar : -1, cs2 = (real && has2) ? l2 - 1 + tstar : -1;
    const int ev1 = (GL && real) ? l1 - 1 : NOEV, ev2 = (GL && real && has2) ? l2 - 1 : NOEV;
    for (int t = 0; t < 32; ++t) tb_off[t] += prev_rows;
    const int par = i & 1;
    lv_i = (i == 0 || !real || lv_i + 1 >= lv.n) ? 0 : lv_i + 1;
    explicit_reset = lv_i == 0;
    const bool inj_next = !GL && i + 1 < n_i && lv_i + 1 < lv.n;
    pend_lvl = lvl;
    lvl = GL ? 0 : lv.lvl0 + lv_i * lv.step;
    if (!GL) {
      const uint32_t lw = b16::addend(lvl, lvl); Hc = Hc0 + lw; Ec = Ec0 + lw; H00 = H000 + lw;
      if (explicit_reset) border_set(Hc, Ec);
    }
    for (int s = 0; s < MM16G_MIN_ROWS; ++s) {
      step(s, s, bufoff, ev1, l1 - 1, ev2, l2 - 1, true, par);
      if (s == pend_s1) capture(0, pend_i1, par ^ 1, pend_lvl);
      if (s == pend_s2) capture(1, pend_i2, par ^ 1, pend_lvl);
      if (s == cs1) capture(0, i, par, lvl);
      if (s == cs2) capture(1, i, par, lvl);
    }
    if (i + 2 < n_i) fill((i + 2) % 3, i + 2);
    int s = MM16G_MIN_ROWS;
    if (GL) {
      const int s_chk = std::min(rows, std::max(MM16G_MIN_ROWS, std::min(ev1, ev2)));
      for (; s < s_chk; ++s) step(s, -1, 0, NOEV, 0, NOEV, 0, false, par);
      for (; s < rows; ++s) {
        step(s, -1, 0, ev1, l1 - 1, ev2, l2 - 1, false, par);
        if (s == cs1) capture(0, i, par, lvl);
        if (s == cs2) capture(1, i, par, lvl);
      }
    } else {
      while (s < rows) {
        int stop = rows;
        if (cs1 >= s && cs1 < stop) stop = cs1 + 1;
        if (cs2 >= s && cs2 < stop) stop = cs2 + 1;
        for (; s < stop; ++s) step(s, -1, 0, NOEV, 0, NOEV, 0, false, par);
        if (s - 1 == cs1) capture(0, i, par, lvl);
        if (s - 1 == cs2) capture(1, i, par, lvl);
      }
      if (inj_next) {      // the separator row of lane 0: the next pair's level is injected from the left
        inject = true;
        step(s, -1, 0, NOEV, 0, NOEV, 0, false, par);
        inject = false;
        { const uint32_t up = b16::addend(lv.step, lv.step); border_set(Hc + up, Ec + up); }
        if (s == cs1) capture(0, i, par, lvl);
        if (s == cs2) capture(1, i, par, lvl);
        rows += 1;
      }
    }
    pend_s1 = cs1 >= rows ? cs1 - rows : -1; pend_i1 = i;
    pend_s2 = cs2 >= rows ? cs2 - rows : -1; pend_i2 = i;
    pend_e1 = ev1 == NOEV ? NOEV : ev1 - rows; pend_r1 = l1 - 1;
    pend_e2 = ev2 == NOEV ? NOEV : ev2 - rows; pend_r2 = l2 - 1;
    prev_rows = rows;
  }
}

template <bool GL, int K>
static void emu_mm16g_v(bool border, const unsigned char *q, int lq, int n_pairs, const unsigned char *const *t1s, const int *l1s,
                        const unsigned char *const *t2s, const int *l2s, int alpha, const int *M, int go, int ge, int *sc1, int *sc2) {
  if constexpr (mm16g_border(K, GL)) {
    if (border) { emu_mm16g_t<GL, K, true>(q, lq, n_pairs, t1s, l1s, t2s, l2s, alpha, M, go, ge, sc1, sc2); return; }
  }
  emu_mm16g_t<GL, K, false>(q, lq, n_pairs, t1s, l1s, t2s, l2s, alpha, M, go, ge, sc1, sc2);
}

template <bool GL>
static int emu_mm16g_k(int K, bool border, const unsigned char *q, int lq, int n_pairs, const unsigned char *const *t1s, const int *l1s,
                       const unsigned char *const *t2s, const int *l2s, int alpha, const int *M, int go, int ge, int *sc1, int *sc2) {
  switch (K) {
    case 2: emu_mm16g_v<GL, 2>(border, q, lq, n_pairs, t1s, l1s, t2s, l2s, alpha, M, go, ge, sc1, sc2); return 0;
    case 3: emu_mm16g_v<GL, 3>(border, q, lq, n_pairs, t1s, l1s, t2s, l2s, alpha, M, go, ge, sc1, sc2); return 0;
    case 4: emu_mm16g_v<GL, 4>(border, q, lq, n_pairs, t1s, l1s, t2s, l2s, alpha, M, go, ge, sc1, sc2); return 0;
    case 5: emu_mm16g_v<GL, 5>(border, q, lq, n_pairs, t1s, l1s, t2s, l2s, alpha, M, go, ge, sc1, sc2); return 0;
    case 6: emu_mm16g_v<GL, 6>(border, q, lq, n_pairs, t1s, l1s, t2s, l2s, alpha, M, go, ge, sc1, sc2); return 0;
    case 7: emu_mm16g_v<GL, 7>(border, q, lq, n_pairs, t1s, l1s, t2s, l2s, alpha, M, go, ge, sc1, sc2); return 0;
    case 8: emu_mm16g_v<GL, 8>(border, q, lq, n_pairs, t1s, l1s, t2s, l2s, alpha, M, go, ge, sc1, sc2); return 0;
    case 9: emu_mm16g_v<GL, 9>(border, q, lq, n_pairs, t1s, l1s, t2s, l2s, alpha, M, go, ge, sc1, sc2); return 0;
    case 10: emu_mm16g_v<GL, 10>(border, q, lq, n_pairs, t1s, l1s, t2s, l2s, alpha, M, go, ge, sc1, sc2); return 0;
    case 12: emu_mm16g_v<GL, 12>(border, q, lq, n_pairs, t1s, l1s, t2s, l2s, alpha, M, go, ge, sc1, sc2); return 0;
    case 16: emu_mm16g_v<GL, 16>(border, q, lq, n_pairs, t1s, l1s, t2s, l2s, alpha, M, go, ge, sc1, sc2); return 0;
  }
  return -1;
}

// query columns the border-lane instantiation of the streaming kernel serves with strips of K (0: there is none for this mode / width)
extern "C" int emu_mm16g_border_capacity(int K, int glocal) { return mm16g_border(K, glocal != 0) ? 31 * K : 0; }

extern "C" int emu_mm16g(int mode, int K, int border, const unsigned char *q, int lq, int n_pairs, const unsigned char *const *t1s, const int *l1s,
                         const unsigned char *const *t2s, const int *l2s, int alpha, const int *M, int go, int ge,
                         int *sc1, int *sc2) {
  const bool glocal = mode == GLOCAL;
  if (32 * K < lq) return -1;
  // the border-lane instantiation: up to 31K columns, and its fixed point needs Ec <= Hc, i.e. go <= ge (align_abi.cu keeps other jobs off it)
  if (border && (!mm16g_border(K, glocal) || 31 * K < lq || go > ge)) return -1;
  try {
    if (mode == GLOBAL) return emu_mm16g_k<false>(K, border != 0, q, lq, n_pairs, t1s, l1s, t2s, l2s, alpha, M, go, ge, sc1, sc2);
    if (mode == GLOCAL) return emu_mm16g_k<true>(K, border != 0, q, lq, n_pairs, t1s, l1s, t2s, l2s, alpha, M, go, ge, sc1, sc2);
  } catch (...) { return -2; }   // an out-of-range symbol read: a pointer bookkeeping error
  return -1;
}

// ---------------------------------------------------------------------------------------------
// Variant scan (dp_core.cuh::variants_walk, the function variants_kernel runs per thread) on the host.
// recs: 5 ints per record (kind, start, pos, len, run_start).  Returns the record count or a VAR_ERR_* code.
extern "C" int emu_variants(const unsigned char *sample, int ls, const unsigned char *ref, int lr, int softmask, int gap,
                            int *recs, int cap) {
  if (ls != lr) return VAR_ERR_LENGTH;
  int k = 0;
  const int rc = variants_walk(sample, ref, ls, softmask != 0, (unsigned char)gap, [&](int kind, int start, int pos, int len, int at) {
    if (k < cap) { int *o = recs + 5 * k; o[0] = kind; o[1] = start; o[2] = pos; o[3] = len; o[4] = at; }
    ++k;
  });
  return rc == VAR_OK ? k : rc;
}

// ---------------------------------------------------------------------------------------------
// tb16 (kernels.cuh::tb16_kernel): two pairs in the 16-bit halves, lane t owns columns [tK, (t+1)K), rows skewed
// by lane.  Mirrors the kernel's item loop for ONE item with the device row function (dp_core.cuh::tb16_row) and
// the device direction-word layout (Tb16Reader); the walk back uses the generic walk over the reader's nibble view.
template <int MODE, int K>
static void emu_tb16_t(const unsigned char *qa0, int la0, const unsigned char *tb0, int lb0,
                       const unsigned char *qa1, int la1, const unsigned char *tb1, int lb1,
                       int alpha, const int *M, int go, int ge, int pad_score,
                       int *score, int *nfrag, FragOut *frags0, FragOut *frags1, int cap) {
  constexpr int V = Tb16Cfg<K>::V, P = Tb16Cfg<K>::P;
  const int a1 = alpha + 1, pad = alpha;
  std::vector<int> mext((size_t)a1 * a1, pad_score);
  for (int i = 0; i < alpha; ++i) for (int j = 0; j < alpha; ++j) mext[i * a1 + j] = M[i * alpha + j];
  const int rows = std::max(lb0, lb1), tl = (std::max(la0, la1) - 1) / K, nsteps = rows + tl;
  const int rows_cap = (rows + 15) & ~15, steps4 = tb16_steps4(rows_cap);
  std::vector<uint16_t> tb(rows);
  for (int r = 0; r < rows; ++r) tb[r] = (uint16_t)((r < lb0 ? tb0[r] : pad) | ((r < lb1 ? tb1[r] : pad) << 8));
  std::vector<uint8_t> qc(2 * P, (uint8_t)pad);
  for (int c = 0; c < 32 * K; ++c) { qc[c] = c < la0 ? qa0[c] : pad; qc[P + c] = c < la1 ? qa1[c] : pad; }
  std::vector<uint32_t> prof((size_t)2 * a1 * P);
  for (int idx = 0; idx < a1 * P; ++idx) {
    const int sym = idx / P, pos = idx - sym * P;
    const int j = pos % V, it = pos / V, t = it & 31, i = it >> 5, k = i * V + j;
    prof[idx] = b16::addend(mext[qc[t * K + k] * a1 + sym], 0);
    prof[a1 * P + idx] = b16::addend(0, mext[qc[P + t * K + k] * a1 + sym]);
  }
  const uint32_t gow = b16::addend(go, go), gew = b16::addend(ge, ge), zerob = b16::biased(0, 0);
  std::vector<uint32_t> dir((size_t)K * steps4 * 32, 0xdeadbeefu);
  struct Lane { uint32_t Hp[K], F[K], acc[K], cand[K], hleft, h_out, e_out; int crl[K], crh[K]; };
  std::vector<Lane> L(32);
  for (int t = 0; t < 32; ++t) {
    const int c0 = t * K;
    for (int k = 0; k < K; ++k) {
      const int c = c0 + k;
      if (MODE == GLOBAL) { L[t].Hp[k] = b16::biased(go + ge * c, go + ge * c); L[t].F[k] = b16::biased(2 * go + ge * c, 2 * go + ge * c); }
      else { L[t].Hp[k] = zerob; L[t].F[k] = b16::biased(go, go); }
      L[t].acc[k] = 0;
      L[t].cand[k] = b16::biased(-1, -1); L[t].crl[k] = 0; L[t].crh[k] = 0;
    }
    L[t].hleft = (MODE == GLOBAL && c0) ? b16::biased(go + ge * (c0 - 1), go + ge * (c0 - 1)) : zerob;
    L[t].h_out = zerob; L[t].e_out = zerob;
  }
  const uint32_t Hv0 = MODE == GLOBAL ? b16::biased(go, go) : zerob;
  const uint32_t Ev0 = MODE == GLOBAL ? b16::biased(2 * go, 2 * go) : b16::biased(go, go);
  if (MODE == LOCAL) L[31].e_out = Ev0;          // local: lane 31 owns no column and holds column -1 for lane 0 (tb16_max_la)
  const int t0 = (la0 - 1) / K, k0 = (la0 - 1) - t0 * K, t1 = (la1 - 1) / K, k1 = (la1 - 1) - t1 * K;
  int corner0 = 0, corner1 = 0;
  for (int s4 = 0; s4 < nsteps; s4 += 4) {
    for (int j = 0; j < 4; ++j) {
      const int s = s4 + j;
      uint32_t h_sh[32], e_sh[32];
      for (int t = 0; t < 32; ++t) { h_sh[t] = L[(t + 31) & 31].h_out; e_sh[t] = L[(t + 31) & 31].e_out; }
      for (int t = 0; t < 32; ++t) {
        const int r = s - t;
        if (t == 0 && MODE == GLOBAL) { h_sh[0] = Hv0 + (uint32_t)r * gew; e_sh[0] = Ev0 + (uint32_t)r * gew; }
        if (r >= 0 && r < rows && t <= tl) {
          const unsigned code = tb[r];
          uint32_t Slo[K], Shi[K];
          for (int k = 0; k < K; ++k) {
            Slo[k] = prof[(code & 255u) * P + tb16_slot(t, k, V)];
            Shi[k] = prof[(a1 + (code >> 8)) * P + tb16_slot(t, k, V)];
          }
          uint32_t e = e_sh[t];
          tb16_row<MODE, K>(L[t].hleft, e, L[t].Hp, L[t].F, Slo, Shi, gow, gew, zerob, 0x03030303u << (2 * j), j == 0, L[t].acc, L[t].cand, L[t].crl, L[t].crh, r);
          L[t].hleft = h_sh[t]; L[t].e_out = e; L[t].h_out = L[t].Hp[K - 1];
          if (MODE == GLOBAL) {
            if (r == lb0 - 1 && t == t0) corner0 = b16::lo(L[t].Hp[k0]);
            if (r == lb1 - 1 && t == t1) corner1 = b16::hi(L[t].Hp[k1]);
          }
        }
      }
    }
    for (int t = 0; t < 32; ++t)
      for (int k = 0; k < K; ++k) dir[((size_t)k * steps4 + (s4 >> 2)) * 32 + t] = L[t].acc[k];
  }
  int sc[2], ec[2], er[2];
  if (MODE == GLOBAL) { sc[0] = corner0; sc[1] = corner1; ec[0] = la0 - 1; er[0] = lb0 - 1; ec[1] = la1 - 1; er[1] = lb1 - 1; }
  else {
    long long key[2] = {-1, -1};
    for (int t = 0; t < 32; ++t) {
      int b[2] = {-1, -1}, bc[2] = {0, 0}, br[2] = {0, 0};
      for (int k = 0; k < K; ++k) {
        const int v0 = b16::lo(L[t].cand[k]), v1 = b16::hi(L[t].cand[k]);
        if (v0 >= 0 && v0 >= b[0]) { b[0] = v0; bc[0] = t * K + k; br[0] = L[t].crl[k]; }
        if (v1 >= 0 && v1 >= b[1]) { b[1] = v1; bc[1] = t * K + k; br[1] = L[t].crh[k]; }
      }
      for (int h = 0; h < 2; ++h)
        key[h] = std::max(key[h], b[h] < 0 ? -1ll : (((long long)b[h] << 40) | ((long long)bc[h] << 20) | br[h]));
    }
    for (int h = 0; h < 2; ++h) {
      if (key[h] < 0) { sc[h] = 0; ec[h] = 0; er[h] = 0; }
      else { sc[h] = (int)(key[h] >> 40); ec[h] = (int)((key[h] >> 20) & 0xfffff); er[h] = (int)(key[h] & 0xfffff); }
    }
  }
  for (int h = 0; h < 2; ++h) {
    Tb16Reader<K> rd; rd.w = dir.data(); rd.steps4 = steps4; rd.half = h;
    score[h] = sc[h];
    nfrag[h] = walk_back<MODE>(rd, ec[h], er[h], h ? frags1 : frags0, cap);
  }
}

template <int MODE>
static int emu_tb16_k(int K, const unsigned char *qa0, int la0, const unsigned char *tb0, int lb0,
                      const unsigned char *qa1, int la1, const unsigned char *tb1, int lb1,
                      int alpha, const int *M, int go, int ge, int pad_score, int *score, int *nfrag, FragOut *f0, FragOut *f1, int cap) {
#define ABV_TB16_CASE(KK) case KK: emu_tb16_t<MODE, KK>(qa0, la0, tb0, lb0, qa1, la1, tb1, lb1, alpha, M, go, ge, pad_score, score, nfrag, f0, f1, cap); return 0;
  switch (K) {
    ABV_TB16_CASE(2) ABV_TB16_CASE(4) ABV_TB16_CASE(5) ABV_TB16_CASE(6) ABV_TB16_CASE(7) ABV_TB16_CASE(8) ABV_TB16_CASE(10)
    default: return -1;
  }
#undef ABV_TB16_CASE
}

// fragments come back in REVERSE sequence order (as the kernel's scratch holds them)
extern "C" int emu_tb16(int mode, int K, const unsigned char *qa0, int la0, const unsigned char *tb0, int lb0,
                        const unsigned char *qa1, int la1, const unsigned char *tb1, int lb1,
                        int alpha, const int *M, int go, int ge, int pad_score, int *score, int *nfrag, FragOut *f0, FragOut *f1, int cap) {
  if (std::max(la0, la1) > tb16_max_la(mode == LOCAL, K)) return -1;      // what the host plan never sends (local: lane 31 stays free)
  if (mode == LOCAL) return emu_tb16_k<LOCAL>(K, qa0, la0, tb0, lb0, qa1, la1, tb1, lb1, alpha, M, go, ge, pad_score, score, nfrag, f0, f1, cap);
  if (mode == GLOBAL) return emu_tb16_k<GLOBAL>(K, qa0, la0, tb0, lb0, qa1, la1, tb1, lb1, alpha, M, go, ge, pad_score, score, nfrag, f0, f1, cap);
  return -1;
}

// the level plan of the streaming global kernel (dp_core.cuh), for the property test
extern "C" void emu_levels(int P, int max_rows, int go, int ge, int maxabs, int maxpos, int *out3) {
  const Mm16gLevels L = mm16g_levels(P, max_rows, go, ge, maxabs, maxpos);
  out3[0] = L.n; out3[1] = L.lvl0; out3[2] = L.step;
}
