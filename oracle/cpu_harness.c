/*
 * oracle/cpu_harness.c -- multi-threaded CPU timing harness.  TEST/BENCH INFRASTRUCTURE ONLY.
 *
 * Drives a reference-ABI aligner (`Alignment *f(sa, la, sb, lb, alpha, matrix, go, ge)`,
 * lib/align/align.h:49) over a pair list or a many-to-many job on N pthreads, the tight-loop
 * CPU baseline of SURVEY.md section 8(d).  The aligner is passed in as a function pointer so the
 * same harness times either the unmodified reference (oracle/_ref/libalign_ref.so) or the
 * restatement in align_oracle.c (harness_port_*), each doing the full work the C extension does
 * per call: allocate and fill the jump matrix, trace back, build and free the fragment list.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "align_oracle.h"

/* ABI mirror of lib/align/align.h:31-41 */
typedef struct HFrag { struct HFrag *next; int type, sa_start, sb_start, hsp_len; } HFrag;
typedef struct { HFrag *align_frag; int frag_count; int score; } HAlignment;

typedef HAlignment *(*raw_align_fn)(const unsigned char *, int, const unsigned char *, int, int, int *, int, int);
typedef void (*free_fn)(HAlignment *);

static double now_s(void) {
  struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ---- the restatement wrapped in the reference ABI (kind = "port") ---- */
static HAlignment *port_call(int mode, const unsigned char *sa, int la, const unsigned char *sb, int lb,
                             int alpha, int *M, int go, int ge) {
  int cap = la + lb + 2, score = 0;
  if (la <= 0 || lb <= 0) return NULL;
  orc_frag *fr = (orc_frag *)malloc(sizeof(orc_frag) * (size_t)cap);
  if (!fr) return NULL;
  int n = orc_align_raw(mode, sa, la, sb, lb, alpha, M, go, ge, &score, fr, cap);
  if (n < 0) { free(fr); return NULL; }
  HAlignment *al = (HAlignment *)malloc(sizeof *al);
  HFrag *head = NULL;
  for (int i = n - 1; i >= 0; i--) {
    HFrag *f = (HFrag *)malloc(sizeof *f);
    f->next = head; f->type = fr[i].type; f->sa_start = fr[i].sa_start; f->sb_start = fr[i].sb_start; f->hsp_len = fr[i].hsp_len;
    head = f;
  }
  free(fr);
  al->align_frag = head; al->frag_count = n; al->score = score;
  return al;
}
HAlignment *harness_port_overlap(const unsigned char *a, int la, const unsigned char *b, int lb, int al, int *M, int go, int ge) { return port_call(ORC_OVERLAP, a, la, b, lb, al, M, go, ge); }
HAlignment *harness_port_local(const unsigned char *a, int la, const unsigned char *b, int lb, int al, int *M, int go, int ge) { return port_call(ORC_LOCAL, a, la, b, lb, al, M, go, ge); }
HAlignment *harness_port_global(const unsigned char *a, int la, const unsigned char *b, int lb, int al, int *M, int go, int ge) { return port_call(ORC_GLOBAL, a, la, b, lb, al, M, go, ge); }
HAlignment *harness_port_glocal(const unsigned char *a, int la, const unsigned char *b, int lb, int al, int *M, int go, int ge) { return port_call(ORC_GLOCAL, a, la, b, lb, al, M, go, ge); }
void harness_port_free(HAlignment *al) {
  if (!al) return;
  HFrag *f = al->align_frag;
  while (f) { HFrag *n = f->next; free(f); f = n; }
  free(al);
}

/* ---- job descriptions ---- */
typedef struct {
  raw_align_fn fn; free_fn ffn;
  const unsigned char *q; const long long *q_off; int n_q;
  const unsigned char *t; const long long *t_off; int n_t;   /* n_t < 0: one-to-one pair list */
  int alpha; int *M; int go, ge, seed;
  int *out_score; int *out_idx;
  volatile int next;      /* work counter */
} job_t;

static void *worker(void *arg) {
  job_t *j = (job_t *)arg;
  for (;;) {
    int i = __sync_fetch_and_add(&j->next, 1);
    if (i >= j->n_q) break;
    const unsigned char *qa = j->q + j->q_off[i];
    int la = (int)(j->q_off[i + 1] - j->q_off[i]);
    if (j->n_t < 0) {                        /* pair list */
      HAlignment *al = j->fn(qa, la, j->t + j->t_off[i], (int)(j->t_off[i + 1] - j->t_off[i]), j->alpha, j->M, j->go, j->ge);
      j->out_score[i] = al ? al->score : -2147483647 - 1;
      if (al) j->ffn(al);
    } else {                                 /* best hit over all targets, first maximum wins */
      int bs = j->seed, bi = -1;
      for (int k = 0; k < j->n_t; k++) {
        HAlignment *al = j->fn(qa, la, j->t + j->t_off[k], (int)(j->t_off[k + 1] - j->t_off[k]), j->alpha, j->M, j->go, j->ge);
        if (!al) continue;
        if (al->score > bs) { bs = al->score; bi = k; }
        j->ffn(al);
      }
      j->out_score[i] = bs; j->out_idx[i] = bi;
    }
  }
  return NULL;
}

static double run_job(job_t *j, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
  j->next = 0;
  double t0 = now_s();
  for (int i = 0; i < nthreads; i++) pthread_create(&th[i], NULL, worker, j);
  for (int i = 0; i < nthreads; i++) pthread_join(th[i], NULL);
  double t1 = now_s();
  free(th);
  return t1 - t0;
}

/* many-to-many best hit; returns elapsed seconds */
double harness_best_hits(raw_align_fn fn, free_fn ffn,
                         const unsigned char *q, const long long *q_off, int n_q,
                         const unsigned char *t, const long long *t_off, int n_t,
                         int alpha, int *M, int go, int ge, int seed,
                         int *best_score, int *best_idx, int nthreads) {
  job_t j; memset(&j, 0, sizeof j);
  j.fn = fn; j.ffn = ffn; j.q = q; j.q_off = q_off; j.n_q = n_q; j.t = t; j.t_off = t_off; j.n_t = n_t;
  j.alpha = alpha; j.M = M; j.go = go; j.ge = ge; j.seed = seed; j.out_score = best_score; j.out_idx = best_idx;
  return run_job(&j, nthreads);
}

/* one-to-one pair list; returns elapsed seconds */
double harness_pairs(raw_align_fn fn, free_fn ffn,
                     const unsigned char *a, const long long *a_off,
                     const unsigned char *b, const long long *b_off, int n,
                     int alpha, int *M, int go, int ge, int *scores, int nthreads) {
  job_t j; memset(&j, 0, sizeof j);
  j.fn = fn; j.ffn = ffn; j.q = a; j.q_off = a_off; j.n_q = n; j.t = b; j.t_off = b_off; j.n_t = -1;
  j.alpha = alpha; j.M = M; j.go = go; j.ge = ge; j.out_score = scores;
  return run_job(&j, nthreads);
}
