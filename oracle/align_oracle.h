/* oracle/align_oracle.h -- TEST INFRASTRUCTURE ONLY (see align_oracle.c). */
#ifndef AMBIVERT_ORACLE_H
#define AMBIVERT_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* alignment modes; numbering shared with include/ambivert_align.h */
enum { ORC_OVERLAP = 0, ORC_LOCAL = 1, ORC_GLOBAL = 2, ORC_GLOCAL = 3 };
/* fragment types: lib/align/align.h:29 */
enum { ORC_A_GAP = 0, ORC_B_GAP = 1, ORC_MATCH = 2 };

#define ORC_ERR_ARGS      (-1)  /* the reference returns NULL (global_align.c:118-122) */
#define ORC_ERR_UNDEFINED (-2)  /* the reference has undefined behaviour (glocal lb==1, overlap la==1) */

typedef struct { int type, sa_start, sb_start, hsp_len; } orc_frag;

/* Returns the fragment count (>= 0) or a negative ORC_ERR_*.  frags (capacity cap) receives the
 * fragments in sequence order; pass frags=NULL, cap=0 for score only. */
int orc_align_raw(int mode, const unsigned char *sa, int la, const unsigned char *sb, int lb,
                  int alpha_len, const int *matrix, int gap_open, int gap_extend,
                  int *score, orc_frag *frags, int cap);

int orc_align(int mode, const char *seqa, int la, const char *seqb, int lb,
              int alpha_len, const unsigned char *map256, const int *matrix,
              int gap_open, int gap_extend, int *score, orc_frag *frags, int cap);

void orc_encode(const char *in, unsigned char *out, int n, const unsigned char *map256);

int orc_best_hits(int mode,
                  const unsigned char *q, const long long *q_off, int n_q,
                  const unsigned char *t, const long long *t_off, int n_t,
                  int alpha_len, const int *matrix, int gap_open, int gap_extend,
                  int seed_score, int *best_score, int *best_idx);

#ifdef __cplusplus
}
#endif
#endif
