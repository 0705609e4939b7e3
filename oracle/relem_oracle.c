/* oracle/relem_oracle.c -- TEST INFRASTRUCTURE ONLY (see relem_oracle.h).
 *
 * Plain-C CPU restatement of the RNAelem hot path.  It keeps the reference's visiting order, its scatter-style
 * outside pass and its pairwise log-space arithmetic (util.hpp:192-229), so that sums and Viterbi ties come
 * out as in the reference.  Each function cites the reference lines it follows.
 *
 * Parity pinning (tests/test_oracle.py): the reference's own known-answer tests (RNAelem-test/test.cpp:93-203
 * path and emission counts under its debug flags; test-exact.cpp:90-137 RNAfold dot plot) and the outputs of the
 * unmodified reference compiled into oracle/_ref (tests/golden/*).
 */
#define _POSIX_C_SOURCE 200809L
#include "relem_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../rnaelem_b200/csrc/energy_data.inc" /* third-party ViennaRNA parameter DATA (integers), not code */

#define NEGINF (-INFINITY)
enum { ST_P = 0, ST_E, ST_M, ST_B, ST_1, ST_2, ST_L, ST_O, NSTATE };
enum { TT_E_H = 0, TT_P_E, TT_P_P, TT_O_O, TT_O_OP, TT_E_P, TT_E_M, TT_M_M, TT_M_B, TT_B_12, TT_1_B, TT_1_2,
       TT_2_2, TT_2_P, TT_L_L, NTRANS };
static const int TT_FIRST[NTRANS] = {ST_E, ST_P, ST_P, ST_O, ST_O, ST_E, ST_E, ST_M, ST_M, ST_B, ST_1, ST_1, ST_2, ST_2, ST_L};
static const int TT_SECOND[NTRANS] = {ST_L, ST_E, ST_P, ST_O, ST_P, ST_P, ST_M, ST_M, ST_B, ST_1, ST_B, ST_2, ST_2, ST_P, ST_L};

/* bio_sequence.hpp:22-28 */
static const int BP[5][5] = {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 5}, {0, 0, 0, 1, 0}, {0, 0, 2, 0, 3}, {0, 6, 0, 4, 0}};
static const char NACGU[] = "NACGU";

/* ---------------------------------------------------------------------------------------- log-space math */
static double lse(double x, double y) { /* util.hpp:195-202 */
  if (y == NEGINF) return x;
  if (x == NEGINF) return y;
  return x < y ? y + log1p(exp(x - y)) : x + log1p(exp(y - x));
}
static void addl(double* x, double y) { *x = lse(*x, y); }

typedef struct { int id, l, r; } IS;
typedef struct { int n, cap; int* v; } ivec;
static void iv_push(ivec* a, int x) {
  if (a->n == a->cap) { a->cap = a->cap ? 2 * a->cap : 8; a->v = (int*)realloc(a->v, sizeof(int) * a->cap); }
  a->v[a->n++] = x;
}
typedef struct { int k, l, t, e1, s1; } Trace;

struct orc_model {
  /* energy (energy_param.hpp:61-88) */
  double hairpin[31], mm_h[7][5][5], mm_i[7][5][5], mm_m[8][5][5], mm_1ni[7][5][5], mm_23i[7][5][5], mm_ext[8][5][5];
  double stack[7][7], bulge[31], term_au, int11[8][8][5][5], int21[8][8][5][5][5], int22[8][8][5][5][5][5];
  double internal[31], dangle5[8][5], dangle3[8][5], ninio[31], mlintern, mlclosing, ml_base, lxc37;
  char *triloops, *tetraloops, *hexaloops;
  double triloop[64], tetraloop[64], hexaloop[64];
  int ntri, ntetra, nhexa;
  /* automaton (profile_hmm.hpp) */
  int M, S;
  int *node, *pair, *theta_id;
  ivec *edge_to, *edge_from;
  unsigned char *reach, *reach_loop;
  IS* state;
  int* n2s;
  ivec loop_state, *right, *left, *pairt;
  int nquad; int* quad;
  int nrow; int* row_size; double** theta;
  double lambda[2], tau, ltau;
  int max_span, max_iloop, no_rss, no_prf, no_ene;
  double min_bpp, min_lnbpp;
  int no_theta, no_turn; char* fix_s;
  /* per sequence */
  int L, W, C;
  const int* seq;
  const double* ws;
  unsigned char *bp_ok, *left_ok;
  double bpp_eff;
  double *ein, *eout, *ein_o, *eout_o;
  double *in, *out, *in_o, *out_o;
  Trace *trace, *trace_o;
  long tab_n, otab_n;
  /* pass state */
  int mode; /* see MODE_* */
  double ZL; double* dEH; double* dEN; /* theta-flat */
  double *PysL, *PyiL, *PyeL;
  int Ys, Ye;
  int* row_off;
  double cnt_struct, cnt_motif;
  int psi_dummy;
  int* psihat; char* rss;
};
enum { MODE_TRAIN = 0, MODE_SCAN_START, MODE_SCAN_END, MODE_CYK };

/* -------------------------------------------------------------------------------------------- energy tables */
static const double kT = (37 + 273.15) * 1.98717;
static double smooth(int a) { /* energy_param.hpp:95-106 */
  double z = (double)a;
  if (z / 10. < -1.2283697) return 0.;
  if (0.8660254 < z / 10.) return z;
  return 10. * 0.38490018 * (1. + sin(z / 10. - 0.34242663)) * (1. + sin(z / 10. - 0.34242663));
}
static double lw(int z, int smo) { /* energy_param.hpp:108-114, 175-180 */
  if (z == RELEM_EINF) return NEGINF;
  return smo ? smooth(-z) * 10. / kT : -z * 10. / kT;
}
static void fill_inf(double* p, size_t n) { for (size_t k = 0; k < n; ++k) p[k] = NEGINF; }

struct eset {
  const int *stack, *mm_h, *mm_i, *mm_1ni, *mm_23i, *mm_m, *mm_ext, *d5, *d3, *i11, *i21, *i22, *hp, *bu, *in, *sc;
  double lxc; int ntri, ntetra, nhexa; const char* const *tri, *const *tetra, *const *hexa; const int *tri_e, *tetra_e, *hexa_e;
};
static char* join_loops(const char* const* names, int n) {
  size_t len = 1;
  for (int k = 0; k < n; ++k) len += strlen(names[k]) + 1;
  char* s = (char*)calloc(len, 1);
  for (int k = 0; k < n; ++k) { strcat(s, names[k]); strcat(s, " "); }
  return s;
}
static void load_energy(orc_model* m, int which) { /* which sub-block each section fills: energy_param.hpp:519-640 */
  struct eset T = {RELEM_T2004_stack, RELEM_T2004_mm_h, RELEM_T2004_mm_i, RELEM_T2004_mm_1ni, RELEM_T2004_mm_23i,
                   RELEM_T2004_mm_m, RELEM_T2004_mm_ext, RELEM_T2004_dangle5, RELEM_T2004_dangle3, RELEM_T2004_int11,
                   RELEM_T2004_int21, RELEM_T2004_int22, RELEM_T2004_hairpin, RELEM_T2004_bulge, RELEM_T2004_interior,
                   RELEM_T2004_scalars, RELEM_T2004_lxc37, RELEM_T2004_ntri, RELEM_T2004_ntetra, RELEM_T2004_nhexa,
                   RELEM_T2004_tri_seq, RELEM_T2004_tetra_seq, RELEM_T2004_hexa_seq, RELEM_T2004_tri_e,
                   RELEM_T2004_tetra_e, RELEM_T2004_hexa_e};
  struct eset A = {RELEM_A2007_stack, RELEM_A2007_mm_h, RELEM_A2007_mm_i, RELEM_A2007_mm_1ni, RELEM_A2007_mm_23i,
                   RELEM_A2007_mm_m, RELEM_A2007_mm_ext, RELEM_A2007_dangle5, RELEM_A2007_dangle3, RELEM_A2007_int11,
                   RELEM_A2007_int21, RELEM_A2007_int22, RELEM_A2007_hairpin, RELEM_A2007_bulge, RELEM_A2007_interior,
                   RELEM_A2007_scalars, RELEM_A2007_lxc37, RELEM_A2007_ntri, RELEM_A2007_ntetra, RELEM_A2007_nhexa,
                   RELEM_A2007_tri_seq, RELEM_A2007_tetra_seq, RELEM_A2007_hexa_seq, RELEM_A2007_tri_e,
                   RELEM_A2007_tetra_e, RELEM_A2007_hexa_e};
  struct eset* e = which ? &A : &T;
  fill_inf(m->hairpin, 31); fill_inf(&m->mm_h[0][0][0], 175); fill_inf(&m->mm_i[0][0][0], 175);
  fill_inf(&m->mm_m[0][0][0], 200); fill_inf(&m->mm_1ni[0][0][0], 175); fill_inf(&m->mm_23i[0][0][0], 175);
  fill_inf(&m->mm_ext[0][0][0], 200); fill_inf(&m->stack[0][0], 49); fill_inf(m->bulge, 31);
  fill_inf(&m->int11[0][0][0][0], 1600); fill_inf(&m->int21[0][0][0][0][0], 8000);
  fill_inf(&m->int22[0][0][0][0][0][0], 40000); fill_inf(m->internal, 31); fill_inf(&m->dangle5[0][0], 40);
  fill_inf(&m->dangle3[0][0], 40); fill_inf(m->ninio, 31);
  for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) m->stack[a + 1][b + 1] = lw(e->stack[a * 6 + b], 0);
  for (int a = 0; a < 6; ++a) for (int k = 0; k < 25; ++k) {
    m->mm_h[a + 1][k / 5][k % 5] = lw(e->mm_h[a * 25 + k], 0);
    m->mm_i[a + 1][k / 5][k % 5] = lw(e->mm_i[a * 25 + k], 0);
    m->mm_1ni[a + 1][k / 5][k % 5] = lw(e->mm_1ni[a * 25 + k], 0);
    m->mm_23i[a + 1][k / 5][k % 5] = lw(e->mm_23i[a * 25 + k], 0);
  }
  for (int a = 0; a < 7; ++a) for (int k = 0; k < 25; ++k) {
    m->mm_m[a + 1][k / 5][k % 5] = lw(e->mm_m[a * 25 + k], 1);
    m->mm_ext[a + 1][k / 5][k % 5] = lw(e->mm_ext[a * 25 + k], 1);
  }
  for (int a = 0; a < 7; ++a) for (int k = 0; k < 5; ++k) {
    m->dangle5[a + 1][k] = lw(e->d5[a * 5 + k], 1);
    m->dangle3[a + 1][k] = lw(e->d3[a * 5 + k], 1);
  }
  for (int a = 0; a < 7; ++a) for (int b = 0; b < 7; ++b) {
    for (int k = 0; k < 25; ++k) m->int11[a + 1][b + 1][k / 5][k % 5] = lw(e->i11[(a * 7 + b) * 25 + k], 0);
    for (int k = 0; k < 125; ++k) m->int21[a + 1][b + 1][k / 25][(k / 5) % 5][k % 5] = lw(e->i21[(a * 7 + b) * 125 + k], 0);
  }
  /* energy_param.hpp:597-598 clears only the first 8000 entries of int22; the rest, unless read from the file, is
   * never written and holds +0.0 in the reference's binaries (tests/golden/tables_*.npz) */
  for (int k = 8000; k < 40000; ++k) (&m->int22[0][0][0][0][0][0])[k] = 0.;
  for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) for (int k = 0; k < 256; ++k)
    m->int22[a + 1][b + 1][1 + k / 64][1 + (k / 16) % 4][1 + (k / 4) % 4][1 + k % 4] = lw(e->i22[(a * 6 + b) * 256 + k], 0);
  for (int d = 0; d <= 30; ++d) {
    m->hairpin[d] = lw(e->hp[d], 0); m->bulge[d] = lw(e->bu[d], 0); m->internal[d] = lw(e->in[d], 0);
    int x = d * e->sc[0]; if (e->sc[1] < x) x = e->sc[1];
    m->ninio[d] = lw(x, 0);
  }
  m->ml_base = lw(e->sc[2], 0); m->mlclosing = lw(e->sc[3], 0); m->mlintern = lw(e->sc[4], 0); m->term_au = lw(e->sc[5], 0);
  m->lxc37 = e->lxc;
  m->ntri = e->ntri; m->ntetra = e->ntetra; m->nhexa = e->nhexa;
  m->triloops = join_loops(e->tri, e->ntri); m->tetraloops = join_loops(e->tetra, e->ntetra);
  m->hexaloops = join_loops(e->hexa, e->nhexa);
  for (int k = 0; k < e->ntri; ++k) m->triloop[k] = lw(e->tri_e[k], 0);
  for (int k = 0; k < e->ntetra; ++k) m->tetraloop[k] = lw(e->tetra_e[k], 0);
  for (int k = 0; k < e->nhexa; ++k) m->hexaloop[k] = lw(e->hexa_e[k], 0);
}

/* ---- energy_param.hpp:686-708 */
static double sum_ext_m(const orc_model* m, const int* s, int L, int i, int j, int ext) {
  int type = BP[s[i]][s[j]];
  double z = 0.;
  if (0 <= i - 1 && j + 1 < L) {
    z = z + (ext ? m->mm_ext[type][s[i - 1]][s[j + 1]] : m->mm_m[type][s[i - 1]][s[j + 1]]);
    if (type > 2) z = z + m->term_au;
  } else {
    if (0 <= i - 1) z = z + m->dangle5[type][s[i - 1]];
    if (j + 1 < L) z = z + m->dangle3[type][s[j + 1]];
    if (type > 2) z = z + m->term_au;
  }
  return z;
}
static void slice(const int* s, int i, int j, char* out) {
  int n = 0;
  for (int k = i; k < j; ++k) out[n++] = NACGU[s[k]];
  out[n] = 0;
}
/* ---- energy_param.hpp:710-742 */
static double hairpin_energy(const orc_model* m, const int* s, int i, int j) {
  int d = j - i - 1;
  if (d < 1) return NEGINF;
  int type = BP[s[i]][s[j]];
  double z = d <= 30 ? m->hairpin[30 < d ? 30 : d]
                     : m->hairpin[30] - m->lxc37 * log((double)d * (1. / 30)) * 10. * (1. / kT);
  char buf[16];
  if (d < 3) {
  } else if (d == 3) {
    slice(s, i, j + 1, buf);
    const char* p = strstr(m->triloops, buf);
    if (p) return m->triloop[(p - m->triloops) / 6];
    if (type > 2) z = z + m->term_au;
  } else if (d == 4) {
    slice(s, i, j + 1, buf);
    const char* p = strstr(m->tetraloops, buf);
    if (p) return m->tetraloop[(p - m->tetraloops) / 7];
  } else if (d == 6) {
    slice(s, i, j + 1, buf);
    const char* p = strstr(m->hexaloops, buf);
    if (p) return m->hexaloop[(p - m->hexaloops) / 9];
  }
  if (3 < d) z = z + m->mm_h[type][s[i + 1]][s[j - 1]];
  return z;
}
/* ---- energy_param.hpp:744-795 */
static double loop_energy(const orc_model* m, const int* s, int i, int j, int p, int q) {
  int type = BP[s[i]][s[j]], type2 = BP[s[q]][s[p]];
  int u1 = p - i - 1, u2 = j - q - 1, u = u1 > u2 ? u1 : u2;
  double z;
  if (u1 < 0 || u2 < 0 || 30 < u1 + u2) z = NEGINF;
  else if (u1 == 0 && u2 == 0) z = m->stack[type][type2];
  else if (u1 == 0 || u2 == 0) {
    z = m->bulge[u];
    if (u == 1) z = z + m->stack[type][type2];
    else {
      if (type > 2) z = z + m->term_au;
      if (type2 > 2) z = z + m->term_au;
    }
  } else if (u <= 2) {
    if (u1 + u2 == 2) z = m->int11[type][type2][s[i + 1]][s[j - 1]];
    else if (u1 == 1 && u2 == 2) z = m->int21[type][type2][s[i + 1]][s[q + 1]][s[j - 1]];
    else if (u1 == 2 && u2 == 1) z = m->int21[type2][type][s[q + 1]][s[i + 1]][s[p - 1]];
    else z = m->int22[type][type2][s[i + 1]][s[p - 1]][s[q + 1]][s[j - 1]];
  } else {
    z = m->internal[u1 + u2] + m->ninio[abs(u1 - u2)];
    if (u1 == 1 || u2 == 1) z = z + (m->mm_1ni[type][s[i + 1]][s[j - 1]] + m->mm_1ni[type2][s[q + 1]][s[p - 1]]);
    else if (u1 + u2 == 5) z = z + (m->mm_23i[type][s[i + 1]][s[j - 1]] + m->mm_23i[type2][s[q + 1]][s[p - 1]]);
    else z = z + (m->mm_i[type][s[i + 1]][s[j - 1]] + m->mm_i[type2][s[q + 1]][s[p - 1]]);
  }
  return z;
}
double orc_loop_energy(const orc_model* m, const int* s, int L, int i, int j, int p, int q) { (void)L; return loop_energy(m, s, i, j, p, q); }
double orc_hairpin_energy(const orc_model* m, const int* s, int L, int i, int j) { (void)L; return hairpin_energy(m, s, i, j); }
double orc_sum_ext_m(const orc_model* m, const int* s, int L, int i, int j, int ext) { return sum_ext_m(m, s, L, i, j, ext); }

/* ----------------------------------------------------------------------------------------------- automaton */
static int is_loop_node(int c) { return c == 'z' || c == '.' || c == '*' || c == 'o'; }
static int is_bg_node(int c) { return c == 'z' || c == 'o' || c == '*'; }
static void closure(unsigned char* a, int n) { /* profile_hmm.hpp:357-366 */
  for (int k = 0; k < n; ++k) for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j)
    if (a[i * n + k] && a[k * n + j]) a[i * n + j] = 1;
}
static int build_hmm(orc_model* m, const char* pattern) { /* profile_hmm.hpp:188-463 */
  char reg[512];
  int n = 0;
  for (const char* p = pattern; *p; ++p) if (!(*p == '*' && n > 0 && reg[n - 1] == '*')) reg[n++] = *p;
  reg[n] = 0;
  int b = 0; while (reg[b] == '*') ++b;
  memmove(reg, reg + b, n - b + 1); n -= b;
  while (n > 0 && reg[n - 1] == '*') reg[--n] = 0;
  int M = n + 2;
  m->M = M;
  m->node = (int*)calloc(M, sizeof(int));
  m->node[0] = 'z'; for (int k = 0; k < n; ++k) m->node[k + 1] = reg[k]; m->node[M - 1] = 'o';
  m->pair = (int*)malloc(sizeof(int) * M);
  int* stk = (int*)malloc(sizeof(int) * M); int sp = 0;
  for (int h = 0; h < M; ++h) m->pair[h] = -1;
  for (int h = 0; h < M; ++h) {
    if (m->node[h] == '(') stk[sp++] = h;
    else if (m->node[h] == ')') { if (!sp) return -1; int hl = stk[--sp]; m->pair[hl] = h; m->pair[h] = hl; }
  }
  free(stk);
  if (sp) return -1;
  m->edge_to = (ivec*)calloc(M, sizeof(ivec)); m->edge_from = (ivec*)calloc(M, sizeof(ivec));
  for (int h = 0; h < M; ++h) {
    if (h > 0) {
      if (m->node[h - 1] == '*') { iv_push(&m->edge_to[h], h - 2); iv_push(&m->edge_from[h - 2], h); }
      iv_push(&m->edge_to[h], h - 1); iv_push(&m->edge_from[h - 1], h);
    }
    iv_push(&m->edge_to[h], h); iv_push(&m->edge_from[h], h);
  }
  m->theta_id = (int*)malloc(sizeof(int) * M);
  m->row_size = (int*)malloc(sizeof(int) * (M + 1));
  m->nrow = 1; m->row_size[0] = 4;
  for (int h = 0; h < M; ++h) {
    m->theta_id[h] = -1;
    switch (m->node[h]) {
      case ')': m->theta_id[h] = m->nrow; m->row_size[m->nrow++] = 6; break;
      case '.': m->theta_id[h] = m->nrow; m->row_size[m->nrow++] = 4; break;
      case '*': case 'z': case 'o': m->theta_id[h] = 0; break;
      case '(': break;
      default: return -1;
    }
  }
  m->theta = (double**)calloc(m->nrow, sizeof(double*));
  m->row_off = (int*)calloc(m->nrow + 1, sizeof(int));
  for (int r = 0; r < m->nrow; ++r) {
    m->theta[r] = (double*)calloc(m->row_size[r], sizeof(double));
    for (int k = 0; k < m->row_size[r]; ++k) m->theta[r][k] = -log((double)m->row_size[r]);
    m->row_off[r + 1] = m->row_off[r] + m->row_size[r];
  }
  m->reach = (unsigned char*)calloc(M * M, 1); m->reach_loop = (unsigned char*)calloc(M * M, 1);
  for (int h = 0; h < M; ++h) {
    int c = m->node[h];
    if (c == ')') { ivec* e = &m->edge_to[m->pair[h]]; for (int k = 0; k < e->n; ++k) m->reach[e->v[k] * M + h] = 1; }
    else if (c == '(') {}
    else { ivec* e = &m->edge_to[h]; for (int k = 0; k < e->n; ++k) { m->reach[e->v[k] * M + h] = 1; m->reach_loop[e->v[k] * M + h] = 1; } }
    m->reach[h * M + h] = 1; m->reach_loop[h * M + h] = 1;
  }
  closure(m->reach, M); closure(m->reach_loop, M);
  m->state = (IS*)malloc(sizeof(IS) * M * M);
  m->n2s = (int*)malloc(sizeof(int) * M * M);
  for (int k = 0; k < M * M; ++k) m->n2s[k] = -1;
  int S = 0;
  for (int hr = 0; hr < M; ++hr) for (int hl = hr; hl >= 0; --hl)
    if (m->reach[hl * M + hr]) { m->state[S].id = S; m->state[S].l = hl; m->state[S].r = hr; m->n2s[hl * M + hr] = S; ++S; }
  m->S = S;
  for (int s = 0; s < S; ++s) if (m->reach_loop[m->state[s].l * M + m->state[s].r]) iv_push(&m->loop_state, s);
  m->right = (ivec*)calloc(S, sizeof(ivec)); m->left = (ivec*)calloc(S, sizeof(ivec)); m->pairt = (ivec*)calloc(S, sizeof(ivec));
  for (int s = 0; s < S; ++s) {
    IS st = m->state[s];
    if (is_loop_node(m->node[st.r])) {
      ivec* e = &m->edge_to[st.r];
      for (int k = 0; k < e->n; ++k) { int h = e->v[k]; if (st.l <= h && m->reach[st.l * M + h]) iv_push(&m->right[s], m->n2s[st.l * M + h]); }
    }
  }
  for (int s = 0; s < S; ++s) {
    IS st = m->state[s];
    if (is_loop_node(m->node[st.l])) {
      ivec* e = &m->edge_to[st.l];
      for (int k = 0; k < e->n; ++k) { int h = e->v[k]; if (h <= st.r && m->reach[h * M + st.r]) iv_push(&m->left[m->n2s[h * M + st.r]], s); }
    }
  }
  for (int hr = 0; hr < M; ++hr) if (m->node[hr] == ')') {
    int kl = m->pair[hr];
    ivec* e = &m->edge_to[kl];
    for (int a = 0; a < e->n; ++a) {
      int hl = e->v[a], s = m->n2s[hl * M + hr];
      ivec* f = &m->edge_to[hr];
      for (int c = 0; c < f->n; ++c) { int kr = f->v[c]; if (m->reach[kl * M + kr]) iv_push(&m->pairt[s], m->n2s[kl * M + kr]); }
    }
  }
  for (int s = 0; s < S; ++s) {
    IS st = m->state[s];
    if (!is_bg_node(m->node[st.r])) continue;
    ivec* e = &m->edge_from[st.l];
    for (int a = 0; a < e->n; ++a) {
      int hl = e->v[a];
      if (!is_bg_node(m->node[hl])) continue;
      ivec* f = &m->edge_to[st.r];
      for (int c = 0; c < f->n; ++c) { int hr = f->v[c]; if (m->reach[hl * M + hr]) iv_push(&m->pairt[s], m->n2s[hl * M + hr]); }
    }
  }
  m->quad = (int*)malloc(sizeof(int) * 4 * (size_t)m->loop_state.n * m->loop_state.n + 16);
  m->nquad = 0;
  for (int a = 0; a < m->loop_state.n; ++a) for (int c = 0; c < m->loop_state.n; ++c) {
    IS s2 = m->state[m->loop_state.v[a]], s3 = m->state[m->loop_state.v[c]];
    if (s3.r < s2.l || !m->reach[s2.r * M + s3.l] || !m->reach[s2.l * M + s3.r]) continue;
    int* q = m->quad + 4 * m->nquad++;
    q[0] = m->n2s[s2.l * M + s3.r]; q[1] = m->n2s[s2.r * M + s3.l]; q[2] = s2.id; q[3] = s3.id;
  }
  return 0;
}

orc_model* orc_model_new(const char* pattern, int energy_set, int max_span, int max_iloop, double min_bpp,
                         int no_rss, int no_prf, int no_ene) {
  orc_model* m = (orc_model*)calloc(1, sizeof(orc_model));
  load_energy(m, energy_set);
  if (build_hmm(m, pattern)) { free(m); return NULL; }
  m->max_span = max_span; m->max_iloop = max_iloop; m->min_bpp = min_bpp; m->min_lnbpp = log(min_bpp);
  m->no_rss = no_rss; m->no_prf = no_prf; m->no_ene = no_ene;
  m->lambda[0] = m->lambda[1] = 1.; m->tau = 1.; m->ltau = 0.;
  return m;
}
void orc_model_free(orc_model* m) { free(m); /* test helper: the process is short-lived */ }
void orc_model_set_debug(orc_model* m, int no_theta, int no_turn, const char* fix_rss) {
  m->no_theta = no_theta; m->no_turn = no_turn;
  free(m->fix_s); m->fix_s = fix_rss ? strdup(fix_rss) : NULL;
}
void orc_model_set_params(orc_model* m, const double* theta_flat, const double* lambda, double tau) {
  int k = 0;
  for (int r = 0; r < m->nrow; ++r) for (int c = 0; c < m->row_size[r]; ++c) m->theta[r][c] = theta_flat[k++];
  m->lambda[0] = lambda[0]; m->lambda[1] = lambda[1]; m->tau = tau; m->ltau = log(tau);
}
int orc_hmm_M(const orc_model* m) { return m->M; }
int orc_hmm_S(const orc_model* m) { return m->S; }
int orc_hmm_nparam(const orc_model* m) { return m->row_off[m->nrow]; }
int orc_hmm_get(const orc_model* m, int kind, int* out) {
  int n = 0;
#define PUT(x) do { if (out) out[n] = (x); ++n; } while (0)
  const ivec* lists = kind == 2 ? m->right : kind == 3 ? m->left : m->pairt;
  switch (kind) {
    case 0: for (int s = 0; s < m->S; ++s) { PUT(m->state[s].l); PUT(m->state[s].r); } break;
    case 1: for (int k = 0; k < m->loop_state.n; ++k) PUT(m->loop_state.v[k]); break;
    case 2: case 3: case 4: {
      int o = 0; PUT(0);
      for (int s = 0; s < m->S; ++s) { o += lists[s].n; PUT(o); }
      for (int s = 0; s < m->S; ++s) for (int k = 0; k < lists[s].n; ++k) PUT(lists[s].v[k]);
      break;
    }
    case 5: for (int k = 0; k < 4 * m->nquad; ++k) PUT(m->quad[k]); break;
    case 6: for (int h = 0; h < m->M; ++h) PUT(m->node[h]); break;
    case 7: for (int h = 0; h < m->M; ++h) PUT(m->theta_id[h]); break;
    case 8: for (int r = 0; r < m->nrow; ++r) PUT(m->row_size[r]); break;
    case 9: for (int k = 0; k < m->M * m->M; ++k) PUT(m->reach[k]); break;
    default: return -1;
  }
#undef PUT
  return n;
}
int orc_energy_get(const orc_model* m, const char* name, double* out, int cap) {
  struct { const char* n; const double* p; int len; } ents[] = {
      {"hairpin", m->hairpin, 31}, {"mismatch_h", &m->mm_h[0][0][0], 175}, {"mismatch_i", &m->mm_i[0][0][0], 175},
      {"mismatch_m", &m->mm_m[0][0][0], 175}, {"mismatch_1ni", &m->mm_1ni[0][0][0], 175},
      {"mismatch_23i", &m->mm_23i[0][0][0], 175}, {"mismatch_ext", &m->mm_ext[0][0][0], 175},
      {"stack", &m->stack[0][0], 49}, {"bulge", m->bulge, 31}, {"term_au", &m->term_au, 1},
      {"int11", &m->int11[0][0][0][0], 1600}, {"int21", &m->int21[0][0][0][0][0], 8000},
      {"int22", &m->int22[0][0][0][0][0][0], 40000}, {"internal", m->internal, 31}, {"dangle5", &m->dangle5[0][0], 40},
      {"dangle3", &m->dangle3[0][0], 40}, {"ninio", m->ninio, 31}, {"mlintern", &m->mlintern, 1},
      {"mlclosing", &m->mlclosing, 1}, {"ml_base", &m->ml_base, 1}, {"lxc37", &m->lxc37, 1},
      {"triloop", m->triloop, m->ntri}, {"tetraloop", m->tetraloop, m->ntetra}, {"hexaloop", m->hexaloop, m->nhexa}};
  for (size_t k = 0; k < sizeof(ents) / sizeof(ents[0]); ++k)
    if (!strcmp(ents[k].n, name)) {
      int n = cap < ents[k].len ? cap : ents[k].len;
      for (int i = 0; i < n; ++i) out[i] = ents[k].p[i];
      return ents[k].len;
    }
  return -1;
}

/* motif_model.hpp:62-70 */
void orc_set_ws(const int* q, int n, double* ws) {
  int cnt[127 - 33]; memset(cnt, 0, sizeof(cnt));
  for (int i = 0; i < n; ++i) cnt[q[i]] += 1;
  int mode = 0, best = -2147483647 - 1;
  for (int i = 0; i < 127 - 33; ++i) if (best <= cnt[i]) { mode = i; best = cnt[i]; }
  for (int i = 0; i < n - 1; ++i) ws[i] = log((0.01 + (double)q[i]) / (0.01 + mode));
  ws[n - 1] = q[n - 1] == 0 ? NEGINF : 0.;
}

/* ------------------------------------------------------------------------------------- structural grammar */
#define BPOK(m, i, d) ((m)->bp_ok[(i) * ((m)->W + 1) + (d)])
#define LFOK(m, i, d) ((m)->left_ok[(i) * ((m)->W + 1) + (d)])
static int parsable(const orc_model* m, int e, int i, int j) { /* energy_model.hpp:289-338 */
  int d = j - i;
  switch (e) {
    case ST_P: return 0 <= i && d <= m->W && BPOK(m, i, d);
    case ST_E: return 0 < i && d + 2 <= m->W && BPOK(m, i - 1, d + 2);
    case ST_M: return 0 < i && j < m->L && d <= m->W && (m->no_turn ? 4 <= d : 10 <= d);
    case ST_B: case ST_1: case ST_2: return d <= m->W && LFOK(m, i, d);
  }
  return 0;
}
typedef void (*trans_fn)(orc_model*, int tt, int i, int j, int k, int l, double tsc);
typedef void (*col_fn)(orc_model*, int a, int b);

static int all_dots(const orc_model* m, int from, int n) {
  for (int k = 0; k < n; ++k) if (m->fix_s[from + k] != '.') return 0;
  return 1;
}
/* energy_model.hpp:340-441 */
static void compute_inside(orc_model* m, trans_fn f, col_fn before) {
  const int L = m->L, W = m->W, C = m->C;
  const int* s = m->seq;
  for (int j = 0; j <= L; ++j) {
    int i0 = j - W > 0 ? j - W : 0;
    if (before) before(m, i0, j);
    for (int i = j; i0 <= i; --i) {
      double tsc;
      if (parsable(m, ST_P, i, j)) {
        if (parsable(m, ST_E, i + 1, j - 1)) f(m, TT_P_E, i, j, i + 1, j - 1, 0.);
        if (parsable(m, ST_P, i + 1, j - 1)) {
          tsc = m->no_ene ? 0. : loop_energy(m, s, i, j - 1, i + 1, j - 2);
          if (tsc != NEGINF) f(m, TT_P_P, i, j, i + 1, j - 1, tsc);
        }
      }
      if (parsable(m, ST_B, i, j))
        for (int k = i; k <= j; ++k)
          if (parsable(m, ST_1, i, k) && parsable(m, ST_2, k, j)) f(m, TT_B_12, i, j, i, k, 0.);
      if (parsable(m, ST_2, i, j)) {
        if (parsable(m, ST_2, i, j - 1)) { if (m->fix_s && m->fix_s[j - 1] != '.') {} else f(m, TT_2_2, i, j, i, j - 1, 0.); }
        if (parsable(m, ST_P, i, j)) {
          tsc = m->no_ene ? 0. : sum_ext_m(m, s, L, i, j - 1, 0) + m->mlintern;
          if (tsc != NEGINF) f(m, TT_2_P, i, j, i, j, tsc);
        }
      }
      if (parsable(m, ST_1, i, j)) {
        if (parsable(m, ST_2, i, j)) f(m, TT_1_2, i, j, i, j, 0.);
        if (parsable(m, ST_B, i, j)) f(m, TT_1_B, i, j, i, j, 0.);
      }
      if (parsable(m, ST_M, i, j)) {
        if (parsable(m, ST_M, i + 1, j)) { if (m->fix_s && m->fix_s[i] != '.') {} else f(m, TT_M_M, i, j, i + 1, j, 0.); }
        if (parsable(m, ST_B, i, j)) f(m, TT_M_B, i, j, i, j, 0.);
      }
      if (parsable(m, ST_E, i, j)) {
        if (parsable(m, ST_M, i, j)) {
          tsc = m->no_ene ? 0. : sum_ext_m(m, s, L, j, i - 1, 0) + (m->mlclosing + m->mlintern);
          if (tsc != NEGINF) f(m, TT_E_M, i, j, i, j, tsc);
        }
        tsc = m->no_ene ? 0. : hairpin_energy(m, s, i - 1, j);
        if (m->fix_s && !all_dots(m, i, j - i)) {} else if (tsc != NEGINF) f(m, TT_E_H, i, j, i, j, tsc);
        for (int l = j; l >= (i > j - C ? i : j - C); --l)
          for (int k = i; k <= (l < i + C - (j - l) ? l : i + C - (j - l)); ++k) {
            if (i == k && l == j) continue;
            if (parsable(m, ST_P, k, l)) {
              tsc = m->no_ene ? 0. : loop_energy(m, s, i - 1, j, k, l - 1);
              if (m->fix_s && (!all_dots(m, i, k - i) || !all_dots(m, l, j - l))) {}
              else if (tsc != NEGINF) f(m, TT_E_P, i, j, k, l, tsc);
            }
          }
      }
      if (parsable(m, ST_P, i, j)) {
        tsc = m->no_ene ? 0. : sum_ext_m(m, s, L, i, j - 1, 1);
        if (tsc != NEGINF) f(m, TT_O_OP, 0, j, 0, i, tsc);
      }
      if (i0 == i && 0 < j) { if (m->fix_s && m->fix_s[j - 1] != '.') {} else f(m, TT_O_O, 0, j, 0, j - 1, 0.); }
    }
  }
}
/* energy_model.hpp:443-547 */
static void compute_outside(orc_model* m, trans_fn f, col_fn after) {
  const int L = m->L, W = m->W, C = m->C;
  const int* s = m->seq;
  for (int j = L; 0 <= j; --j) {
    int i0 = j - W > 0 ? j - W : 0;
    for (int i = i0; i <= j; ++i) {
      double tsc;
      if (i0 == i && j < L) { if (m->fix_s && m->fix_s[j] != '.') {} else f(m, TT_O_O, 0, j, 0, j + 1, 0.); }
      if (parsable(m, ST_2, i, j)) {
        if (parsable(m, ST_2, i, j + 1)) { if (m->fix_s && m->fix_s[j] != '.') {} else f(m, TT_2_2, i, j, i, j + 1, 0.); }
        if (parsable(m, ST_1, i, j)) f(m, TT_1_2, i, j, i, j, 0.);
      }
      if (parsable(m, ST_P, i, j)) {
        tsc = m->no_ene ? 0. : sum_ext_m(m, s, L, i, j - 1, 1);
        if (tsc != NEGINF) f(m, TT_O_OP, 0, i, 0, j, tsc);
        if (parsable(m, ST_P, i - 1, j + 1)) {
          tsc = m->no_ene ? 0. : loop_energy(m, s, i - 1, j, i, j - 1);
          if (tsc != NEGINF) f(m, TT_P_P, i, j, i - 1, j + 1, tsc);
        }
        if (parsable(m, ST_2, i, j)) {
          tsc = m->no_ene ? 0. : sum_ext_m(m, s, L, i, j - 1, 0) + m->mlintern;
          if (tsc != NEGINF) f(m, TT_2_P, i, j, i, j, tsc);
        }
      }
      if (parsable(m, ST_E, i, j)) {
        if (parsable(m, ST_P, i - 1, j + 1)) f(m, TT_P_E, i, j, i - 1, j + 1, 0.);
        tsc = m->no_ene ? 0. : hairpin_energy(m, s, i - 1, j);
        if (m->fix_s && !all_dots(m, i, j - i)) {} else if (tsc != NEGINF) f(m, TT_E_H, i, j, i, j, tsc);
        if (parsable(m, ST_M, i, j)) {
          tsc = m->no_ene ? 0. : sum_ext_m(m, s, L, j, i - 1, 0) + (m->mlclosing + m->mlintern);
          if (tsc != NEGINF) f(m, TT_E_M, i, j, i, j, tsc);
        }
      }
      if (parsable(m, ST_M, i, j) && parsable(m, ST_M, i - 1, j)) {
        if (m->fix_s && m->fix_s[i - 1] != '.') {} else f(m, TT_M_M, i, j, i - 1, j, 0.);
      }
      if (parsable(m, ST_B, i, j)) {
        if (parsable(m, ST_1, i, j)) f(m, TT_1_B, i, j, i, j, 0.);
        if (parsable(m, ST_M, i, j)) f(m, TT_M_B, i, j, i, j, 0.);
        for (int k = j; k >= i; --k)
          if (parsable(m, ST_1, i, k) && parsable(m, ST_2, k, j)) f(m, TT_B_12, i, k, i, j, 0.);
      }
      if (parsable(m, ST_E, i, j)) {
        /* the reference's inner bound is self-referential (energy_model.hpp:529): every l >= k+2 is visited */
        for (int k = i; k <= (j - 2 < i + C ? j - 2 : i + C); ++k)
          for (int l = j; l >= k + 2; --l) {
            if (i == k && l == j) continue;
            if (parsable(m, ST_P, k, l)) {
              tsc = m->no_ene ? 0. : loop_energy(m, s, i - 1, j, k, l - 1);
              if (m->fix_s && (!all_dots(m, i, k - i) || !all_dots(m, l, j - l))) {}
              else if (tsc != NEGINF) f(m, TT_E_P, k, l, i, j, tsc);
            }
          }
      }
    }
    if (1 <= j && after) after(m, i0, j - 1);
  }
}

/* ---- energy-only inside / outside (energy_model.hpp:559-661), tables [i][d][7] */
#define EIN(m, i, j, e) ((m)->ein[((i) * ((m)->W + 1) + ((j) - (i))) * 7 + (e)])
#define EOUT(m, i, j, e) ((m)->eout[((i) * ((m)->W + 1) + ((j) - (i))) * 7 + (e)])
static void e_inside(orc_model* m, int t, int i, int j, int k, int l, double tsc) {
  switch (t) {
    case TT_O_OP: addl(&m->ein_o[j], m->ein_o[l] + (EIN(m, l, j, ST_P) + tsc)); break;
    case TT_O_O: addl(&m->ein_o[j], m->ein_o[l] + tsc); break;
    case TT_E_H: addl(&EIN(m, i, j, ST_E), tsc); break;
    case TT_B_12: addl(&EIN(m, i, j, ST_B), EIN(m, k, l, ST_1) + (EIN(m, l, j, ST_2) + tsc)); break;
    default: addl(&EIN(m, i, j, TT_FIRST[t]), EIN(m, k, l, TT_SECOND[t]) + tsc);
  }
}
static void e_outside(orc_model* m, int t, int i, int j, int k, int l, double tsc) {
  switch (t) {
    case TT_O_OP:
      addl(&m->eout_o[j], EIN(m, j, l, ST_P) + (m->eout_o[l] + tsc));
      addl(&EOUT(m, j, l, ST_P), m->ein_o[j] + (m->eout_o[l] + tsc));
      break;
    case TT_O_O: addl(&m->eout_o[j], m->eout_o[l] + tsc); break;
    case TT_E_H: break;
    case TT_B_12:
      addl(&EOUT(m, i, j, ST_1), EIN(m, j, l, ST_2) + (EOUT(m, k, l, ST_B) + tsc));
      addl(&EOUT(m, j, l, ST_2), EIN(m, i, j, ST_1) + (EOUT(m, k, l, ST_B) + tsc));
      break;
    default: addl(&EOUT(m, i, j, TT_SECOND[t]), EOUT(m, k, l, TT_FIRST[t]) + tsc);
  }
}
static void fill_left(orc_model* m) { /* energy_model.hpp:203-209 */
  int L = m->L, W = m->W;
  memset(m->left_ok, 0, (size_t)(L + 1) * (W + 1));
  for (int i = 0; i <= L; ++i)
    for (int j = i + 1; j <= (L < i + W ? L : i + W); ++j)
      if (LFOK(m, i, j - i - 1) || BPOK(m, i, j - i)) LFOK(m, i, j - i) = 1;
}
static double ln_bpp(orc_model* m, int i, int j) { /* energy_model.hpp:195-201 */
  if (0 <= i && j <= m->L && j - i <= m->W && (m->no_turn ? 1 : 5 <= j - i))
    return (EIN(m, i, j, ST_P) + EOUT(m, i, j, ST_P)) - m->ein_o[m->L];
  return NEGINF;
}
/* EnergyModel::set_seq (energy_model.hpp:268-276) + fill_bpp_tables (211-266) */
static void set_seq(orc_model* m, const int* seq, int L, double* lnbpp_out) {
  m->seq = seq; m->L = L;
  m->W = L < m->max_span ? L : m->max_span;
  m->C = (m->W - 2 - (m->no_turn ? 2 : 5)) < m->max_iloop ? (m->W - 2 - (m->no_turn ? 2 : 5)) : m->max_iloop;
  int W = m->W;
  size_t nc = (size_t)(L + 1) * (W + 1);
  m->bp_ok = (unsigned char*)realloc(m->bp_ok, nc); m->left_ok = (unsigned char*)realloc(m->left_ok, nc);
  memset(m->bp_ok, 0, nc); memset(m->left_ok, 0, nc);
  int total = 0, nbp = 0;
  for (int i = 0; i <= L; ++i)
    for (int j = m->no_turn ? i + 1 : i + 5; j <= (L < i + W ? L : i + W); ++j)
      if ((BPOK(m, i, j - i) = 0 < BP[seq[i]][seq[j - 1]])) ++total;
  if (m->fix_s) {
    memset(m->bp_ok, 0, nc);
    int* stk = (int*)malloc(sizeof(int) * (L + 1)); int sp = 0;
    for (int i = 0; i < L; ++i) {
      if (m->fix_s[i] == '(') stk[sp++] = i;
      else if (m->fix_s[i] == ')') { int j = stk[--sp]; BPOK(m, j, i + 1 - j) = 1; ++nbp; }
    }
    free(stk);
  } else if (m->min_bpp == 0) {
    nbp = total;
  } else {
    fill_left(m);
    m->ein = (double*)realloc(m->ein, sizeof(double) * nc * 7); m->eout = (double*)realloc(m->eout, sizeof(double) * nc * 7);
    m->ein_o = (double*)realloc(m->ein_o, sizeof(double) * (L + 1)); m->eout_o = (double*)realloc(m->eout_o, sizeof(double) * (L + 1));
    for (size_t k = 0; k < nc * 7; ++k) m->ein[k] = m->eout[k] = NEGINF;
    for (int k = 0; k <= L; ++k) m->ein_o[k] = m->eout_o[k] = NEGINF;
    m->ein_o[0] = 0.; m->eout_o[L] = 0.;
    compute_inside(m, e_inside, NULL);
    compute_outside(m, e_outside, NULL);
    unsigned char* keep = (unsigned char*)calloc(nc, 1);
    for (int i = 0; i <= L; ++i)
      for (int j = m->no_turn ? i + 1 : i + 5; j <= (L < i + W ? L : i + W); ++j) {
        double v = ln_bpp(m, i, j);
        if (lnbpp_out && BPOK(m, i, j - i)) lnbpp_out[i * (W + 1) + (j - i)] = v;
        if ((keep[i * (W + 1) + (j - i)] = m->min_lnbpp <= v)) ++nbp;
      }
    memcpy(m->bp_ok, keep, nc);
    free(keep);
  }
  fill_left(m);
  m->bpp_eff = (double)nbp / (double)total;
}

double orc_bpp(orc_model* m, const int* seq, int L, unsigned char* bp_ok, unsigned char* left_ok, double* lnbpp, double* lnZ) {
  int W = L < m->max_span ? L : m->max_span;
  size_t nc = (size_t)(L + 1) * (W + 1);
  if (lnbpp) for (size_t k = 0; k < nc; ++k) lnbpp[k] = NEGINF;
  set_seq(m, seq, L, lnbpp);
  if (bp_ok) memcpy(bp_ok, m->bp_ok, nc);
  if (left_ok) memcpy(left_ok, m->left_ok, nc);
  if (lnZ) *lnZ = (m->min_bpp != 0 && !m->fix_s) ? m->ein_o[L] : NEGINF;
  return m->bpp_eff;
}

/* ------------------------------------------------------------------------------------------ coupled grammar */
#define TAB(T, m, i, j, e, s) ((T)[(((size_t)(i) * ((m)->W + 1) + ((j) - (i))) * 7 + (e)) * (m)->S + (s)])
#define OTAB(T, m, j, s) ((T)[(size_t)(j) * (m)->S + (s)])
#define N2S(m, a, b) ((m)->n2s[(a) * (m)->M + (b)])

static double theta1(const orc_model* m, int h, int b) { /* profile_hmm.hpp:137-141 */
  return (b == 0 || m->no_theta) ? 0. : m->theta[m->theta_id[h]][b - 1];
}
static double theta2(const orc_model* m, int h, int h1, int a, int b) { /* profile_hmm.hpp:113-135 */
  if (m->node[h1] == ')') return (BP[a][b] == 0 || m->no_theta) ? 0. : m->theta[m->theta_id[h1]][BP[a][b] - 1];
  if (m->no_theta) return 0.;
  return (a == 0 ? 0. : m->theta[m->theta_id[h]][a - 1]) + (b == 0 ? 0. : m->theta[m->theta_id[h1]][b - 1]);
}
static double weight(const orc_model* m, int h, int i) { /* motif_model.hpp:131-134 */
  int c = m->node[h];
  return (c == '.' || c == '(' || c == ')') ? m->ws[i] : 0.;
}
static double lam_of(const orc_model* m, int s) { return m->state[s].l == m->state[s].r ? m->lambda[0] : m->lambda[1]; }
static void add_emit2(const orc_model* m, double* e, int h, int h1, int a, int b, double w) { /* profile_hmm.hpp:144-179 */
  if (m->node[h1] == ')') { if (0 < BP[a][b]) e[m->row_off[m->theta_id[h1]] + BP[a][b] - 1] += w; }
  else {
    if (a != 0) e[m->row_off[m->theta_id[h]] + a - 1] += w;
    if (b != 0) e[m->row_off[m->theta_id[h1]] + b - 1] += w;
  }
}
static void add_emit1(const orc_model* m, double* e, int h, int b, double w) { if (b != 0) e[m->row_off[m->theta_id[h]] + b - 1] += w; }

/* constraint vetoes of InsideEndFun (motif_scanner.hpp:606-639) and CYKFun (:843-880); parent s, child s1 */
static int vetoed(const orc_model* m, int e, int i, int j, int k, int l, int s, int s1) {
  if (m->mode != MODE_SCAN_END && m->mode != MODE_CYK) return 0;
  IS a = m->state[s], b = s1 >= 0 ? m->state[s1] : a;
  int ys = m->Ys, ye = m->mode == MODE_CYK ? m->Ye : -2, M = m->M;
  switch (e) {
    case ST_P:
      if (i == k - 1 && l == j - 1) {
        if (i == ys && !(0 == a.l && 1 == b.l)) return 1;
        if (l == ys && !(0 == b.r && 1 == a.r)) return 1;
        if (m->mode == MODE_CYK) {
          if (i == ye && !(M - 2 == a.l && M - 1 == b.l)) return 1;
          if (l == ye && !(M - 2 == b.r && M - 1 == a.r)) return 1;
          if ((j == ye && m->L == j) && M - 2 != a.r) return 1;
        }
      }
      break;
    case ST_O: case ST_2: case ST_L:
      if (i == k && l == j - 1) {
        if (l == ys && !(0 == b.r && 1 == a.r)) return 1;
        if (m->mode == MODE_CYK) {
          if (l == ye && !(M - 2 == b.r && M - 1 == a.r)) return 1;
          if ((j == ye && m->L == j) && M - 2 != a.r) return 1;
        }
      }
      break;
    case ST_M:
      if (i == k - 1 && l == j) {
        if (i == ys && !(0 == a.l && 1 == b.l)) return 1;
        if (m->mode == MODE_CYK && i == ye && !(M - 2 == a.l && M - 1 == b.l)) return 1;
      }
      break;
  }
  return 0;
}

static void cyk_compare(orc_model* m, int e, int e1, int i, int j, int k, int l, int s, int s1, double* x, double y, int tt) {
  if (*x < y) { /* motif_scanner.hpp:815-826 */
    *x = y;
    Trace* t = e == ST_O ? &OTAB(m->trace_o, m, j, s) : &TAB(m->trace, m, i, j, e, s);
    t->k = k; t->l = l; t->t = tt; t->e1 = e1; t->s1 = s1;
  }
}
/* on_inside_transition of RNAelemTrainDP::InsideFun / RNAelemScanDP::{InsideFun,InsideEndFun,CYKFun} */
static void on_inside(orc_model* m, int tt, int e, int e1, int i, int j, int k, int l, int s, int s1, int s2, int s3,
                      double tsc, double wt, double lam) {
  m->cnt_motif += 1;
  if (vetoed(m, e, i, j, k, l, s, s1)) return;
  double diff = wt + lam * tsc;
  double* T = m->in; double* O = m->in_o;
  if (m->mode == MODE_CYK) {
    if (e == ST_E && e1 == ST_P)
      cyk_compare(m, e, e1, i, j, k, l, s, s1, &TAB(T, m, i, j, e, s),
                  TAB(T, m, k, l, e1, s1) + (TAB(T, m, i, k, ST_L, s2) + (TAB(T, m, l, j, ST_L, s3) + diff)), tt);
    else if (e == ST_O && e1 == ST_P)
      cyk_compare(m, e, e1, i, j, k, l, s, s1, &OTAB(O, m, j, s),
                  OTAB(O, m, k, N2S(m, m->state[s].l, m->state[s1].l)) + (TAB(T, m, k, l, e1, s1) + diff), tt);
    else if (e == ST_B && e1 == ST_1)
      cyk_compare(m, e, e1, i, j, k, l, s, s1, &TAB(T, m, i, j, e, s), TAB(T, m, k, l, ST_1, s1) + (TAB(T, m, l, j, ST_2, s2) + diff), tt);
    else if (e == ST_O && e1 == ST_O)
      cyk_compare(m, e, e1, i, j, k, l, s, s1, &OTAB(O, m, j, s), OTAB(O, m, l, s1) + diff, tt);
    else
      cyk_compare(m, e, e1, i, j, k, l, s, s1, &TAB(T, m, i, j, e, s), TAB(T, m, k, l, e1, s1) + diff, tt);
    return;
  }
  if (e == ST_E && e1 == ST_P)
    addl(&TAB(T, m, i, j, e, s), TAB(T, m, k, l, e1, s1) + (TAB(T, m, i, k, ST_L, s2) + (TAB(T, m, l, j, ST_L, s3) + diff)));
  else if (e == ST_O && e1 == ST_P) addl(&OTAB(O, m, j, s), OTAB(O, m, k, s2) + (TAB(T, m, k, l, e1, s1) + diff));
  else if (e == ST_B && e1 == ST_1) addl(&TAB(T, m, i, j, e, s), TAB(T, m, k, l, ST_1, s1) + (TAB(T, m, l, j, ST_2, s2) + diff));
  else if (e == ST_O && e1 == ST_O) addl(&OTAB(O, m, j, s), OTAB(O, m, l, s1) + diff);
  else addl(&TAB(T, m, i, j, e, s), TAB(T, m, k, l, e1, s1) + diff);
}

/* on_outside_transition: child (i,j,e,s), parent (k,l,e1,s1)  (motif_trainer.hpp:347-457, motif_scanner.hpp:438-579,693-800) */
static void on_outside(orc_model* m, int e, int e1, int i, int j, int k, int l, int s, int s1, int s2, int s3,
                       double tsc, double wt, double lam) {
  double diff = wt + lam * tsc;
  double *I = m->in, *IO = m->in_o, *X = m->out, *XO = m->out_o;
  const int* seq = m->seq;
  IS cs = m->state[s], ps = m->state[s1];
  int M = m->M;
  double inner = e == ST_O ? OTAB(IO, m, j, s) : TAB(I, m, i, j, e, s);
  double rest;
  if (e1 == ST_E && e == ST_P) rest = TAB(X, m, k, l, e1, s1) + (TAB(I, m, k, i, ST_L, s2) + TAB(I, m, j, l, ST_L, s3));
  else if (e1 == ST_O && e == ST_P) rest = OTAB(XO, m, l, s1) + OTAB(IO, m, i, s2);
  else if (e1 == ST_B && e == ST_1) rest = TAB(X, m, k, l, e1, s1) + TAB(I, m, j, l, ST_2, s2);
  else if (e1 == ST_O && e == ST_O) rest = OTAB(XO, m, l, s1);
  else rest = TAB(X, m, k, l, e1, s1);
  double z = (diff + (inner + rest)) - m->ZL;
  if (z == NEGINF) return;
  if (m->mode == MODE_TRAIN) {
    if (lam == m->lambda[0]) m->dEH[0] += tsc * exp(z); else m->dEH[1] += tsc * exp(z);
  }
  if (m->mode == MODE_SCAN_END) {
    int ys = m->Ys;
    switch (e1) {
      case ST_P:
        if (k == i - 1 && j == l - 1) {
          if (ys == k && (0 != ps.l || 1 != cs.l)) return;
          if (ys == j && (0 != cs.r || 1 != ps.r)) return;
          if (M - 2 == ps.l && M - 1 == cs.l) addl(&m->PyeL[k], z);
          if (M - 2 == cs.r && M - 1 == ps.r) addl(&m->PyeL[j], z);
          if (M - 2 == ps.r && m->L == l) addl(&m->PyeL[m->L], z);
        }
        break;
      case ST_O: case ST_2: case ST_L:
        if (i == k && j == l - 1) {
          if (ys == j && (0 != cs.r || 1 != ps.r)) return;
          if (M - 2 == cs.r && M - 1 == ps.r) addl(&m->PyeL[j], z);
          if (M - 2 == ps.r && m->L == l) addl(&m->PyeL[m->L], z);
        }
        break;
      case ST_M:
        if (k == i - 1 && j == l) {
          if (ys == k && (0 != ps.l || 1 != cs.l)) return;
          if (M - 2 == ps.l && M - 1 == cs.l) addl(&m->PyeL[k], z);
        }
        break;
    }
  } else {
    switch (e1) {
      case ST_P: if (k == i - 1 && j == l - 1 && !m->no_prf) add_emit2(m, m->dEN, cs.l, ps.r, seq[k], seq[j], exp(z)); break;
      case ST_2: case ST_O: case ST_L: if (k == i && j == l - 1 && !m->no_prf) add_emit1(m, m->dEN, ps.r, seq[j], exp(z)); break;
      case ST_M: if (k == i - 1 && j == l && !m->no_prf) add_emit1(m, m->dEN, cs.l, seq[k], exp(z)); break;
    }
  }
  if (e1 == ST_E && e == ST_P) {
    addl(&TAB(X, m, i, j, e, s), TAB(X, m, k, l, e1, s1) + (TAB(I, m, k, i, ST_L, s2) + (TAB(I, m, j, l, ST_L, s3) + diff)));
    addl(&TAB(X, m, k, i, ST_L, s2), TAB(X, m, k, l, e1, s1) + (TAB(I, m, i, j, e, s) + (TAB(I, m, j, l, ST_L, s3) + diff)));
    addl(&TAB(X, m, j, l, ST_L, s3), TAB(X, m, k, l, e1, s1) + (TAB(I, m, i, j, e, s) + (TAB(I, m, k, i, ST_L, s2) + diff)));
  } else if (e1 == ST_O && e == ST_P) {
    addl(&TAB(X, m, i, j, e, s), OTAB(XO, m, l, s1) + (OTAB(IO, m, i, s2) + diff));
    addl(&OTAB(XO, m, i, s2), OTAB(XO, m, l, s1) + (TAB(I, m, i, j, e, s) + diff));
  } else if (e1 == ST_B && e == ST_1) {
    addl(&TAB(X, m, i, j, e, s), TAB(X, m, k, l, e1, s1) + (TAB(I, m, j, l, ST_2, s2) + diff));
    addl(&TAB(X, m, j, l, ST_2, s2), TAB(I, m, i, j, e, s) + (TAB(X, m, k, l, e1, s1) + diff));
  } else if (e1 == ST_O && e == ST_O) addl(&OTAB(XO, m, j, s), OTAB(XO, m, l, s1) + diff);
  else addl(&TAB(X, m, i, j, e, s), TAB(X, m, k, l, e1, s1) + diff);
  if (m->mode == MODE_SCAN_START) {
    switch (e1) {
      case ST_P:
        if (k == i - 1 && j == l - 1) {
          if (0 == ps.l && 1 == cs.l) addl(&m->PysL[k], z);
          if (0 == cs.r && 1 == ps.r) addl(&m->PysL[j], z);
          if (0 != cs.l && M - 1 != cs.l) addl(&m->PyiL[k], z);
          if (0 != ps.r && M - 1 != ps.r) addl(&m->PyiL[j], z);
        }
        break;
      case ST_2: case ST_O: case ST_L:
        if (i == k && j == l - 1) {
          if (0 == cs.r && 1 == ps.r) addl(&m->PysL[j], z);
          if (0 != ps.r && M - 1 != ps.r) addl(&m->PyiL[j], z);
        }
        break;
      case ST_M:
        if (k == i - 1 && j == l) {
          if (0 == ps.l && 1 == cs.l) addl(&m->PysL[k], z);
          if (0 != cs.l && M - 1 != cs.l) addl(&m->PyiL[k], z);
        }
        break;
    }
  }
}

static double tau_if(const orc_model* m, int cond) { return cond ? m->ltau : 0.; }
static double theta1p(const orc_model* m, int h, int b) { return m->no_prf ? 0. : theta1(m, h, b); }

/* RNAelem::InsideFun::before_transition (motif_model.hpp:243-257) */
static void c_before(orc_model* m, int i0, int j) {
  for (int i = j - 1; i0 <= i; --i)
    for (int a = 0; a < m->loop_state.n; ++a) {
      int s = m->loop_state.v[a]; IS st = m->state[s]; double lam = lam_of(m, s);
      for (int b = 0; b < m->right[s].n; ++b) {
        int s1 = m->right[s].v[b];
        double w = theta1p(m, st.r, m->seq[j - 1]), ws = weight(m, st.r, j - 1);
        double t = tau_if(m, st.r == m->state[s1].r && '.' == m->node[st.r]);
        on_inside(m, TT_L_L, ST_L, ST_L, i, j, i, j - 1, s, s1, -1, -1, 0., w + (t + ws), lam);
      }
    }
}
/* RNAelem::InsideFun::on_transition (motif_model.hpp:259-422) */
static void c_inside(orc_model* m, int tt, int i, int j, int k, int l, double tsc) {
  const int* seq = m->seq; int S = m->S;
  m->cnt_struct += 1;
  switch (tt) {
    case TT_E_H:
      for (int a = 0; a < m->loop_state.n; ++a) { int s = m->loop_state.v[a]; on_inside(m, tt, ST_E, ST_L, i, j, k, l, s, s, -1, -1, tsc, 0., lam_of(m, s)); }
      break;
    case TT_P_E: case TT_P_P:
      for (int s = 0; s < S; ++s) {
        double lam = lam_of(m, s); IS st = m->state[s];
        for (int b = 0; b < m->pairt[s].n; ++b) {
          int s1 = m->pairt[s].v[b]; IS c = m->state[s1];
          int rp = tt == TT_P_E ? j - 1 : l;
          double w = m->no_prf ? 0. : theta2(m, c.l, st.r, seq[i], seq[rp]);
          double ws = weight(m, c.l, i) + weight(m, st.r, rp);
          double t = tau_if(m, st.r == c.r && ')' == m->node[c.r]);
          on_inside(m, tt, ST_P, tt == TT_P_E ? ST_E : ST_P, i, j, k, l, s, s1, -1, -1, tsc, w + (t + ws), lam);
        }
      }
      break;
    case TT_O_O: case TT_2_2:
      for (int s = 0; s < S; ++s) {
        double lam = lam_of(m, s); IS st = m->state[s];
        for (int b = 0; b < m->right[s].n; ++b) {
          int s1 = m->right[s].v[b];
          double w = theta1p(m, st.r, seq[l]), ws = weight(m, st.r, l);
          double t = tau_if(m, st.r == m->state[s1].r && '.' == m->node[st.r]);
          int e = tt == TT_O_O ? ST_O : ST_2;
          on_inside(m, tt, e, e, i, j, k, l, s, s1, -1, -1, tsc, w + (t + ws), lam);
        }
      }
      break;
    case TT_O_OP:
      for (int s = 0; s < S; ++s) {
        double lam = lam_of(m, s); IS st = m->state[s];
        for (int h = st.l; h <= st.r; ++h)
          if (m->reach[st.l * m->M + h] && m->reach[h * m->M + st.r])
            on_inside(m, tt, ST_O, ST_P, i, j, l, j, s, N2S(m, h, st.r), N2S(m, st.l, h), -1, tsc, 0., lam);
      }
      break;
    case TT_E_P:
      for (int a = 0; a < m->nquad; ++a) {
        const int* q = m->quad + 4 * a;
        on_inside(m, tt, ST_E, ST_P, i, j, k, l, q[0], q[1], q[2], q[3], tsc, 0., lam_of(m, q[0]));
      }
      break;
    case TT_E_M: for (int s = 0; s < S; ++s) on_inside(m, tt, ST_E, ST_M, i, j, k, l, s, s, -1, -1, tsc, 0., lam_of(m, s)); break;
    case TT_M_M:
      for (int s = 0; s < S; ++s) {
        double lam = lam_of(m, s); IS st = m->state[s];
        for (int b = 0; b < m->left[s].n; ++b) {
          int s1 = m->left[s].v[b]; IS c = m->state[s1];
          double w = theta1p(m, c.l, seq[i]), ws = weight(m, c.l, i);
          double t = tau_if(m, st.l == c.l && '.' == m->node[st.l]);
          on_inside(m, tt, ST_M, ST_M, i, j, k, l, s, s1, -1, -1, tsc, w + (t + ws), lam);
        }
      }
      break;
    case TT_M_B: for (int s = 0; s < S; ++s) on_inside(m, tt, ST_M, ST_B, i, j, k, l, s, s, -1, -1, tsc, 0., lam_of(m, s)); break;
    case TT_B_12:
      for (int s = 0; s < S; ++s) {
        double lam = lam_of(m, s); IS st = m->state[s];
        for (int h = st.l; h <= st.r; ++h) {
          if (!m->reach[st.l * m->M + h] || !m->reach[h * m->M + st.r]) continue;
          on_inside(m, tt, ST_B, ST_1, i, j, k, l, s, N2S(m, st.l, h), N2S(m, h, st.r), -1, tsc, 0., lam);
        }
      }
      break;
    case TT_2_P: for (int s = 0; s < S; ++s) on_inside(m, tt, ST_2, ST_P, i, j, k, l, s, s, -1, -1, tsc, 0., lam_of(m, s)); break;
    case TT_1_2: for (int s = 0; s < S; ++s) on_inside(m, tt, ST_1, ST_2, i, j, k, l, s, s, -1, -1, tsc, 0., lam_of(m, s)); break;
    case TT_1_B: for (int s = 0; s < S; ++s) on_inside(m, tt, ST_1, ST_B, i, j, k, l, s, s, -1, -1, tsc, 0., lam_of(m, s)); break;
  }
}
/* RNAelem::OutsideFun::after_transition (motif_model.hpp:433-447) */
static void c_after(orc_model* m, int j0, int j) {
  for (int i = j0; i <= j; ++i)
    for (int a = 0; a < m->loop_state.n; ++a) {
      int s1 = m->loop_state.v[a]; IS p = m->state[s1]; double lam = lam_of(m, s1);
      for (int b = 0; b < m->right[s1].n; ++b) {
        int s = m->right[s1].v[b];
        double w = theta1p(m, p.r, m->seq[j]), ws = weight(m, p.r, j);
        double t = tau_if(m, m->state[s].r == p.r && '.' == m->node[p.r]);
        on_outside(m, ST_L, ST_L, i, j, i, j + 1, s, s1, -1, -1, 0., w + (t + ws), lam);
      }
    }
}
/* RNAelem::OutsideFun::on_transition (motif_model.hpp:449-612) */
static void c_outside(orc_model* m, int tt, int i, int j, int k, int l, double tsc) {
  const int* seq = m->seq; int S = m->S;
  switch (tt) {
    case TT_E_H:
      for (int a = 0; a < m->loop_state.n; ++a) { int s = m->loop_state.v[a]; on_outside(m, ST_L, ST_E, i, j, k, l, s, s, -1, -1, tsc, 0., lam_of(m, s)); }
      break;
    case TT_P_E: case TT_P_P:
      for (int s1 = 0; s1 < S; ++s1) {
        double lam = lam_of(m, s1); IS p = m->state[s1];
        for (int b = 0; b < m->pairt[s1].n; ++b) {
          int s = m->pairt[s1].v[b]; IS c = m->state[s];
          double w = m->no_prf ? 0. : theta2(m, c.l, p.r, seq[k], seq[j]);
          double ws = weight(m, c.l, k) + weight(m, p.r, j);
          double t = tau_if(m, c.r == p.r && ')' == m->node[p.r]);
          on_outside(m, tt == TT_P_E ? ST_E : ST_P, ST_P, i, j, k, l, s, s1, -1, -1, tsc, w + (t + ws), lam);
        }
      }
      break;
    case TT_O_O: case TT_2_2:
      for (int s1 = 0; s1 < S; ++s1) {
        double lam = lam_of(m, s1); IS p = m->state[s1];
        for (int b = 0; b < m->right[s1].n; ++b) {
          int s = m->right[s1].v[b];
          double w = theta1p(m, p.r, seq[j]), ws = weight(m, p.r, j);
          double t = tau_if(m, m->state[s].r == p.r && '.' == m->node[p.r]);
          int e = tt == TT_O_O ? ST_O : ST_2;
          on_outside(m, e, e, i, j, k, l, s, s1, -1, -1, tsc, w + (t + ws), lam);
        }
      }
      break;
    case TT_O_OP:
      for (int s1 = 0; s1 < S; ++s1) {
        double lam = lam_of(m, s1); IS p = m->state[s1];
        for (int h = p.l; h <= p.r; ++h)
          if (m->reach[p.l * m->M + h] && m->reach[h * m->M + p.r])
            on_outside(m, ST_P, ST_O, j, l, i, l, N2S(m, h, p.r), s1, N2S(m, p.l, h), -1, tsc, 0., lam);
      }
      break;
    case TT_E_P:
      for (int a = 0; a < m->nquad; ++a) {
        const int* q = m->quad + 4 * a;
        on_outside(m, ST_P, ST_E, i, j, k, l, q[1], q[0], q[2], q[3], tsc, 0., lam_of(m, q[0]));
      }
      break;
    case TT_E_M: for (int s = 0; s < S; ++s) on_outside(m, ST_M, ST_E, i, j, k, l, s, s, -1, -1, tsc, 0., lam_of(m, s)); break;
    case TT_M_M:
      for (int s1 = 0; s1 < S; ++s1) {
        double lam = lam_of(m, s1); IS p = m->state[s1];
        for (int b = 0; b < m->left[s1].n; ++b) {
          int s = m->left[s1].v[b]; IS c = m->state[s];
          double w = theta1p(m, c.l, seq[k]), ws = weight(m, c.l, k);
          double t = tau_if(m, c.l == p.l && '.' == m->node[p.l]);
          on_outside(m, ST_M, ST_M, i, j, k, l, s, s1, -1, -1, tsc, w + (t + ws), lam);
        }
      }
      break;
    case TT_2_P: for (int s = 0; s < S; ++s) on_outside(m, ST_P, ST_2, i, j, k, l, s, s, -1, -1, tsc, 0., lam_of(m, s)); break;
    case TT_1_2: for (int s = 0; s < S; ++s) on_outside(m, ST_2, ST_1, i, j, k, l, s, s, -1, -1, tsc, 0., lam_of(m, s)); break;
    case TT_1_B: for (int s = 0; s < S; ++s) on_outside(m, ST_B, ST_1, i, j, k, l, s, s, -1, -1, tsc, 0., lam_of(m, s)); break;
    case TT_M_B: for (int s = 0; s < S; ++s) on_outside(m, ST_B, ST_M, i, j, k, l, s, s, -1, -1, tsc, 0., lam_of(m, s)); break;
    case TT_B_12:
      for (int s1 = 0; s1 < S; ++s1) {
        double lam = lam_of(m, s1); IS p = m->state[s1];
        for (int h = p.l; h <= p.r; ++h) {
          if (!m->reach[p.l * m->M + h] || !m->reach[h * m->M + p.r]) continue;
          on_outside(m, ST_1, ST_B, i, j, k, l, N2S(m, p.l, h), s1, N2S(m, h, p.r), -1, tsc, 0., lam);
        }
      }
      break;
  }
}

/* table set-up (motif_trainer.hpp:89-106) */
static void alloc_tables(orc_model* m) {
  long n = (long)(m->L + 1) * (m->W + 1) * 7 * m->S, no = (long)(m->L + 1) * m->S;
  if (n > m->tab_n) {
    m->in = (double*)realloc(m->in, sizeof(double) * n); m->out = (double*)realloc(m->out, sizeof(double) * n);
    m->trace = (Trace*)realloc(m->trace, sizeof(Trace) * n); m->tab_n = n;
  }
  if (no > m->otab_n) {
    m->in_o = (double*)realloc(m->in_o, sizeof(double) * no); m->out_o = (double*)realloc(m->out_o, sizeof(double) * no);
    m->trace_o = (Trace*)realloc(m->trace_o, sizeof(Trace) * no); m->otab_n = no;
  }
}
static void init_inside(orc_model* m) {
  long n = (long)(m->L + 1) * (m->W + 1) * 7 * m->S, no = (long)(m->L + 1) * m->S;
  for (long k = 0; k < n; ++k) m->in[k] = NEGINF;
  for (long k = 0; k < no; ++k) m->in_o[k] = NEGINF;
  for (int i = 0; i <= m->L; ++i) for (int h = 0; h < m->M; ++h) TAB(m->in, m, i, i, ST_L, N2S(m, h, h)) = 0.;
  OTAB(m->in_o, m, 0, N2S(m, 0, 0)) = 0.;
}
static void init_outside(orc_model* m, int ari, int nasi) {
  long n = (long)(m->L + 1) * (m->W + 1) * 7 * m->S, no = (long)(m->L + 1) * m->S;
  for (long k = 0; k < n; ++k) m->out[k] = NEGINF;
  for (long k = 0; k < no; ++k) m->out_o[k] = NEGINF;
  OTAB(m->out_o, m, m->L, N2S(m, 0, 0)) = nasi ? 0. : NEGINF;
  OTAB(m->out_o, m, m->L, N2S(m, 0, m->M - 1)) = ari ? 0. : NEGINF;
  OTAB(m->out_o, m, m->L, N2S(m, 0, m->M - 2)) = ari ? 0. : NEGINF;
}
static double part_func(const orc_model* m, int ari, int nasi) { /* motif_trainer.hpp:108-112 */
  double a = nasi ? OTAB(m->in_o, m, m->L, N2S(m, 0, 0)) : NEGINF;
  double b = ari ? OTAB(m->in_o, m, m->L, N2S(m, 0, m->M - 2)) : NEGINF;
  double c = ari ? OTAB(m->in_o, m, m->L, N2S(m, 0, m->M - 1)) : NEGINF;
  return lse(a, lse(b, c));
}

int orc_estep_seq(orc_model* m, const int* seq, int L, const double* ws, int restricted, int is_negative, double* Z,
                  double* Zx_out, double* ENo, double* ENx, double* EHo, double* EHx, double* bpp_eff) {
  /* motif_trainer.hpp:204-245 */
  set_seq(m, seq, L, NULL);
  m->ws = ws;
  alloc_tables(m);
  init_inside(m);
  m->mode = MODE_TRAIN; m->cnt_struct = m->cnt_motif = 0;
  compute_inside(m, c_inside, c_before);
  double Ztt = part_func(m, 1, 1), Ztf = part_func(m, 1, 0), Zft = part_func(m, 0, 1);
  Z[0] = Ztt; Z[1] = Ztf; Z[2] = Zft;
  if (bpp_eff) *bpp_eff = m->bpp_eff;
  int ok = is_negative ? isfinite(Ztt) : (isfinite(Ztt) && isfinite(Ztf));
  if (!ok) return 1;
  int np = orc_hmm_nparam(m);
  memset(ENo, 0, sizeof(double) * np); memset(ENx, 0, sizeof(double) * np);
  EHo[0] = EHo[1] = EHx[0] = EHx[1] = 0.;
  init_outside(m, 1, 1);
  m->ZL = Ztt; m->dEH = EHo; m->dEN = ENo;
  compute_outside(m, c_outside, c_after);
  int ari = restricted;
  if (restricted < 0) ari = !is_negative && !(NEGINF < ws[L]);
  double Zx;
  if (ari) { init_outside(m, 1, 0); Zx = Ztf; } else { init_outside(m, 0, 1); Zx = Zft; }
  m->ZL = Zx; m->dEH = EHx; m->dEN = ENx;
  compute_outside(m, c_outside, c_after);
  if (Zx_out) *Zx_out = Zx;
  return 0;
}

/* the reference's test fixture RNAelemDP::dp() (RNAelem-test/motif_test.hpp:26-34): inside, then one outside pass
 * with the normaliser ZL given by the caller (its tests pass oneL = 0 to get un-normalised counts) */
int orc_debug_eval(orc_model* m, const int* seq, int L, const double* ws, double ZL, double* pf, double* pf_out,
                   double* EN, double* EH) {
  set_seq(m, seq, L, NULL);
  m->ws = ws;
  alloc_tables(m);
  init_inside(m); init_outside(m, 1, 1);
  m->mode = MODE_TRAIN;
  compute_inside(m, c_inside, c_before);
  int np = orc_hmm_nparam(m);
  memset(EN, 0, sizeof(double) * np); EH[0] = EH[1] = 0.;
  m->ZL = ZL; m->dEH = EH; m->dEN = EN;
  compute_outside(m, c_outside, c_after);
  *pf = part_func(m, 1, 1);
  *pf_out = OTAB(m->out_o, m, 0, N2S(m, 0, 0));
  return 0;
}

static int max_index(const double* v, int n) { /* util.hpp:231-241 */
  int s = 0; double mx = -1.7976931348623157e308;
  for (int i = 0; i < n; ++i) if (mx <= v[i]) { s = i; mx = v[i]; }
  return s;
}

/* RNAelemScanDP::trace_back (motif_scanner.hpp:262-362) */
static void fill_chars(char* s, int from, int n, char c) { for (int k = 0; k < n; ++k) s[from + k] = c; }
static void trace_back(orc_model* m, int s0) {
  typedef struct { int i, j, e, s; } T2;
  int cap = 8 * (m->L + 4), sp = 0;
  T2* st = (T2*)malloc(sizeof(T2) * cap);
  st[sp++] = (T2){0, m->L, ST_O, s0};
  while (sp > 0) {
    T2 t2 = st[--sp];
    Trace t = t2.e == ST_O ? OTAB(m->trace_o, m, t2.j, t2.s) : TAB(m->trace, m, t2.i, t2.j, t2.e, t2.s);
    if (t.t < 0) continue;
    IS ps = m->state[t2.s], c = m->state[t.s1];
    switch (t.t) {
      case TT_L_L: m->psihat[t.l] = ps.r; st[sp++] = (T2){t.k, t.l, t.e1, t.s1}; break;
      case TT_O_O: m->psihat[t.l] = ps.r; m->rss[t.l] = 'O'; st[sp++] = (T2){t.k, t.l, t.e1, t.s1}; break;
      case TT_2_2: m->psihat[t.l] = ps.r; m->rss[t.l] = 'M'; st[sp++] = (T2){t.k, t.l, t.e1, t.s1}; break;
      case TT_E_H: fill_chars(m->rss, t2.i, t2.j - t2.i, 'H'); st[sp++] = (T2){t.k, t.l, t.e1, t2.s}; break;
      case TT_E_M: case TT_M_B: case TT_2_P: case TT_1_2: case TT_1_B: st[sp++] = (T2){t.k, t.l, t.e1, t2.s}; break;
      case TT_P_E: case TT_P_P:
        m->psihat[t2.i] = c.l; m->rss[t2.i] = 'L'; m->psihat[t.l] = ps.r; m->rss[t.l] = 'R';
        st[sp++] = (T2){t.k, t.l, t.e1, t.s1};
        break;
      case TT_O_OP:
        st[sp++] = (T2){t.k, t.l, t.e1, t.s1};
        st[sp++] = (T2){ps.l, t.k, ST_O, N2S(m, ps.l, c.l)};
        break;
      case TT_E_P: {
        int n1 = t2.j - t.l, n2 = t.k - t2.i;
        if (n1 == 0) fill_chars(m->rss, t2.i, n2, 'B');
        else if (n2 == 0) fill_chars(m->rss, t.l, n1, 'B');
        else { fill_chars(m->rss, t2.i, n2, 'I'); fill_chars(m->rss, t.l, n1, 'I'); }
        st[sp++] = (T2){t.l, t2.j, ST_L, N2S(m, c.r, ps.r)};
        st[sp++] = (T2){t2.i, t.k, ST_L, N2S(m, ps.l, c.l)};
        st[sp++] = (T2){t.k, t.l, t.e1, t.s1};
        break;
      }
      case TT_B_12:
        st[sp++] = (T2){t.l, t2.j, ST_2, N2S(m, c.r, ps.r)};
        st[sp++] = (T2){t.k, t.l, t.e1, t.s1};
        break;
      case TT_M_M: m->psihat[t2.i] = c.l; m->rss[t2.i] = 'M'; st[sp++] = (T2){t.k, t.l, ST_M, t.s1}; break;
    }
  }
  free(st);
}

int orc_scan_seq(orc_model* m, const int* seq, int L, const double* ws, double* PysL, double* PyeL, double* PyiL,
                 int* psihat, char* rss, int* Ys, int* Ye, double* exist_prob, double* EN) {
  /* motif_scanner.hpp:172-260 */
  set_seq(m, seq, L, NULL);
  m->ws = ws;
  alloc_tables(m);
  for (int k = 0; k < L; ++k) PysL[k] = PyiL[k] = NEGINF;
  for (int k = 0; k <= L; ++k) PyeL[k] = NEGINF;
  m->PysL = PysL; m->PyiL = PyiL; m->PyeL = PyeL; m->dEN = EN; m->dEH = NULL;
  init_inside(m); init_outside(m, 1, 1);
  m->mode = MODE_SCAN_START;
  compute_inside(m, c_inside, c_before);
  m->ZL = part_func(m, 1, 1);
  compute_outside(m, c_outside, c_after);
  m->Ys = max_index(PysL, L);
  init_inside(m); init_outside(m, 1, 1);
  m->mode = MODE_SCAN_END;
  compute_inside(m, c_inside, c_before);
  m->ZL = part_func(m, 1, 1);
  compute_outside(m, c_outside, c_after);
  m->Ye = max_index(PyeL, L + 1);
  *Ys = m->Ys; *Ye = m->Ye;
  double sum = NEGINF;
  for (int k = 0; k < L; ++k) addl(&sum, PysL[k]);
  *exist_prob = exp(sum);
  /* constrained Viterbi (calc_viterbi_alignment, :172-184) */
  init_inside(m);
  long n = (long)(L + 1) * (m->W + 1) * 7 * m->S, no = (long)(L + 1) * m->S;
  for (long k = 0; k < n; ++k) m->trace[k].t = -1;
  for (long k = 0; k < no; ++k) m->trace_o[k].t = -1;
  for (int k = 0; k < L; ++k) { psihat[k] = 0; rss[k] = ' '; }
  rss[L] = 0;
  m->psihat = psihat; m->rss = rss;
  m->mode = MODE_CYK;
  compute_inside(m, c_inside, c_before);
  int sa = N2S(m, 0, m->M - 2), sb = N2S(m, 0, m->M - 1);
  int s0 = OTAB(m->in_o, m, L, sa) < OTAB(m->in_o, m, L, sb) ? sb : sa;
  trace_back(m, s0);
  return 0;
}

long orc_last_table(const orc_model* m, int which, double* out) {
  long n = (long)(m->L + 1) * (m->W + 1) * 7 * m->S, no = (long)(m->L + 1) * m->S;
  const double* src = which == 0 ? m->in : which == 1 ? m->out : which == 2 ? m->in_o : m->out_o;
  long cnt = which < 2 ? n : no;
  if (out) {
    if (which < 2) { /* stored [i][d][e][s] already */ memcpy(out, src, sizeof(double) * cnt); }
    else memcpy(out, src, sizeof(double) * cnt);
  }
  return cnt;
}
void orc_last_counts(const orc_model* m, double* structural, double* motif_terms) {
  if (structural) *structural = m->cnt_struct;
  if (motif_terms) *motif_terms = m->cnt_motif;
}
