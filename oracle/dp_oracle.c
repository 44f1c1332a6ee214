/* TEST INFRASTRUCTURE ONLY -- see dp_oracle.h.  Scalar restatement of the reference's SIMD dynprog.
 * Written cell-by-cell and list-by-list (no striping, no vectors) on purpose: it shares no
 * structure with the CUDA kernels it checks.  File:line citations are into /root/reference/src. */
#include "dp_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <ctype.h>

/* ------------------------------------------------------------------------------------------------
 * Tables (dynprog.c:1007-1173)
 * ---------------------------------------------------------------------------------------------- */
static short pd[4][128][128];
static unsigned char cons[128][128];
static int nt2int[128];
static int use8p[4];
static int tables_ready = 0;

/* intron score arrays, dynprog_genome.c:135-187.  Index: 0 sense, 1 antisense, 2 either; [finalp] */
static int intron_score[3][2][256];

/* intron.h:11-33 */
#define LEFT_GT 0x21
#define LEFT_GC 0x10
#define LEFT_AT 0x08
#define LEFT_CT 0x06
#define RIGHT_AG 0x30
#define RIGHT_AC 0x0C
#define RIGHT_GC 0x02
#define RIGHT_AT 0x01
#define GTAG_FWD 0x20
#define GCAG_FWD 0x10
#define ATAC_FWD 0x08
#define GTAG_REV 0x04
#define GCAG_REV 0x02
#define ATAC_REV 0x01
#define NONINTRON 0x00

/* comp.h:6-11 */
#define MATCH_COMP '|'
#define DYNPROG_MATCH_COMP '*'
#define AMBIGUOUS_COMP ':'
#define MISMATCH_COMP ' '
#define INDEL_COMP '-'
#define SHORTGAP_COMP '~'

/* scores.h:5-10 */
#define T_MATCH 1
#define T_MISMATCH (-3)
#define T_OPEN (-3)
#define T_INDEL (-1)

#define MICROINTRON_LENGTH 9	/* pairpool.c:19 */
#define LAZY_INDEL 1		/* dynprog_simd.c / dynprog_end.c:97 */
#define INSERT_PAIRS 9		/* dynprog_cdna.c:40 */
#define NEG_INFINITY_32 (-32768)

static void set_sym (int A, int B, int score) {
  /* permute_cases, dynprog.c:903-970: all case combinations, both orders */
  int a = tolower(A), b = tolower(B), i;
  int xs[2], ys[2], x, y;
  xs[0] = A; xs[1] = a; ys[0] = B; ys[1] = b;
  for (x = 0; x < 2; x++) for (y = 0; y < 2; y++) {
    cons[xs[x]][ys[y]] = 1; cons[ys[y]][xs[x]] = 1;
    for (i = 0; i < 4; i++) { pd[i][xs[x]][ys[y]] = (short) score; pd[i][ys[y]][xs[x]] = (short) score; }
  }
}

void orc_init (void) {
  static const int mism[4] = {-3, -2, -1, -5};	/* dynprog.h:47-49, dynprog.c:104 */
  static const char *half[] = {"RAG","YTC","WAT","SGC","MAC","KGT",0};
  static const char *full[] = {"HATC","BGCT","VGAC","DGAT","NTCAG","XTCAG",0};
  int i, c1, c2, k, j;
  if (tables_ready) return;
  memset(pd,0,sizeof(pd)); memset(cons,0,sizeof(cons));
  for (j = 0; j < 128; j++) nt2int[j] = 4;
  nt2int['A'] = nt2int['a'] = 0; nt2int['C'] = nt2int['c'] = 1;
  nt2int['G'] = nt2int['g'] = 2; nt2int['T'] = nt2int['t'] = 3;
  for (i = 0; i < 4; i++) use8p[i] = (-128) / mism[i] - 1;		/* dynprog.c:1022-1025 */
  for (i = 0; i < 4; i++)
    for (c1 = 'A'; c1 <= 'z'; c1++) for (c2 = 'A'; c2 < 'z'; c2++) pd[i][c1][c2] = (short) mism[i];  /* :1068-1077 */
  for (c1 = 'A'; c1 < 'Z'; c1++) set_sym(c1,c1,3);			/* :1089 */
  set_sym('U','T',3);
  for (k = 0; half[k]; k++) for (j = 1; half[k][j]; j++) set_sym(half[k][0],half[k][j],1);
  for (k = 0; full[k]; k++) for (j = 1; full[k][j]; j++) set_sym(full[k][0],full[k][j],3);
  set_sym('N','N',3); set_sym('X','X',3);

  memset(intron_score,0,sizeof(intron_score));
  /* sense */
  intron_score[0][1][GTAG_FWD] = 16; intron_score[0][1][GCAG_FWD] = 10; intron_score[0][1][ATAC_FWD] = 8;
  intron_score[0][0][GTAG_FWD] = 14; intron_score[0][0][GCAG_FWD] = 8;  intron_score[0][0][ATAC_FWD] = 4;
  /* antisense */
  intron_score[1][1][GTAG_REV] = 16; intron_score[1][1][GCAG_REV] = 10; intron_score[1][1][ATAC_REV] = 8;
  intron_score[1][0][GTAG_REV] = 14; intron_score[1][0][GCAG_REV] = 8;  intron_score[1][0][ATAC_REV] = 4;
  /* either: asymmetric on purpose (dynprog_genome.c:172-184) */
  intron_score[2][1][GTAG_FWD] = 16; intron_score[2][1][GCAG_FWD] = 10; intron_score[2][1][ATAC_FWD] = 8;
  intron_score[2][1][GTAG_REV] = 14; intron_score[2][1][GCAG_REV] = 10; intron_score[2][1][ATAC_REV] = 8;
  intron_score[2][0][GTAG_FWD] = 16; intron_score[2][0][GCAG_FWD] = 8;  intron_score[2][0][ATAC_FWD] = 4;
  intron_score[2][0][GTAG_REV] = 14; intron_score[2][0][GCAG_REV] = 8;  intron_score[2][0][ATAC_REV] = 4;
  tables_ready = 1;
}

int orc_pairdistance (int mt, int a, int b) { orc_init(); return pd[mt][a & 127][b & 127]; }
int orc_consistent (int a, int b) { orc_init(); return cons[a & 127][b & 127]; }
int orc_use8p_size (int mt) { orc_init(); return use8p[mt]; }

void orc_compute_bands (int *lband, int *uband, int rlength, int glength, int extraband, int widebandp) {
  if (!widebandp) { *lband = extraband; *uband = extraband; }
  else if (glength >= rlength) { *uband = glength - rlength + extraband; *lband = extraband; }
  else { *lband = rlength - glength + extraband; *uband = extraband; }
}

long orc_cells (int kind, int rlength, int glength, int lband, int uband) {
  long n = 0; int r, c, lo, hi;
  if (kind == 0) {
    for (c = 0; c <= glength; c++) { lo = c - uband; if (lo < 0) lo = 0; hi = c + lband; if (hi > rlength) hi = rlength; if (hi >= lo) n += hi - lo + 1; }
  } else if (kind == 1) {
    for (r = 0; r <= rlength; r++) { hi = r + uband; if (hi > glength) hi = glength; if (hi >= r) n += hi - r + 1; }
  } else {
    for (c = 0; c <= glength; c++) { hi = c + lband; if (hi > rlength) hi = rlength; if (hi >= c) n += hi - c + 1; }
  }
  return n;
}

/* ------------------------------------------------------------------------------------------------
 * Fills
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int bits, NEG, POS;
  const char *rs, *gs, *ga;	/* forward arrays */
  int rlength, glength, revp, mt;
} seqs_t;

static int satv (const seqs_t *s, int v) { return v < s->NEG ? s->NEG : (v > s->POS ? s->POS : v); }
static int qch (const seqs_t *s, int r) { return (unsigned char) (s->revp ? s->rs[s->rlength - r] : s->rs[r-1]) & 127; }
static int gch (const seqs_t *s, int c) { return (unsigned char) (s->revp ? s->gs[s->glength - c] : s->gs[c-1]) & 127; }
static int ach (const seqs_t *s, int c) { return (unsigned char) (s->revp ? s->ga[s->glength - c] : s->ga[c-1]) & 127; }

/* profile for upper and full fills: raw query char vs genome class (dynprog_simd.c:4424-4449) */
static int S_up (const seqs_t *s, int r, int c) {
  static const char L[5] = {'A','C','G','T','N'};
  int q = (r == 0) ? 'N' : qch(s,r);
  int k = (c == 0) ? 4 : nt2int[gch(s,c)], ka = (c == 0) ? 4 : nt2int[ach(s,c)];
  int a = pd[s->mt][q][(int) L[k]], b = pd[s->mt][q][(int) L[ka]];
  return a > b ? a : b;
}

/* profile for lower fills: query class vs raw genome char (dynprog_simd.c:5459-5524, 8690) */
static int S_lo (const seqs_t *s, int r, int c) {
  static const char L[5] = {'A','C','G','T','N'};
  int k = (r == 0) ? 4 : nt2int[qch(s,r)];
  int a, b;
  if (c == 0) return pd[s->mt][(int) L[k]][s->bits == 8 ? 4 : 'N'];
  a = pd[s->mt][(int) L[k]][gch(s,c)]; b = pd[s->mt][(int) L[k]][ach(s,c)];
  return a > b ? a : b;
}

#define IDX(r,c) ((r)*G1+(c))

static void fill_upper (const seqs_t *s, int open, int extend, int uband, int late,
			short *H, signed char *dN, signed char *dE) {
  int r, c, G1 = s->glength + 1, E, T1, Hd, diag, chigh;
  for (r = 0; r <= s->rlength; r++) {
    E = s->NEG;
    chigh = r + uband; if (chigh > s->glength) chigh = s->glength;
    for (c = r; c <= chigh; c++) {
      diag = (c == 0) ? 0 : (r == 0 ? s->NEG : H[IDX(r-1,c-1)]);
      Hd = satv(s,diag + S_up(s,r,c));
      if (c == r) {
	E = s->NEG; H[IDX(r,c)] = (short) Hd; dN[IDX(r,c)] = 0; dE[IDX(r,c)] = 0;
      } else {
	T1 = satv(s,H[IDX(r,c-1)] + open);
	dE[IDX(r,c)] = (late ? E >= T1 : E > T1) ? -1 : 0;
	E = satv(s,(E > T1 ? E : T1) + extend);
	dN[IDX(r,c)] = (late ? E >= Hd : E > Hd) ? -1 : 0;
	H[IDX(r,c)] = (short) (Hd > E ? Hd : E);
      }
    }
  }
}

static void fill_lower (const seqs_t *s, int open, int extend, int lband, int late,
			short *H, signed char *dN, signed char *dE) {
  int r, c, G1 = s->glength + 1, E, T1, Hd, diag, rhigh;
  for (c = 0; c <= s->glength; c++) {
    E = s->NEG;
    rhigh = c + lband; if (rhigh > s->rlength) rhigh = s->rlength;
    for (r = c; r <= rhigh; r++) {
      diag = (r == 0) ? 0 : (c == 0 ? s->NEG : H[IDX(r-1,c-1)]);
      Hd = satv(s,diag + S_lo(s,r,c));
      if (r == c) {
	E = s->NEG; H[IDX(r,c)] = (short) Hd; dN[IDX(r,c)] = 0; dE[IDX(r,c)] = 0;
      } else {
	T1 = satv(s,H[IDX(r-1,c)] + open);
	dE[IDX(r,c)] = (late ? E >= T1 : E > T1) ? -1 : 0;
	E = satv(s,(E > T1 ? E : T1) + extend);
	dN[IDX(r,c)] = (late ? E >= Hd : E > Hd) ? -1 : 0;
	H[IDX(r,c)] = (short) (Hd > E ? Hd : E);
      }
    }
  }
}

/* Full fill, stripe-faithful (dynprog_simd.c:2987-3700 / 6562-7200): W rows per stripe, E vectorised,
 * F a scalar pass with an int carry FF[c] between stripes.  W = 32 (8-bit) / 16 (16-bit) = AVX2. */
static void fill_full (const seqs_t *s, int open, int extend, int lband, int uband, int late, int W,
		       short *H, signed char *dN, signed char *dE, signed char *dF) {
  int rlength = s->rlength, glength = s->glength, G1 = glength + 1;
  int NEG = s->NEG;
  int *FF = (int *) calloc(glength + 1,sizeof(int));
  /* a private stripe-sized copy of the matrix, so that out-of-band lanes (which the SIMD code does
     write) never touch the caller's dense planes */
  int rows = ((rlength + W) / W) * W;
  short *M = (short *) calloc((size_t) rows * G1,sizeof(short));
  int *El = (int *) malloc(W * sizeof(int)), *Hl = (int *) malloc(W * sizeof(int)), *Hs = (int *) malloc(W * sizeof(int));
  int rlo, rhigh, c, c0, chigh, l, r, Xprev, T1, sc, Hd, rlo_calc, rhigh_calc, cgap, last, score;
#define MI(r,c) ((r)*G1+(c))

  for (rlo = 0; rlo <= rlength; rlo += W) {
    rhigh = rlo + W - 1; if (rhigh > rlength) rhigh = rlength;
    c0 = rlo - lband; if (c0 < 0) c0 = 0;
    for (l = 0; l < W; l++) {
      El[l] = late ? NEG : NEG + 1;
      Hl[l] = NEG - open;			/* fits: open < 0 */
      if (s->bits == 8) Hl[l] = (signed char) Hl[l]; else Hl[l] = (short) Hl[l];
    }
    chigh = rhigh + uband; if (chigh > glength) chigh = glength;
    for (c = c0; c <= chigh; c++) {
      if (c == 0) Xprev = (rlo == 0) ? 0 : NEG;
      else Xprev = (rlo == 0) ? NEG : M[MI(rlo-1,c-1)];

      for (l = 0; l < W; l++) {
	r = rlo + l;
	T1 = satv(s,Hl[l] + open);
	if (r <= rlength && r >= c - uband && r <= c + lband) dE[IDX(r,c)] = (late ? El[l] >= T1 : El[l] > T1) ? -1 : 0;
	El[l] = satv(s,(El[l] > T1 ? El[l] : T1) + extend);
      }
      for (l = 0; l < W; l++) Hs[l] = (l == 0) ? Xprev : Hl[l-1];
      for (l = 0; l < W; l++) {
	r = rlo + l;
	if (c == 0) sc = (r == 0) ? 0 : NEG;		/* pairscores_col0, dynprog_simd.c:3128-3140 */
	else if (r <= rlength) sc = S_up(s,r,c);
	else sc = 0;					/* pad lanes: never reach the band */
	Hd = satv(s,Hs[l] + sc);
	if (r <= rlength && r >= c - uband && r <= c + lband) dN[IDX(r,c)] = (late ? El[l] >= Hd : El[l] > Hd) ? -1 : 0;
	Hl[l] = Hd > El[l] ? Hd : El[l];
	M[MI(r,c)] = (short) Hl[l];
      }

      /* F pass */
      rlo_calc = rlo; if (rlo_calc < c - uband) rlo_calc = c - uband;
      rhigh_calc = rhigh;
      if (rhigh >= c + lband) {
	rhigh_calc = c + lband;
	if (c > 0) {
	  M[MI(rhigh_calc,c)] = (short) satv(s,M[MI(rhigh_calc-1,c-1)] + S_up(s,rhigh_calc,c));
	  dE[IDX(rhigh_calc,c)] = 0; dN[IDX(rhigh_calc,c)] = 0;
	}
      }
      if (rlo == 0 || c >= rlo + uband) { cgap = NEG_INFINITY_32; last = NEG_INFINITY_32; }
      else { cgap = FF[c]; last = M[MI(rlo_calc-1,c)]; }

      r = rlo_calc;
      if (r == c - uband) {
	score = last + open; cgap = score + extend;
	last = M[MI(r,c)];
	r++;
      }
      for ( ; r <= rhigh_calc; r++) {
	score = last + open;
	if (late ? cgap >= score : cgap > score) { cgap += extend; dF[IDX(r,c)] = -2; }
	else cgap = score + extend;
	last = M[MI(r,c)];
	if (late ? cgap >= last : cgap > last) {
	  last = cgap;
	  M[MI(r,c)] = (short) ((cgap < NEG) ? NEG : (s->bits == 8 ? (signed char) cgap : (short) cgap));
	  dN[IDX(r,c)] = -2;
	}
      }
      FF[c] = cgap;
      for (l = 0; l < W; l++) Hl[l] = M[MI(rlo+l,c)];
    }
  }

  for (c = 0; c <= glength; c++) for (r = 0; r <= rlength; r++)
    if (r >= c - uband && r <= c + lband) H[IDX(r,c)] = M[MI(r,c)];
  free(Hs); free(Hl); free(El); free(M); free(FF);
#undef MI
}

/* Algorithmic in-band cells of the fills run since the last orc_cells_reset(): how the CPU arms of
   bench.py count their work without the product library.  In count-only mode the entry points follow their control
   flow up to the fills (the no-fill shortcuts count zero cells), count, and return -1 without filling anything. */
static long cells_filled = 0;
static int count_only = 0;
void orc_cells_reset (void) { cells_filled = 0; }
long orc_cells_filled (void) { return cells_filled; }
void orc_set_count_only (int on) { count_only = on; }

int orc_fill (int kind, int bits, const char *rseq, const char *gseq, const char *galt,
	      int rlength, int glength, int mismatchtype, int open, int extend,
	      int lband, int uband, int jump_late_p, int revp,
	      short *H, signed char *dN, signed char *dE, signed char *dF) {
  seqs_t s;
  orc_init();
  cells_filled += orc_cells(kind,rlength,glength,lband,uband);
  if (count_only) return 0;
  s.bits = bits; s.NEG = bits == 8 ? -128 : -32768; s.POS = bits == 8 ? 127 : 32767;
  s.rs = rseq; s.gs = gseq; s.ga = galt; s.rlength = rlength; s.glength = glength; s.revp = revp; s.mt = mismatchtype;
  if (kind == 0) fill_full(&s,open,extend,lband,uband,jump_late_p,bits == 8 ? 32 : 16,H,dN,dE,dF);
  else if (kind == 1) fill_upper(&s,open,extend,uband,jump_late_p,H,dN,dE);
  else fill_lower(&s,open,extend,lband,jump_late_p,H,dN,dE);
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Pair lists.  The reference conses onto the head; here a list is a vector in PUSH order, so the
 * reference's list (head first) is the vector read backwards.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { orc_pair *v; int n, cap; } plist;

static void pl_init (plist *l) { l->v = NULL; l->n = 0; l->cap = 0; }
static void pl_free (plist *l) { free(l->v); l->v = NULL; l->n = l->cap = 0; }
static orc_pair *pl_new (plist *l) {
  if (l->n == l->cap) { l->cap = l->cap ? 2*l->cap : 256; l->v = (orc_pair *) realloc(l->v,l->cap * sizeof(orc_pair)); }
  memset(&l->v[l->n],0,sizeof(orc_pair));
  return &l->v[l->n++];
}
static void pl_reverse (plist *l) {
  int i, j; orc_pair t;
  for (i = 0, j = l->n - 1; i < j; i++, j--) { t = l->v[i]; l->v[i] = l->v[j]; l->v[j] = t; }
}

/* Pairpool_push, pairpool.c:180 */
static void push (plist *l, int querypos, int genomepos, char cdna, char comp, char genome, char genomealt, int dynprogindex) {
  orc_pair *p;
  if (querypos < 0 || genomepos < 0) return;
  p = pl_new(l);
  p->querypos = querypos; p->genomepos = genomepos; p->cdna = cdna; p->comp = comp;
  p->genome = genome; p->genomealt = genomealt; p->dynprogindex = dynprogindex;
  p->introntype = NONINTRON;
}

/* Pairpool_push_gapholder with NULL left/right pairs, pairpool.c:375 */
static orc_pair *push_gapholder (plist *l, int queryjump, int genomejump) {
  orc_pair *p = pl_new(l);
  p->querypos = -1; p->genomepos = -1; p->cdna = p->comp = p->genome = p->genomealt = ' ';
  p->gapp = 1; p->queryjump = queryjump; p->genomejump = genomejump; p->introntype = NONINTRON;
  return p;
}

/* One traceback context: pointers use the reference's convention (rev pointers address the last char) */
typedef struct {
  const char *rseq, *rsequc, *gseq, *galt;
  int queryoffset, genomeoffset, revp, dynprogindex;
  int score, nmatches, nmismatches, nopens, nindels;
} tb_t;

/* Pairpool_add_queryskip, pairpool.c:981 */
static void add_queryskip (plist *l, tb_t *t, int r, int c, int dist) {
  int querycoord = r - 1, genomecoord = c - 1, step = -1, j;
  if (t->revp) { querycoord = -querycoord; genomecoord = -genomecoord; step = +1; }
  for (j = 0; j < dist; j++) {
    push(l,t->queryoffset + querycoord,t->genomeoffset + genomecoord,t->rseq[querycoord],INDEL_COMP,' ',' ',t->dynprogindex);
    querycoord += step;
  }
}

/* Pairpool_add_genomeskip, pairpool.c:1068 (genomesequence == NULL path: chars come from the genome,
 * which inside the fetched segment are the segment's chars) */
static int add_genomeskip (plist *l, tb_t *t, int r, int c, int dist) {
  int querycoord = r - 1, left = c - dist, right = c - 1, step = -1, tmp, genomecoord, j;
  if (t->revp) { querycoord = -querycoord; tmp = left; left = -right; right = -tmp; step = +1; }
  if (dist < MICROINTRON_LENGTH) {
    genomecoord = t->revp ? left : right;
    for (j = 0; j < dist; j++) {
      push(l,t->queryoffset + querycoord,t->genomeoffset + genomecoord,' ',INDEL_COMP,t->gseq[genomecoord],t->galt[genomecoord],t->dynprogindex);
      genomecoord += step;
    }
    return 1;
  } else {
    push_gapholder(l,0,dist);
    return 0;
  }
}

/* the DIAG step shared by every traceback (e.g. dynprog_simd.c:9222-9268) */
static void diag_step (plist *l, tb_t *t, int r, int c) {
  int querycoord = r - 1, genomecoord = c - 1;
  char c1, c1_uc, c2, c2_alt;
  if (t->revp) { querycoord = -querycoord; genomecoord = -genomecoord; }
  c1 = t->rseq[querycoord]; c1_uc = t->rsequc[querycoord];
  c2 = t->gseq[genomecoord]; c2_alt = t->galt[genomecoord];
  if (c2 == '*') {
    /* nothing */
  } else if (c1_uc == c2 || c1_uc == c2_alt) {
    t->score += T_MATCH; t->nmatches++;
    push(l,t->queryoffset + querycoord,t->genomeoffset + genomecoord,c1,DYNPROG_MATCH_COMP,c2,c2_alt,t->dynprogindex);
  } else if (cons[(unsigned char) c1_uc & 127][(unsigned char) c2 & 127] || cons[(unsigned char) c1_uc & 127][(unsigned char) c2_alt & 127]) {
    t->score += T_MATCH; t->nmatches++;
    push(l,t->queryoffset + querycoord,t->genomeoffset + genomecoord,c1,AMBIGUOUS_COMP,c2,c2_alt,t->dynprogindex);
  } else {
    t->score += T_MISMATCH; t->nmismatches++;
    push(l,t->queryoffset + querycoord,t->genomeoffset + genomecoord,c1,MISMATCH_COMP,c2,c2_alt,t->dynprogindex);
  }
}

static void hgap (plist *l, tb_t *t, int r, int c, int dist) {
  if (add_genomeskip(l,t,r,c,dist)) { t->score += T_OPEN + dist*T_INDEL; t->nopens++; t->nindels += dist; }
}
static void vgap (plist *l, tb_t *t, int r, int c, int dist) {
  add_queryskip(l,t,r,c,dist);
  t->score += T_OPEN + dist*T_INDEL; t->nopens++; t->nindels += dist;
}

/* Dynprog_traceback_{8,16}, dynprog_simd.c:9154/9553 */
static void traceback_full (plist *l, tb_t *t, const signed char *dN, const signed char *dE, const signed char *dF,
			    int G1, int r, int c) {
  int dist, dir;
  while (r > 0 && c > 0) {
    dir = dN[IDX(r,c)];
    if (dir == -1) {
      dist = 1;
      while (c > 0 && dE[IDX(r,c--)] != 0) dist++;
      hgap(l,t,r,c+dist,dist);
    } else if (dir == -2) {
      dist = 1;
      while (r > 0 && dF[IDX(r--,c)] != 0) dist++;
      vgap(l,t,r+dist,c,dist);
    } else {
      diag_step(l,t,r,c); r--; c--;
    }
  }
  if (r == 0 && c == 0) {
  } else if (c == 0) vgap(l,t,r,0+LAZY_INDEL,r);
  else hgap(l,t,0+LAZY_INDEL,c,c);
}

/* Dynprog_traceback_{8,16}_upper, dynprog_simd.c:9319/9716 */
static void traceback_upper (plist *l, tb_t *t, const signed char *dN, const signed char *dE, int G1, int r, int c) {
  int dist;
  while (r > 0 && c > 0) {
    if (dN[IDX(r,c)] != 0) {
      dist = 1;
      while (dE[IDX(r,c--)] != 0) dist++;
      hgap(l,t,r,c+dist,dist);
    } else { diag_step(l,t,r,c); r--; c--; }
  }
  if (c != 0) hgap(l,t,0+LAZY_INDEL,c,c);
}

/* Dynprog_traceback_{8,16}_lower, dynprog_simd.c:9439/9836 */
static void traceback_lower (plist *l, tb_t *t, const signed char *dN, const signed char *dE, int G1, int r, int c) {
  int dist;
  while (r > 0 && c > 0) {
    if (dN[IDX(r,c)] != 0) {
      dist = 1;
      while (dE[IDX(r--,c)] != 0) dist++;
      vgap(l,t,r+dist,c,dist);
    } else { diag_step(l,t,r,c); r--; c--; }
  }
  if (r != 0) vgap(l,t,r,0+LAZY_INDEL,r);
}

/* traceback_nogaps, dynprog_end.c:649 */
static void traceback_nogaps (plist *l, tb_t *t, int r, int c) {
  while (r > 0 && c > 0) { diag_step(l,t,r,c); r--; c--; }
}

/* matrices for one upper+lower pair */
typedef struct {
  short *Hu, *Hl; signed char *dNu, *dEu, *dNl, *dEl;
  int G1;
} ul_t;

static void ul_alloc (ul_t *m, int rlength, int glength) {
  size_t n = (size_t) (rlength + 1) * (glength + 1);
  m->G1 = glength + 1;
  m->Hu = (short *) calloc(n,sizeof(short)); m->Hl = (short *) calloc(n,sizeof(short));
  m->dNu = (signed char *) calloc(n,1); m->dEu = (signed char *) calloc(n,1);
  m->dNl = (signed char *) calloc(n,1); m->dEl = (signed char *) calloc(n,1);
}
static void ul_free (ul_t *m) { free(m->Hu); free(m->Hl); free(m->dNu); free(m->dEu); free(m->dNl); free(m->dEl); }

static int dump (plist *l, int headfirst_is_reverse, orc_pair *pairs, int maxpairs) {
  /* headfirst_is_reverse: 1 if the reference returns the consed list as is (head = last pushed),
     0 if it returns List_reverse (push order) */
  int i, n = l->n;
  if (n == 0) return -1;
  for (i = 0; i < n && i < maxpairs; i++) pairs[i] = headfirst_is_reverse ? l->v[n-1-i] : l->v[i];
  return n;
}

static void bump (int *dynprogindex) { *dynprogindex += (*dynprogindex > 0 ? +1 : -1); }

/* --indel-open / --indel-extend: the user_open / user_extend / user_dynprog_p statics that Dynprog_single_setup,
   Dynprog_genome_setup and Dynprog_end_setup store (dynprog_single.c:101, dynprog_genome.c:192, dynprog_end.c:120);
   they override the defect_rate table in single, genome and end gaps (dynprog_single.c:470, dynprog_genome.c:3367,
   dynprog_end.c:1334,1964), not in cdna gaps */
static int user_open = 0, user_extend = 0, user_dynprog_p = 0;
void orc_set_user_dynprog (int open, int extend, int enabled) { user_open = open; user_extend = extend; user_dynprog_p = enabled; }

static void penalties (double defect_rate, const int *opens, const int *extends, int *open, int *extend) {
  int q = defect_rate < 0.003 ? 0 : (defect_rate < 0.014 ? 1 : 2);	/* dynprog.h:57-58 */
  if (user_dynprog_p) { *open = user_open; *extend = user_extend; return; }
  *open = opens[q]; *extend = extends[q];
}
static int mismatchtype_of (double defect_rate) { return defect_rate < 0.003 ? ORC_HIGHQ : (defect_rate < 0.014 ? ORC_MEDQ : ORC_LOWQ); }

/* ------------------------------------------------------------------------------------------------
 * Dynprog_single_gap, dynprog_single.c:428
 * ---------------------------------------------------------------------------------------------- */
int orc_single_gap (int *iout, const char *queryseq, const char *queryuc,
		    int rlength, int glength, int roffset, int goffset,
		    const char *gseg, const char *gseg_alt, int jump_late_p,
		    int extraband_single, int widebandp, double defect_rate,
		    int max_rlength, int max_glength, orc_pair *pairs, int maxpairs) {
  static const int opens[3] = {-8,-7,-6}, extends[3] = {-3,-2,-1};
  const char *rsequence = queryseq + roffset, *rsequenceuc = queryuc + roffset;
  int mt = mismatchtype_of(defect_rate), open, extend, lband, uband, bits, r, n;
  plist l; tb_t t;
  size_t ncell;
  short *H; signed char *dN, *dE, *dF;

  orc_init();
  penalties(defect_rate,opens,extends,&open,&extend);
  if (rlength <= 0 || glength <= 0 || rlength > max_rlength || glength > max_glength) {
    iout[1] = NEG_INFINITY_32; iout[2] = iout[3] = iout[4] = iout[5] = 0; bump(&iout[0]); return -1;
  }
  if (gseg[0] == '\0') { iout[1] = NEG_INFINITY_32; iout[2] = iout[3] = iout[4] = iout[5] = 0; return -1; }

  pl_init(&l);
  memset(&t,0,sizeof(t));
  t.rseq = rsequence; t.rsequc = rsequenceuc; t.gseq = gseg; t.galt = gseg_alt;
  t.queryoffset = roffset; t.genomeoffset = goffset; t.revp = 0; t.dynprogindex = iout[0];

  if (glength == rlength) {
    /* single_gap_simple, dynprog_single.c:345 */
    for (r = 1; r <= rlength; r++) diag_step(&l,&t,r,r);
    if (t.nmismatches <= 1 && l.n > 0) {
      iout[1] = t.score; iout[2] = t.nmatches; iout[3] = t.nmismatches; iout[4] = iout[5] = 0;
      bump(&iout[0]);
      n = dump(&l,1,pairs,maxpairs); pl_free(&l); return n;	/* consed left to right: head = last pushed */
    }
    l.n = 0; t.score = t.nmatches = t.nmismatches = 0;
  }

  orc_compute_bands(&lband,&uband,rlength,glength,extraband_single,widebandp);
  bits = (rlength < use8p[mt] && glength < use8p[mt]) ? 8 : 16;
  ncell = (size_t) (rlength + 1) * (glength + 1);
  H = (short *) calloc(ncell,sizeof(short)); dN = (signed char *) calloc(ncell,1);
  dE = (signed char *) calloc(ncell,1); dF = (signed char *) calloc(ncell,1);
  orc_fill(0,bits,rsequence,gseg,gseg_alt,rlength,glength,mt,open,extend,lband,uband,jump_late_p,0,H,dN,dE,dF);
  if (count_only) { free(H); free(dN); free(dE); free(dF); return -1; }
  traceback_full(&l,&t,dN,dE,dF,glength+1,rlength,glength);
  free(H); free(dN); free(dE); free(dF);
  iout[1] = t.score; iout[2] = t.nmatches; iout[3] = t.nmismatches; iout[4] = t.nopens; iout[5] = t.nindels;
  bump(&iout[0]);
  n = dump(&l,0,pairs,maxpairs); pl_free(&l); return n;
}

/* ------------------------------------------------------------------------------------------------
 * Dynprog_end5_gap / Dynprog_end3_gap, dynprog_end.c:1293/1924
 * ---------------------------------------------------------------------------------------------- */
static void best_endpoint (int *finalscore, int *bestr, int *bestc, const ul_t *m, int rlength, int glength,
			   int lband, int uband, int late, int lastrow_only, int NEG) {
  /* find_best_endpoint_{8,16} dynprog_end.c:143/220 ; _to_queryend_indels_{8,16} :358/437 */
  int G1 = m->G1, r, c, clo, chigh, best, v, r0;
  if (lastrow_only) { best = NEG; *bestr = rlength; *bestc = 0; r0 = rlength; }
  else { best = 0; *bestr = *bestc = 0; r0 = 1; }
  for (r = r0; r <= rlength; r++) {
    clo = r - lband; if (clo < 1) clo = 1;
    chigh = r + uband; if (chigh > glength) chigh = glength;
    for (c = clo; c < r; c++) {
      v = m->Hl[IDX(r,c)];
      if (late ? v >= best : v > best) { *bestr = r; *bestc = c; best = v; }
    }
    for ( ; c <= chigh; c++) {
      v = m->Hu[IDX(r,c)];
      if (late ? v >= best : v > best) { *bestr = r; *bestc = c; best = v; }
    }
  }
  *finalscore = best;
}

int orc_end_gap (int end5p, int *iout, const char *queryseq, const char *queryuc,
		 int rlength, int glength, int roffset, int goffset,
		 const char *gseg, const char *gseg_alt, int jump_late_p,
		 int extraband_end, double defect_rate, int endalign, int require_pos_score_p,
		 int max_rlength, int max_glength, orc_pair *pairs, int maxpairs) {
  static const int opens[3] = {-10,-8,-6}, extends[3] = {-2,-2,-2};
  int open, extend, lband, uband, bits, late, bestr = 0, bestc = 0, finalscore, n, i, k;
  plist l; tb_t t; ul_t m; int have_m = 0;
  const char *rs_fwd, *gs_fwd, *ga_fwd;

  orc_init();
  penalties(defect_rate,opens,extends,&open,&extend);
#define END_NULL() do { iout[1] = 0; iout[2] = iout[3] = iout[4] = iout[5] = 0; return -1; } while (0)
  if (rlength <= 0) END_NULL();
  if (endalign != ORC_QUERYEND_NOGAPS && rlength > max_rlength) rlength = max_rlength;
  if (end5p && goffset < 0) END_NULL();
  if (glength <= 0) END_NULL();
  if (endalign != ORC_QUERYEND_NOGAPS && glength > max_glength) glength = max_glength;
  if (gseg[0] == '\0') END_NULL();

  /* forward arrays covering the box, for orc_fill's (array, revp) convention.  5': the query is read
     backwards from queryseq[roffset]; the segment array ends at genome position goffset. */
  if (end5p) { rs_fwd = queryseq + roffset - (rlength - 1); gs_fwd = gseg; ga_fwd = gseg_alt; late = !jump_late_p; }
  else { rs_fwd = queryuc + roffset; gs_fwd = gseg; ga_fwd = gseg_alt; late = jump_late_p; }

  memset(&t,0,sizeof(t));
  if (end5p) { t.rseq = queryseq + roffset; t.rsequc = queryuc + roffset; t.gseq = gseg + glength - 1; t.galt = gseg_alt + glength - 1; t.revp = 1; }
  else { t.rseq = queryseq + roffset; t.rsequc = queryuc + roffset; t.gseq = gseg; t.galt = gseg_alt; t.revp = 0; }
  t.queryoffset = roffset; t.genomeoffset = goffset; t.dynprogindex = iout[0];

  if (endalign == ORC_QUERYEND_GAP || endalign == ORC_BEST_LOCAL || endalign == ORC_QUERYEND_INDELS) {
    int wide = (endalign != ORC_QUERYEND_INDELS);
    orc_compute_bands(&lband,&uband,rlength,glength,extraband_end,wide);
    bits = (rlength < use8p[ORC_ENDQ] || glength < use8p[ORC_ENDQ]) ? 8 : 16;
    const long cells_before = cells_filled;
    ul_alloc(&m,rlength,glength); have_m = 1;
    orc_fill(1,bits,rs_fwd,gs_fwd,ga_fwd,rlength,glength,ORC_ENDQ,open,extend,lband,uband,late,end5p,m.Hu,m.dNu,m.dEu,NULL);
    orc_fill(2,bits,rs_fwd,gs_fwd,ga_fwd,rlength,glength,ORC_ENDQ,open,extend,lband,uband,late,end5p,m.Hl,m.dNl,m.dEl,NULL);
    /* require_pos_score_p: the reference runs these fills and then throws the result away (dynprog_end.c:1558-1574);
       SURVEY.md section 8d counts such calls with zero cells */
    if (require_pos_score_p) cells_filled = cells_before;
    if (count_only) { ul_free(&m); return -1; }
    best_endpoint(&finalscore,&bestr,&bestc,&m,rlength,glength,lband,uband,late,!wide,bits == 8 ? -128 : -32768);
  } else {
    bestr = bestc = (glength < rlength) ? glength : rlength;	/* dynprog_end.c:577 */
  }

  pl_init(&l);
  if (endalign == ORC_QUERYEND_NOGAPS) traceback_nogaps(&l,&t,bestr,bestc);
  else if (require_pos_score_p) { /* traceback skipped: *traceback_score was just zeroed, dynprog_end.c:1558-1574 */ }
  else if (bestc >= bestr) traceback_upper(&l,&t,m.dNu,m.dEu,m.G1,bestr,bestc);
  else traceback_lower(&l,&t,m.dNl,m.dEl,m.G1,bestr,bestc);
  if (have_m) ul_free(&m);

  iout[1] = t.score; iout[2] = t.nmatches; iout[3] = t.nmismatches; iout[4] = t.nopens; iout[5] = t.nindels;
  if ((endalign == ORC_QUERYEND_GAP || endalign == ORC_BEST_LOCAL) && (t.nmatches + 1) < t.nmismatches) {
    iout[1] = 0; l.n = 0;
  } else {
    /* List_reverse, then strip leading INDEL_COMP pairs: in push order these are the earliest pushes */
    for (k = 0; k < l.n && l.v[k].comp == INDEL_COMP; k++) ;
    if (k > 0) { for (i = k; i < l.n; i++) l.v[i-k] = l.v[i]; l.n -= k; }
  }
  bump(&iout[0]);
  /* end5 returns List_reverse(pairs) of the push-ordered list = head is last pushed; end3 returns push order */
  n = dump(&l,end5p ? 1 : 0,pairs,maxpairs);
  pl_free(&l);
  return n;
#undef END_NULL
}

/* ------------------------------------------------------------------------------------------------
 * Dynprog_genome_gap, dynprog_genome.c:3287
 * ---------------------------------------------------------------------------------------------- */
static int left_dinucl (char a, char aa, char b, char ba) {
  if ((a == 'G' || aa == 'G') && (b == 'T' || ba == 'T')) return LEFT_GT;
  if ((a == 'G' || aa == 'G') && (b == 'C' || ba == 'C')) return LEFT_GC;
  if ((a == 'A' || aa == 'A') && (b == 'T' || ba == 'T')) return LEFT_AT;
  if ((a == 'C' || aa == 'C') && (b == 'T' || ba == 'T')) return LEFT_CT;
  return 0;
}
static int right_dinucl (char r2, char r2a, char r1, char r1a) {
  if ((r2 == 'A' || r2a == 'A') && (r1 == 'G' || r1a == 'G')) return RIGHT_AG;
  if ((r2 == 'A' || r2a == 'A') && (r1 == 'C' || r1a == 'C')) return RIGHT_AC;
  if ((r2 == 'G' || r2a == 'G') && (r1 == 'C' || r1a == 'C')) return RIGHT_GC;
  if ((r2 == 'A' || r2a == 'A') && (r1 == 'T' || r1a == 'T')) return RIGHT_AT;
  return 0;
}

static int best_pairscore (int mt, int q, int g, int ga) {
  int a = pd[mt][q & 127][g & 127], b = pd[mt][q & 127][ga & 127];
  return a > b ? a : b;
}

/* Pair_maxnegscore, pair.c:8528.  `seq' is the list head first. */
static int maxnegscore (const orc_pair *seq, int n) {
  int maxneg = 0, prevhigh = 0, score = 0, i = 0;
  while (i < n) {
    if (seq[i].gapp) i++;
    else if (seq[i].comp == MISMATCH_COMP) { score += T_MISMATCH; if (score - prevhigh < maxneg) maxneg = score - prevhigh; i++; }
    else if (seq[i].comp == INDEL_COMP) {
      score += T_OPEN + T_INDEL; i++;
      while (i < n && seq[i].comp == INDEL_COMP) { score += T_INDEL; i++; }
      if (score - prevhigh < maxneg) maxneg = score - prevhigh;
    } else { score += T_MATCH; if (score > prevhigh) prevhigh = score; i++; }
  }
  return maxneg;
}

int orc_genome_gap (int *iout, double *dout, const char *queryseq, const char *queryuc,
		    int rlength, int glengthL, int glengthR, int roffset, int goffsetL, int rev_goffsetR,
		    const char *gsegL, const char *gsegL_alt, const char *gsegR, const char *gsegR_alt,
		    const double *left_probs, const double *right_probs,
		    int cdna_direction, int jump_late_p, int extraband_paired, double defect_rate,
		    int maxpeelback, int halfp, int finalp,
		    int max_rlength, int max_glength, orc_pair *pairs, int maxpairs) {
  static const int opens[3] = {-8,-7,-6}, extends[3] = {-3,-2,-1};	/* SINGLE_* == PAIRED_*, dynprog.h:59-75 */
  const char *rsequence = queryseq + roffset, *rsequenceuc = queryuc + roffset;
  const char *rev_rsequence = rsequence + rlength - 1, *rev_rsequenceuc = rsequenceuc + rlength - 1;
  const char *rev_gR = gsegR + glengthR - 1, *rev_gR_alt = gsegR_alt + glengthR - 1;
  int rev_roffset = roffset + rlength - 1;
  int mt, open, extend, bits, NEG;
  int *isc;
  int *dynprogindex = &iout[0], *new_left = &iout[1], *new_right = &iout[2], *traceback_score = &iout[3];
  int *nmatches = &iout[4], *nmismatches = &iout[5], *nopens = &iout[6], *nindels = &iout[7], *exonhead = &iout[8], *introntype = &iout[9];
  plist l;
  int n;

  orc_init();
  *nmatches = *nmismatches = *nopens = *nindels = 0;
  dout[0] = dout[1] = 0.0;
  *introntype = NONINTRON;
  if (rlength <= 1) { *traceback_score = NEG_INFINITY_32; return -1; }
  mt = mismatchtype_of(defect_rate);
  penalties(defect_rate,opens,extends,&open,&extend);
  (void) maxpeelback;

  if (rlength > max_rlength || glengthL > max_glength || glengthR > max_glength) {
    *new_left = goffsetL - 1; *new_right = rev_goffsetR + 1; *exonhead = roffset + rlength - 1;
    bump(dynprogindex); *traceback_score = NEG_INFINITY_32; *introntype = NONINTRON; return -1;
  }
  if (gsegL[0] == '\0' || gsegR[0] == '\0') { *traceback_score = NEG_INFINITY_32; return -1; }

  pl_init(&l);
  if (!finalp && defect_rate < 0.014) {
    /* genome_gap_simple, dynprog_genome.c:3005 (prelim tables; no known splice sites) */
    int *iscp = intron_score[cdna_direction > 0 ? 0 : (cdna_direction < 0 ? 1 : 2)][0];
    int scoreR = 0, scoreL = 0, rL, rR, bestscore = 0, bestscoreI = 0, bestrL = 0, bestrR = 0, itype = NONINTRON, scoreI, score, ldi, rdi, r, result, finalscore;
    tb_t t;
    for (rR = 1; rR < rlength; rR++) scoreR += best_pairscore(mt,rev_rsequenceuc[1-rR],rev_gR[1-rR],rev_gR_alt[1-rR]);
    for (rL = 1, rR = rlength-1; rL < rlength; rL++, rR--) {
      scoreL += best_pairscore(mt,rsequenceuc[rL-1],gsegL[rL-1],gsegL_alt[rL-1]);
      ldi = left_dinucl(gsegL[rL],gsegL_alt[rL],gsegL[rL+1],gsegL_alt[rL+1]);
      rdi = right_dinucl(rev_gR[-rR-1],rev_gR_alt[-rR-1],rev_gR[-rR],rev_gR_alt[-rR]);
      itype = ldi & rdi;
      scoreI = iscp[itype];
      if (itype != NONINTRON && (score = scoreL + scoreI + scoreR) >= bestscore) {
	bestscore = score; bestscoreI = scoreI; bestrL = rL; bestrR = rR; *introntype = itype;
      }
      scoreR -= best_pairscore(mt,rev_rsequenceuc[1-rR],rev_gR[1-rR],rev_gR_alt[1-rR]);
    }
    finalscore = halfp ? bestscore - bestscoreI/2 : bestscore;
    result = finalscore > 0;
    if (result) {
      dout[0] = left_probs[bestrL]; dout[1] = right_probs[bestrR];
      if (dout[0] < 0.90 || dout[1] < 0.90) result = 0;
    }
    *traceback_score = *nmatches = *nmismatches = 0;
    if (result) {
      orc_pair *gp;
      memset(&t,0,sizeof(t));
      t.rseq = rsequence; t.rsequc = rsequenceuc; t.gseq = gsegL; t.galt = gsegL_alt;
      t.queryoffset = roffset; t.genomeoffset = goffsetL; t.revp = 0; t.dynprogindex = *dynprogindex;
      for (r = 1; r <= bestrL; r++) diag_step(&l,&t,r,r);
      *new_left = goffsetL + (bestrL - 1);
      *new_right = *exonhead = rev_goffsetR - (bestrR - 1);
      gp = push_gapholder(&l,0,(*new_right) - (*new_left) - 1);
      gp->introntype = itype;		/* the LAST iteration's local introntype, dynprog_genome.c:3230 */
      gp->donor_prob = dout[0]; gp->acceptor_prob = dout[1];
      t.rseq = rev_rsequence; t.rsequc = rev_rsequenceuc; t.gseq = rev_gR; t.galt = rev_gR_alt;
      t.queryoffset = rev_roffset; t.genomeoffset = rev_goffsetR; t.revp = 1;
      for (r = bestrR; r > 0; r--) diag_step(&l,&t,r,r);
      *traceback_score = t.score; *nmatches = t.nmatches; *nmismatches = t.nmismatches;
      if (l.n > 0) {
	bump(dynprogindex);
	n = dump(&l,1,pairs,maxpairs);	/* pushes cons onto the head: head = last pushed */
	pl_free(&l); return n;
      }
    }
    l.n = 0;
  }

  bits = (rlength < use8p[mt] || (glengthL < use8p[mt] && glengthR < use8p[mt])) ? 8 : 16;
  NEG = bits == 8 ? -128 : -32768;
  isc = intron_score[cdna_direction > 0 ? 0 : (cdna_direction < 0 ? 1 : 2)][finalp ? 1 : 0];
  {
    ul_t L, R;
    int lbandL, ubandL, lbandR, ubandR;
    int *leftdi = (int *) calloc(glengthL + 1,sizeof(int)), *rightdi = (int *) calloc(glengthR + 1,sizeof(int));
    int rL, rR, cL, cR, cloL, chighL, cloR, chighR, G1;
    int bestrL = -1, bestrR = 0, bestcL = 0, bestcR = 0;
    int b2rL = 0, b2rR = 0, b2cL = 0, b2cR = 0;
    int bestscore = NEG, bestscore_with_dinucl = NEG, score, scoreL, scoreR, scoreI, use_dinucl, finalscore;
    double probL, probR, bestprob_with_score = 0.0, bestprob_with_dinucl = 0.0;
    int leftoffset = goffsetL, rightoffset = rev_goffsetR;
    tb_t t;

    orc_compute_bands(&lbandL,&ubandL,rlength,glengthL,extraband_paired,1);
    ul_alloc(&L,rlength,glengthL);
    orc_fill(1,bits,rsequence,gsegL,gsegL_alt,rlength,glengthL,mt,open,extend,lbandL,ubandL,jump_late_p,0,L.Hu,L.dNu,L.dEu,NULL);
    orc_fill(2,bits,rsequence,gsegL,gsegL_alt,rlength,glengthL,mt,open,extend,lbandL,ubandL,jump_late_p,0,L.Hl,L.dNl,L.dEl,NULL);
    orc_compute_bands(&lbandR,&ubandR,rlength,glengthR,extraband_paired,1);
    ul_alloc(&R,rlength,glengthR);
    orc_fill(1,bits,rsequence,gsegR,gsegR_alt,rlength,glengthR,mt,open,extend,lbandR,ubandR,!jump_late_p,1,R.Hu,R.dNu,R.dEu,NULL);
    orc_fill(2,bits,rsequence,gsegR,gsegR_alt,rlength,glengthR,mt,open,extend,lbandR,ubandR,!jump_late_p,1,R.Hl,R.dNl,R.dEl,NULL);
    if (count_only) { ul_free(&L); ul_free(&R); free(leftdi); free(rightdi); return -1; }

    /* bridge_intron_gap_{8,16}_site_level, dynprog_genome.c:866/1742 */
    for (cL = 0; cL < glengthL - 1; cL++) leftdi[cL] = left_dinucl(gsegL[cL],gsegL_alt[cL],gsegL[cL+1],gsegL_alt[cL+1]);
    for (cR = 0; cR < glengthR - 1; cR++) rightdi[cR] = right_dinucl(rev_gR[-cR-1],rev_gR_alt[-cR-1],rev_gR[-cR],rev_gR_alt[-cR]);

#define CONSIDER() do { \
      if ((score = scoreL + scoreI + scoreR) > bestscore) { \
	bestscore = score; bestrL = rL; bestrR = rR; bestcL = cL; bestcR = cR; bestprob_with_score = probL + probR; \
      } else if (score == bestscore && probL + probR > bestprob_with_score) { \
	bestrL = rL; bestrR = rR; bestcL = cL; bestcR = cR; bestprob_with_score = probL + probR; \
      } } while (0)

    for (rL = 1, rR = rlength-1; rL < rlength; rL++, rR--) {
      cloL = rL - lbandL; if (cloL < 1) cloL = 1;
      chighL = rL + ubandL; if (chighL > glengthL-1) chighL = glengthL-1;
      cloR = rR - lbandR; if (cloR < 1) cloR = 1;
      chighR = rR + ubandR; if (chighR > glengthR-1) chighR = glengthR-1;

      cL = rL; probL = left_probs[cL]; G1 = L.G1; scoreL = L.Hu[IDX(rL,cL)];
      cR = rR; probR = right_probs[cR]; G1 = R.G1; scoreR = R.Hu[IDX(rR,cR)];
      scoreI = isc[leftdi[cL] & rightdi[cR]];
      CONSIDER();
      if (scoreI > 0 && probL + probR > bestprob_with_dinucl) {
	bestscore_with_dinucl = scoreL + scoreI + scoreR;
	b2cL = cL; b2cR = cR; b2rL = rL; b2rR = rR; bestprob_with_dinucl = probL + probR;
      }

      /* indel on right */
      cL = rL; probL = left_probs[cL]; G1 = L.G1; scoreL = L.Hu[IDX(rL,cL)];
      G1 = R.G1;
      for (cR = cloR; cR < rR && cR < rightoffset-leftoffset-cL; cR++) {
	probR = right_probs[cR]; scoreR = R.Hl[IDX(rR,cR)]; scoreI = isc[leftdi[cL] & rightdi[cR]]; CONSIDER();
      }
      for (cR++; cR < chighR && cR < rightoffset-leftoffset-cL; cR++) {
	probR = right_probs[cR]; scoreR = R.Hu[IDX(rR,cR)]; scoreI = isc[leftdi[cL] & rightdi[cR]]; CONSIDER();
      }

      /* indel on left */
      cR = rR; probR = right_probs[cR]; G1 = R.G1; scoreR = R.Hu[IDX(rR,cR)];
      G1 = L.G1;
      for (cL = cloL; cL < rL && cL < rightoffset-leftoffset-cR; cL++) {
	probL = left_probs[cL]; scoreL = L.Hl[IDX(rL,cL)]; scoreI = isc[leftdi[cL] & rightdi[cR]]; CONSIDER();
      }
      for (cL++; cL < chighL && cL < rightoffset-leftoffset-cR; cL++) {
	probL = left_probs[cL]; scoreL = L.Hu[IDX(rL,cL)]; scoreI = isc[leftdi[cL] & rightdi[cR]]; CONSIDER();
      }
    }
#undef CONSIDER

    if (bestprob_with_score > 2*0.85) use_dinucl = 0;
    else if (bestprob_with_dinucl == 0.0) use_dinucl = 0;
    else if (bestscore_with_dinucl < 0 || bestscore_with_dinucl < bestscore - 9) use_dinucl = 0;
    else use_dinucl = 1;
    if (use_dinucl) { bestcL = b2cL; bestcR = b2cR; bestrL = b2rL; bestrR = b2rR; bestscore = bestscore_with_dinucl; }
    if (bestscore < 0) finalscore = bestscore;
    else if (halfp) finalscore = bestscore - isc[leftdi[bestcL] & rightdi[bestcR]]/2;
    else finalscore = bestscore;
    free(leftdi); free(rightdi);

    if (finalscore < 0) {
      ul_free(&L); ul_free(&R); pl_free(&l);
      *traceback_score = -100; return -1;
    }
    dout[0] = left_probs[bestcL]; dout[1] = right_probs[bestcR];	/* get_splicesite_probs :332 */

    *new_left = goffsetL + (bestcL - 1);
    *new_right = rev_goffsetR - (bestcR - 1);
    *exonhead = rev_roffset - (bestrR - 1);

    memset(&t,0,sizeof(t));
    t.dynprogindex = *dynprogindex;
    t.rseq = rev_rsequence; t.rsequc = rev_rsequenceuc; t.gseq = rev_gR; t.galt = rev_gR_alt;
    t.queryoffset = rev_roffset; t.genomeoffset = rev_goffsetR; t.revp = 1;
    if (bestcR >= bestrR) traceback_upper(&l,&t,R.dNu,R.dEu,R.G1,bestrR,bestcR);
    else traceback_lower(&l,&t,R.dNl,R.dEl,R.G1,bestrR,bestcR);
    pl_reverse(&l);		/* List_reverse: now as if pushed in the opposite order */
    {
      orc_pair *gp = push_gapholder(&l,(rev_roffset - bestrR) - (roffset + bestrL) + 1,(*new_right) - (*new_left) - 1);
      gp->introntype = *introntype; gp->donor_prob = dout[0]; gp->acceptor_prob = dout[1];
    }
    t.rseq = rsequence; t.rsequc = rsequenceuc; t.gseq = gsegL; t.galt = gsegL_alt;
    t.queryoffset = roffset; t.genomeoffset = goffsetL; t.revp = 0;
    if (bestcL >= bestrL) traceback_upper(&l,&t,L.dNu,L.dEu,L.G1,bestrL,bestcL);
    else traceback_lower(&l,&t,L.dNl,L.dEl,L.G1,bestrL,bestcL);
    ul_free(&L); ul_free(&R);

    *traceback_score = t.score; *nmatches = t.nmatches; *nmismatches = t.nmismatches; *nopens = t.nopens; *nindels = t.nindels;
    if (l.n == 1) l.n = 0;
    bump(dynprogindex);
    {
      /* head-first view for Pair_maxnegscore */
      orc_pair *hf = (orc_pair *) malloc((l.n + 1) * sizeof(orc_pair));
      int i, mn;
      for (i = 0; i < l.n; i++) hf[i] = l.v[l.n-1-i];
      mn = maxnegscore(hf,l.n);
      free(hf);
      if (mn < -10) { *traceback_score = -100; pl_free(&l); return -1; }
    }
    n = dump(&l,0,pairs,maxpairs);	/* List_reverse(pairs) = push order */
    pl_free(&l);
    return n;
  }
}

/* ------------------------------------------------------------------------------------------------
 * Dynprog_cdna_gap, dynprog_cdna.c:786
 * ---------------------------------------------------------------------------------------------- */
int orc_cdna_gap (int *iout, const char *queryseq, const char *queryuc,
		  int rlengthL, int rlengthR, int glength, int roffsetL, int rev_roffsetR, int goffset,
		  const char *gseg, const char *gseg_alt, const char *rev_gseg, const char *rev_gseg_alt,
		  int jump_late_p, int extraband_paired, double defect_rate,
		  int max_rlength, int max_glength, orc_pair *pairs, int maxpairs) {
  const int open = -10, extend = -7;			/* dynprog_cdna.c:32-38 */
  const char *rsequenceL = queryseq + roffsetL, *rsequence_ucL = queryuc + roffsetL;
  const char *rev_rsequenceR = queryseq + rev_roffsetR, *rev_rsequence_ucR = queryuc + rev_roffsetR;
  int mt, bits, NEG, rev_goffset, n;
  int lbandL, ubandL, lbandR, ubandR;
  ul_t L, R;
  plist l; tb_t t;
  int bestscore, score, scoreL, scoreR, pen, rL, rR, cL, cR, rloL, rhighL, rloR, rhighR, G1;
  int bestcL = 0, bestcR = 0, bestrL = 0, bestrR = 0, late = jump_late_p;
  int leftoffset = roffsetL, rightoffset = rev_roffsetR;
  int queryjump, genomejump, k;

  orc_init();
  if (glength <= 1) return -1;
  mt = mismatchtype_of(defect_rate);
  if (glength > max_glength || rlengthR > max_rlength || rlengthL > max_rlength) { bump(&iout[0]); return -1; }
  rev_goffset = goffset + glength - 1;
  bits = (glength < use8p[mt] || (rlengthL < use8p[mt] && rlengthR <= use8p[mt])) ? 8 : 16;
  NEG = bits == 8 ? -128 : -32768;
  if (gseg[0] == '\0' || rev_gseg[0] == '\0') return -1;

  orc_compute_bands(&lbandL,&ubandL,rlengthL,glength,extraband_paired,1);
  ul_alloc(&L,rlengthL,glength);
  orc_fill(1,bits,rsequenceL,gseg,gseg_alt,rlengthL,glength,mt,open,extend,lbandL,ubandL,jump_late_p,0,L.Hu,L.dNu,L.dEu,NULL);
  orc_fill(2,bits,rsequenceL,gseg,gseg_alt,rlengthL,glength,mt,open,extend,lbandL,ubandL,jump_late_p,0,L.Hl,L.dNl,L.dEl,NULL);
  orc_compute_bands(&lbandR,&ubandR,rlengthR,glength,extraband_paired,1);
  ul_alloc(&R,rlengthR,glength);
  {
    const char *rR_fwd = rev_rsequenceR - (rlengthR - 1);
    orc_fill(1,bits,rR_fwd,rev_gseg,rev_gseg_alt,rlengthR,glength,mt,open,extend,lbandR,ubandR,!jump_late_p,1,R.Hu,R.dNu,R.dEu,NULL);
    orc_fill(2,bits,rR_fwd,rev_gseg,rev_gseg_alt,rlengthR,glength,mt,open,extend,lbandR,ubandR,!jump_late_p,1,R.Hl,R.dNl,R.dEl,NULL);
  }
  if (count_only) { ul_free(&L); ul_free(&R); return -1; }

  /* bridge_cdna_gap_{8,16}_ud, dynprog_cdna.c:123/387 */
  bestscore = NEG;
#define CONSIDER() do { score = scoreL + scoreR + pen; \
    if (late ? score >= bestscore : score > bestscore) { bestscore = score; bestcL = cL; bestcR = cR; bestrL = rL; bestrR = rR; } } while (0)
  for (cL = 1; cL < glength; cL++) {
    for (cR = glength - cL, pen = 0; cR >= 0; cR--, pen += extend) {
      rloL = cL - ubandL; if (rloL < 1) rloL = 1;
      rhighL = cL + lbandL; if (rhighL > rlengthL-1) rhighL = rlengthL-1;
      rloR = cR - ubandR; if (rloR < 1) rloR = 1;
      rhighR = cR + lbandR; if (rhighR > rlengthR-1) rhighR = rlengthR-1;
      for (rL = rloL; rL <= rhighL; rL++) {
	G1 = L.G1;
	scoreL = (rL < cL) ? L.Hu[IDX(rL,cL)] : L.Hl[IDX(rL,cL)];
	G1 = R.G1;
	for (rR = rloR; rR < cR && rR < rightoffset-leftoffset-rL; rR++) { scoreR = R.Hu[IDX(rR,cR)]; CONSIDER(); }
	for ( ; rR <= rhighR && rR < rightoffset-leftoffset-rL; rR++) { scoreR = R.Hl[IDX(rR,cR)]; CONSIDER(); }
      }
      pen = open - extend;
    }
  }
#undef CONSIDER

  pl_init(&l);
  memset(&t,0,sizeof(t));
  t.dynprogindex = iout[0];
  t.rseq = rev_rsequenceR; t.rsequc = rev_rsequence_ucR; t.gseq = rev_gseg + glength - 1; t.galt = rev_gseg_alt + glength - 1;
  t.queryoffset = rev_roffsetR; t.genomeoffset = rev_goffset; t.revp = 1;
  if (bestcR >= bestrR) traceback_upper(&l,&t,R.dNu,R.dEu,R.G1,bestrR,bestcR);
  else traceback_lower(&l,&t,R.dNl,R.dEl,R.G1,bestrR,bestcR);
  pl_reverse(&l);

  queryjump = (rev_roffsetR - bestrR) - (roffsetL + bestrL) + 1;
  genomejump = (rev_goffset - bestcR) - (goffset + bestcL) + 1;
  if (queryjump == INSERT_PAIRS && genomejump == INSERT_PAIRS) {
    for (k = rev_roffsetR - bestrR; k >= roffsetL + bestrL; k--)
      push(&l,k,rev_goffset - bestcR + 1,rsequenceL[k - roffsetL],SHORTGAP_COMP,' ',' ',iout[0]);
    for (k = rev_goffset - bestcR; k >= goffset + bestcL; k--)
      push(&l,roffsetL + bestrL,k,' ',SHORTGAP_COMP,gseg[k - goffset],gseg_alt[k - goffset],iout[0]);
  } else {
    push_gapholder(&l,queryjump,genomejump);
    iout[2] = 1;
  }

  t.rseq = rsequenceL; t.rsequc = rsequence_ucL; t.gseq = gseg; t.galt = gseg_alt;
  t.queryoffset = roffsetL; t.genomeoffset = goffset; t.revp = 0;
  if (bestcL >= bestrL) traceback_upper(&l,&t,L.dNu,L.dEu,L.G1,bestrL,bestcL);
  else traceback_lower(&l,&t,L.dNl,L.dEl,L.G1,bestrL,bestcL);
  ul_free(&L); ul_free(&R);

  iout[1] = t.score;
  if (l.n == 1) l.n = 0;
  bump(&iout[0]);
  n = dump(&l,0,pairs,maxpairs);
  pl_free(&l);
  return n;
}
