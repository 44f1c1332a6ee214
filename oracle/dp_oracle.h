/* TEST INFRASTRUCTURE ONLY.
 *
 * dp_oracle: a scalar CPU restatement, in plain C, of GMAP 2024-02-22's alignment dynamic
 * programming (the SIMD variant that `gmap.avx2` runs).  It is the checker for the CUDA path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product library (gmap_2024_b200/csrc) never links or calls it.
 *
 * Parity status: PINNED.  Every function here is compared against the compiled, unmodified
 * reference (oracle/_ref, built by oracle/Makefile from /root/reference/src) by
 * tests/test_oracle_vs_ref.py -- cell for cell for the six fills, and out-parameter by
 * out-parameter plus the whole pair list for the five entry points -- and against the committed
 * golden vectors under tests/golden/ that were generated from that same compiled reference.
 *
 * File:line citations are into /root/reference/src.
 */
#ifndef DP_ORACLE_H
#define DP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* One alignment pair, the observable fields of struct Pair_T (pairdef.h:12-51). */
typedef struct {
  int querypos, genomepos, queryjump, genomejump, dynprogindex, introntype, gapp;
  char cdna, comp, genome, genomealt;
  double donor_prob, acceptor_prob;
} orc_pair;

/* Mismatchtype_T (dynprog.h:54) and Endalign_T (dynprog.h:27) */
enum { ORC_HIGHQ = 0, ORC_MEDQ = 1, ORC_LOWQ = 2, ORC_ENDQ = 3 };
enum { ORC_QUERYEND_GAP = 0, ORC_QUERYEND_INDELS = 1, ORC_QUERYEND_NOGAPS = 2, ORC_BEST_LOCAL = 3 };

void orc_init (void);						/* Dynprog_init dynprog.c:1007 (STANDARD mode) */
void orc_set_user_dynprog (int open, int extend, int enabled);	/* --indel-open / --indel-extend (the *_setup statics) */
int orc_pairdistance (int mismatchtype, int a, int b);
int orc_consistent (int a, int b);
int orc_use8p_size (int mismatchtype);
void orc_compute_bands (int *lband, int *uband, int rlength, int glength, int extraband, int widebandp); /* dynprog.c:1246 */

/* kind 0 full / 1 upper / 2 lower; bits 8 or 16.  Same calling convention as refdrv_fill
 * (oracle/ref_driver.c): forward arrays + revp, dense (rlength+1)x(glength+1) row-major outputs,
 * only in-band cells written.  Restates dynprog_simd.c:2987,4304,5340,6562,7714,8586. */
int orc_fill (int kind, int bits, const char *rseq, const char *gseq, const char *galt,
	      int rlength, int glength, int mismatchtype, int open, int extend,
	      int lband, int uband, int jump_late_p, int revp,
	      short *H, signed char *dN, signed char *dE, signed char *dF);

/* Algorithmic in-band cells of one fill. */
long orc_cells (int kind, int rlength, int glength, int lband, int uband);
/* cells of the fills run by the entry points since the last reset; count-only mode: the entry points count and return -1
   without filling (used by bench.py's CPU arms, which must not count their work with the product library) */
void orc_cells_reset (void);
long orc_cells_filled (void);
void orc_set_count_only (int on);

/* The five entry points.  Genomic segments are passed as the char arrays the reference fetches
 * with Genome_get_segment_right/left (always in ascending memory order; "rev" users read them from
 * the last char backwards).  Return value: number of pairs (list head first), or -1 for a NULL list.
 * iout layouts match oracle/ref_driver.c. */
int orc_single_gap (int *iout, const char *queryseq, const char *queryuc,
		    int rlength, int glength, int roffset, int goffset,
		    const char *gseg, const char *gseg_alt, int jump_late_p,
		    int extraband_single, int widebandp, double defect_rate,
		    int max_rlength, int max_glength, orc_pair *pairs, int maxpairs);

int orc_end_gap (int end5p, int *iout, const char *queryseq, const char *queryuc,
		 int rlength, int glength, int roffset, int goffset,
		 const char *gseg, const char *gseg_alt, int jump_late_p,
		 int extraband_end, double defect_rate, int endalign, int require_pos_score_p,
		 int max_rlength, int max_glength, orc_pair *pairs, int maxpairs);

/* left_probs[0..glengthL-2], right_probs[0..glengthR-2]: the MaxEnt splice-site probabilities the
 * reference computes at dynprog_genome.c:970-1061 (they are inputs here: maxent_hr.c is generated
 * table code that stays on the reference's side of the boundary). */
int orc_genome_gap (int *iout, double *dout, const char *queryseq, const char *queryuc,
		    int rlength, int glengthL, int glengthR, int roffset, int goffsetL, int rev_goffsetR,
		    const char *gsegL, const char *gsegL_alt, const char *gsegR, const char *gsegR_alt,
		    const double *left_probs, const double *right_probs,
		    int cdna_direction, int jump_late_p, int extraband_paired, double defect_rate,
		    int maxpeelback, int halfp, int finalp,
		    int max_rlength, int max_glength, orc_pair *pairs, int maxpairs);

int orc_cdna_gap (int *iout, const char *queryseq, const char *queryuc,
		  int rlengthL, int rlengthR, int glength, int roffsetL, int rev_roffsetR, int goffset,
		  const char *gseg, const char *gseg_alt, const char *rev_gseg, const char *rev_gseg_alt,
		  int jump_late_p, int extraband_paired, double defect_rate,
		  int max_rlength, int max_glength, orc_pair *pairs, int maxpairs);

#ifdef __cplusplus
}
#endif
#endif
