/* ref_maxent_tables.c -- test infrastructure.  The parameters of the reference's MaxEnt splice-site model are static
 * arrays of maxent_hr.c; this file compiles that source IN PLACE (#include, nothing copied) into its own shared
 * object, oracle/_ref/ref_maxent.so, and hands out the sixteen table pointers in the order of gmapdp_maxent_tables
 * (include/gmapdp_b200.h).  Tests use them to feed the product's on-device MaxEnt with the reference's own numbers. */
#include "maxent_hr.c"

void refme_tables (const double **p) {
  p[0] = donor_score_plus; p[1] = donor_discore_plus; p[2] = acc_score1_plus; p[3] = acc_score2_plus; p[4] = acc_score3_plus;
  p[5] = acc_discore_plus; p[6] = acc_score467_plus; p[7] = acc_score589_plus;
  p[8] = donor_score_minus; p[9] = donor_discore_minus; p[10] = acc_score1_minus; p[11] = acc_score2_minus; p[12] = acc_score3_minus;
  p[13] = acc_discore_minus; p[14] = acc_score467_minus; p[15] = acc_score589_minus;
}
