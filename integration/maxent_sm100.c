/* maxent_sm100.c -- the reference's maxent_hr.c compiled IN PLACE (#include, nothing copied; this object replaces
 * maxent_hr.o in gmap.sm100's link, exactly as stage2_sm100.c replaces stage2.o), plus one accessor: the parameters of
 * the MaxEnt splice-site model are static arrays of that file, and the device evaluates the model itself
 * (csrc/gmapdp_genome.h), so the binding hands the table pointers to the runtime once (gmapdp_stream_genome). */
#include "maxent_hr.c"
#include "gmapdp_b200.h"

void sm100_maxent_tables (gmapdp_maxent_tables *t) {
  t->donor_plus = donor_score_plus; t->donor_di_plus = donor_discore_plus;
  t->acc1_plus = acc_score1_plus; t->acc2_plus = acc_score2_plus; t->acc3_plus = acc_score3_plus; t->acc_di_plus = acc_discore_plus;
  t->acc467_plus = acc_score467_plus; t->acc589_plus = acc_score589_plus;
  t->donor_minus = donor_score_minus; t->donor_di_minus = donor_discore_minus;
  t->acc1_minus = acc_score1_minus; t->acc2_minus = acc_score2_minus; t->acc3_minus = acc_score3_minus; t->acc_di_minus = acc_discore_minus;
  t->acc467_minus = acc_score467_minus; t->acc589_minus = acc_score589_minus;
}
