/* Score tables of the DP engine, built on the host and uploaded once per context.
 *
 * They are the slices of the reference's pairdistance_array / consistent_array
 * (/root/reference/src/dynprog.c:1007-1173, STANDARD mode) that the kernels index:
 *   U [mt][q][k]  = pairdistance[mt][q][ "ACGTN"[k] ]   raw query char  x genome class  (upper and full fills)
 *   Lw[mt][g][k]  = pairdistance[mt]["ACGTN"[k]][g]     query class     x raw genome char (lower fills)
 *   cons[a]       = 128-bit mask of the chars consistent with a (Dynprog_consistent_p, dynprog.c:895)
 *   isc[dir][finalp][bit] = intron_score_array_* of dynprog_genome.c:135-187 indexed by the single set
 *                   bit of (leftdi & rightdi): GTAG_FWD 0x20 -> 5 ... ATAC_REV 0x01 -> 0 (intron.h:27-33)
 */
#ifndef GMAPDP_TABLES_H
#define GMAPDP_TABLES_H

#include <stdint.h>
#include <string.h>
#include <ctype.h>

struct GdpTables {
  int8_t U[4][128][8];
  int8_t Lw[4][128][8];
  uint32_t cons[128][4];
  int32_t isc[3][2][8];
};

struct GdpHostTables {
  int16_t pd[4][128][128];
  uint8_t cons[128][128];
  int use8p[4];

  void pairing (int A, int B, int score) {
    int xs[2] = {A, tolower(A)}, ys[2] = {B, tolower(B)};
    for (int x = 0; x < 2; x++) for (int y = 0; y < 2; y++) {
      cons[xs[x]][ys[y]] = cons[ys[y]][xs[x]] = 1;
      for (int m = 0; m < 4; m++) pd[m][xs[x]][ys[y]] = pd[m][ys[y]][xs[x]] = (int16_t) score;
    }
  }

  GdpHostTables () {
    const int mismatch[4] = {-3, -2, -1, -5};	/* HIGHQ MEDQ LOWQ ENDQ */
    memset(pd,0,sizeof(pd)); memset(cons,0,sizeof(cons));
    for (int m = 0; m < 4; m++) {
      use8p[m] = -128 / mismatch[m] - 1;
      for (int a = 'A'; a <= 'z'; a++) for (int b = 'A'; b < 'z'; b++) pd[m][a][b] = (int16_t) mismatch[m];
    }
    for (int a = 'A'; a < 'Z'; a++) pairing(a,a,3);
    pairing('U','T',3);
    const char *two[] = {"RAG","YTC","WAT","SGC","MAC","KGT"};
    const char *many[] = {"HATC","BGCT","VGAC","DGAT","NTCAG","XTCAG"};
    for (auto s : two) for (int j = 1; s[j]; j++) pairing(s[0],s[j],1);
    for (auto s : many) for (int j = 1; s[j]; j++) pairing(s[0],s[j],3);
    pairing('N','N',3); pairing('X','X',3);
  }

  void device_tables (GdpTables *t) const {
    static const char L[5] = {'A','C','G','T','N'};
    memset(t,0,sizeof(*t));
    for (int m = 0; m < 4; m++) for (int x = 0; x < 128; x++) for (int k = 0; k < 5; k++) {
      t->U[m][x][k] = (int8_t) pd[m][x][(int) L[k]];
      t->Lw[m][x][k] = (int8_t) pd[m][(int) L[k]][x];
    }
    for (int a = 0; a < 128; a++) for (int b = 0; b < 128; b++) if (cons[a][b]) t->cons[a][b >> 5] |= 1u << (b & 31);
    /* bit index: 5 GTAG_FWD, 4 GCAG_FWD, 3 ATAC_FWD, 2 GTAG_REV, 1 GCAG_REV, 0 ATAC_REV */
    const int sense_final[6] = {0,0,0,8,10,16}, sense_prelim[6] = {0,0,0,4,8,14};
    const int anti_final[6] = {8,10,16,0,0,0}, anti_prelim[6] = {4,8,14,0,0,0};
    const int either_final[6] = {8,10,14,8,10,16}, either_prelim[6] = {4,8,14,4,8,16};	/* asymmetric: dynprog_genome.c:172-184 */
    for (int k = 0; k < 6; k++) {
      t->isc[0][1][k] = sense_final[k]; t->isc[0][0][k] = sense_prelim[k];
      t->isc[1][1][k] = anti_final[k]; t->isc[1][0][k] = anti_prelim[k];
      t->isc[2][1][k] = either_final[k]; t->isc[2][0][k] = either_prelim[k];
    }
  }
};

#endif
