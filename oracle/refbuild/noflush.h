/* Build variant of the reference for the fairness figure of BASELINE.md section 3 / SURVEY.md F5: the stock
 * Dynprog_simd_* fills execute _mm_clflush(&H_nogap_r) once per column step (12 live sites in dynprog_simd.c), which
 * costs the CPU path 4-5x and changes no result.  This header is force-included (-include) in front of the UNMODIFIED
 * dynprog_simd.c for the *_noflush objects only: the intrinsic headers are read first, then the call is defined away. */
#include <emmintrin.h>
#include <immintrin.h>
#undef _mm_clflush
#define _mm_clflush(p) ((void) 0)
