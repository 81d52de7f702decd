/* Force-included (-include) when compiling the reference's HybridSearch.c for the oracle.
 * TEST INFRASTRUCTURE.  Two observation hooks, no change to any arithmetic:
 *   dwalltime()  -> oracle_dwalltime(__LINE__): lets the shim (a) make the fake FPGA look
 *                   infinitely slow (HybridSearch.c:218,224 / SSE :1494,1500) so the whole
 *                   database goes to the host SIMD team, (b) log the real CPU timestamps.
 *   sort_scores  -> oracle_sort_scores: dumps the raw int32 score row of each query
 *                   (HybridSearch.c:1217) before handing it to the reference's own sort. */
#ifndef OSWALD_ORACLE_HOOKS_H
#define OSWALD_ORACLE_HOOKS_H
#include <time.h>
#include "utils.h"          /* declare the real functions before the macros exist */
double oracle_dwalltime(int line);
void oracle_sort_scores(int *scores, char **titles, unsigned long int size, int threads);
#define dwalltime() oracle_dwalltime(__LINE__)
#define sort_scores oracle_sort_scores
#endif
