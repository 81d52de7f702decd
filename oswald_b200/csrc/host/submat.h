/* submat.h - substitution matrices (reference host/src/submat.h, submat.c). */
#ifndef OSW_SUBMAT_H
#define OSW_SUBMAT_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
int osw_matrix_count(void);
const char *osw_matrix_name(int k);
/* Fills out[24*32] (reference layout, m[r*32+c]).  0 on success, -1 for an unknown name. */
int osw_matrix_by_name(const char *name, int8_t *out);
#ifdef __cplusplus
}
#endif
#endif
