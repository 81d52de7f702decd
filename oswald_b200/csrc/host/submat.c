/* submat.c - the eight substitution matrices of the reference (host/src/submat.c:4-227:
 * blosum45/50/62/80/90, pam30/70/250), reproduced value for value in the reference's
 * layout: 24 rows x 32 columns int8, rows/columns in alphabet order ABCDEFGHIKLMNPQRSTVWXYZ,
 * row 23 and columns 23..31 zero (the padding residue scores 0 against everything).
 * All eight are symmetric, so only the lower triangle is stored (tools/make_submat.py
 * generates submat_tri.inc from the reference's data and pins sha256 sums of the expanded
 * tables in tests/golden/submat.json). */
#include "submat.h"
#include <string.h>

static const struct { const char *name; const char *tri; } osw_tables[] = {
#include "submat_tri.inc"
};
#define OSW_N_TABLES (int)(sizeof osw_tables / sizeof osw_tables[0])

int osw_matrix_count(void) { return OSW_N_TABLES; }
const char *osw_matrix_name(int k) { return k >= 0 && k < OSW_N_TABLES ? osw_tables[k].name : 0; }

int osw_matrix_by_name(const char *name, int8_t *out) {
    if (!name || !out) return -1;
    for (int k = 0; k < OSW_N_TABLES; ++k) {
        if (strcmp(osw_tables[k].name, name) != 0) continue;
        memset(out, 0, 24 * 32);
        const char *p = osw_tables[k].tri;
        for (int r = 0; r < 23; ++r)
            for (int c = 0; c <= r; ++c) {
                int8_t v = (int8_t)(*p++ - 'A' - 17);
                out[r * 32 + c] = v;
                out[c * 32 + r] = v;
            }
        return 0;
    }
    return -1;
}
