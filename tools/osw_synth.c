/* tools/osw_synth.c - deterministic synthetic protein data for tests and benches.
 *
 * Residues are i.i.d. uniform over the 20 standard amino acids; lengths are log-normal
 * (mu, sigma of ln L) clipped to [min, max]  (SURVEY.md section 8(d)).  Every sequence has
 * its own counter-based stream (seed, index), so output does not depend on thread count.
 *
 * As a program:
 *   osw_synth db -n N [-mu 5.6 -sigma 0.6 -min 10 -max 65535 -seed 1]
 *             [-plant q.fasta -plantmin 3000] [-long K -longmin 35000 -longmax 65535]
 *             [-tandem q.fasta -copies 6] -o db.fasta
 *   osw_synth queries -lengths 144,189,... [-seed 7] -o q.fasta
 * As a library (libosw_synth.so): osw_synth_lengths / osw_synth_codes fill arrays directly
 * (residue codes 0..22 in the reference's alphabet order), so a bench can build a 1.3 G
 * residue database without going through FASTA.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const char AA20[21] = "ACDEFGHIKLMNPQRSTVWY";
/* codes of AA20 in the alphabet ABCDEFGHIKLMNPQRSTVWXYZ */
static const uint8_t AA20_CODE[20] = {0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 21};

typedef struct { uint64_t s; } rng_t;
static uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static rng_t rng_for(uint64_t seed, uint64_t stream) { rng_t r; r.s = mix64(seed * 0x100000001B3ull ^ mix64(stream)); return r; }
static uint64_t rng_next(rng_t *r) { r->s += 0x9E3779B97F4A7C15ull; return mix64(r->s); }
static double rng_unit(rng_t *r) { return ((rng_next(r) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

/* ---- library -------------------------------------------------------------------------- */
void osw_synth_lengths(uint64_t n, double mu, double sigma, uint32_t lo, uint32_t hi,
                       uint64_t seed, uint16_t *out) {
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)n; ++i) {
        rng_t r = rng_for(seed ^ 0xA5A5A5A5ull, (uint64_t)i);
        double u1 = rng_unit(&r), u2 = rng_unit(&r);
        double g = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
        double L = floor(exp(mu + sigma * g) + 0.5);
        if (L < lo) L = lo;
        if (L > hi) L = hi;
        out[i] = (uint16_t)L;
    }
}
/* Fill sequence i (stream index `stream[i]`, or i when stream is NULL) at out+off[i]. */
void osw_synth_codes(uint64_t n, const uint16_t *lengths, const uint64_t *off,
                     const uint64_t *stream, uint64_t seed, int as_letters, uint8_t *out) {
#pragma omp parallel for schedule(dynamic, 256)
    for (long long i = 0; i < (long long)n; ++i) {
        rng_t r = rng_for(seed, stream ? stream[i] : (uint64_t)i);
        uint8_t *p = out + off[i];
        uint32_t L = lengths[i];
        uint32_t k = 0;
        while (k < L) {                       /* 12 residues per 64-bit draw (20^12 < 2^52) */
            uint64_t x = rng_next(&r) >> 12;
            for (int t = 0; t < 12 && k < L; ++t, ++k) {
                unsigned d = (unsigned)(x % 20u); x /= 20u;
                p[k] = as_letters ? (uint8_t)AA20[d] : AA20_CODE[d];
            }
        }
    }
}

/* ---- program -------------------------------------------------------------------------- */
#ifndef OSW_SYNTH_NO_MAIN
typedef struct { char **seq; uint32_t *len; int n; } qset_t;

static qset_t read_fasta(const char *path) {
    qset_t q = {0, 0, 0};
    FILE *f = fopen(path, "r");
    if (!f) { fprintf(stderr, "osw_synth: cannot open %s\n", path); exit(2); }
    char line[4096]; int cap = 0;
    while (fgets(line, sizeof line, f)) {
        size_t l = strlen(line);
        while (l && (line[l - 1] == '\n' || line[l - 1] == '\r')) line[--l] = 0;
        if (line[0] == '>') {
            if (q.n == cap) { cap = cap ? 2 * cap : 16; q.seq = realloc(q.seq, cap * sizeof *q.seq); q.len = realloc(q.len, cap * sizeof *q.len); }
            q.seq[q.n] = calloc(1, 1); q.len[q.n] = 0; q.n++;
        } else if (q.n) {
            q.seq[q.n - 1] = realloc(q.seq[q.n - 1], q.len[q.n - 1] + l + 1);
            memcpy(q.seq[q.n - 1] + q.len[q.n - 1], line, l + 1);
            q.len[q.n - 1] += (uint32_t)l;
        }
    }
    fclose(f);
    return q;
}
static void put_record(FILE *o, const char *title, const uint8_t *s, uint32_t L) {
    fprintf(o, ">%s len%u\n", title, L);
    for (uint32_t k = 0; k < L; k += 60) {
        uint32_t w = L - k < 60 ? L - k : 60;
        fwrite(s + k, 1, w, o); fputc('\n', o);
    }
}
static void mutate(uint8_t *s, uint32_t L, double rate, rng_t *r) {
    for (uint32_t k = 0; k < L; ++k)
        if (rng_unit(r) < rate) s[k] = (uint8_t)AA20[rng_next(r) % 20];
}
static const char *opt(int argc, char **argv, const char *name, const char *dflt) {
    for (int i = 2; i + 1 < argc; ++i) if (strcmp(argv[i], name) == 0) return argv[i + 1];
    return dflt;
}

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: osw_synth db|queries ... (see header comment)\n"); return 1; }
    const char *outp = opt(argc, argv, "-o", NULL);
    if (!outp) { fprintf(stderr, "osw_synth: -o required\n"); return 1; }
    uint64_t seed = strtoull(opt(argc, argv, "-seed", "1"), 0, 10);
    FILE *o = fopen(outp, "w");
    if (!o) { fprintf(stderr, "osw_synth: cannot write %s\n", outp); return 2; }
    char title[128];
    if (strcmp(argv[1], "queries") == 0) {
        char *spec = strdup(opt(argc, argv, "-lengths", "144"));
        int qi = 0;
        for (char *tok = strtok(spec, ","); tok; tok = strtok(NULL, ","), ++qi) {
            uint16_t L = (uint16_t)atoi(tok); uint64_t off = 0;
            uint8_t *s = malloc(L + 1);
            uint64_t st = (uint64_t)qi;
            osw_synth_codes(1, &L, &off, &st, seed ^ 0x51ull, 1, s);
            snprintf(title, sizeof title, "q%d", qi);
            put_record(o, title, s, L);
            free(s);
        }
    } else {
        uint64_t n = strtoull(opt(argc, argv, "-n", "1000"), 0, 10);
        double mu = atof(opt(argc, argv, "-mu", "5.6")), sigma = atof(opt(argc, argv, "-sigma", "0.6"));
        uint32_t lo = (uint32_t)atoi(opt(argc, argv, "-min", "10")), hi = (uint32_t)atoi(opt(argc, argv, "-max", "65535"));
        uint16_t *len = malloc((n ? n : 1) * sizeof *len);
        osw_synth_lengths(n, mu, sigma, lo, hi, seed, len);
        uint8_t *buf = malloc(65536);
        for (uint64_t i = 0; i < n; ++i) {
            uint64_t off = 0;
            osw_synth_codes(1, &len[i], &off, &i, seed, 1, buf);
            snprintf(title, sizeof title, "s%llu", (unsigned long long)i);
            put_record(o, title, buf, len[i]);
        }
        rng_t r = rng_for(seed, 0xFFFF0001ull);
        const char *plant = opt(argc, argv, "-plant", NULL);
        if (plant) {
            qset_t q = read_fasta(plant);
            uint32_t pmin = (uint32_t)atoi(opt(argc, argv, "-plantmin", "3000"));
            static const double rate[3] = {0.0, 0.10, 0.30};
            for (int k = 0; k < q.n; ++k) {
                if (q.len[k] < pmin) continue;
                for (int v = 0; v < 3; ++v) {
                    memcpy(buf, q.seq[k], q.len[k]);
                    mutate(buf, q.len[k], rate[v], &r);
                    snprintf(title, sizeof title, "plant%d_mut%d", k, (int)(rate[v] * 100));
                    put_record(o, title, buf, q.len[k]);
                }
            }
        }
        int nlong = atoi(opt(argc, argv, "-long", "0"));
        uint32_t lmin = (uint32_t)atoi(opt(argc, argv, "-longmin", "35000")), lmax = (uint32_t)atoi(opt(argc, argv, "-longmax", "65535"));
        for (int k = 0; k < nlong; ++k) {
            uint16_t L = (uint16_t)(lmin + rng_next(&r) % (lmax - lmin + 1)); uint64_t off = 0, st = 0xFFFF1000ull + k;
            osw_synth_codes(1, &L, &off, &st, seed, 1, buf);
            snprintf(title, sizeof title, "long%d", k);
            put_record(o, title, buf, L);
        }
        const char *tandem = opt(argc, argv, "-tandem", NULL);
        if (tandem) {
            qset_t q = read_fasta(tandem);
            int copies = atoi(opt(argc, argv, "-copies", "6"));
            int big = 0;
            for (int k = 1; k < q.n; ++k) if (q.len[k] > q.len[big]) big = k;
            uint32_t L = 0;
            for (int c = 0; c < copies && L + q.len[big] + 50 <= 65535; ++c) {
                memcpy(buf + L, q.seq[big], q.len[big]); L += q.len[big];
                for (int g = 0; g < 50; ++g) buf[L++] = (uint8_t)AA20[rng_next(&r) % 20];
            }
            snprintf(title, sizeof title, "tandem%d_x%d", big, copies);
            put_record(o, title, buf, L);
        }
    }
    fclose(o);
    return 0;
}
#endif
