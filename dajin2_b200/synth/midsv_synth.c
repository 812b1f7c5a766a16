/*
 * Synthetic nanopore-like MIDSV generator (test + bench infrastructure, not part of the product path).
 *
 * Produces the wire format of the hot path (the "MIDSV" string written by the
 * reference's generate_midsv, src/DAJIN2/core/preprocess/alignment/midsv_caller.py:266-272): one row per
 * read, exactly L comma-separated tokens, token i <-> reference position i.  Rows are laid out back to
 * back in one byte buffer, separated by a single '\n', described by row_off[r] / row_len[r].
 *
 * Every random draw comes from a counter-based generator keyed by (seed, read index), so the output for a
 * read does not depend on the number of threads or on which shard generates it: rank r of G can generate
 * reads r, r+G, ... and obtain exactly the rows a single process would.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int32_t kind; /* 1 substitution, 2 deletion, 3 insertion (before pos), 4 inversion (lowercase block) */
    int32_t pos;
    int32_t len;
} DjSynthFeature;

typedef struct {
    double sub_rate;       /* per-token substitution error rate            */
    double del_rate;       /* per-token deletion error rate                */
    double ins_rate;       /* per-token insertion error rate               */
    double ins_geom_p;     /* insertion length ~ Geometric(p), support >=1 */
    double homopolymer_x;  /* error-rate multiplier inside homopolymers>=4 */
    double trunc_prob;     /* probability that a read carries an =N tail   */
    double trunc_min;      /* tail length, fraction of L                   */
    double trunc_max;
    double plus_strand;    /* probability of '+' strand                    */
} DjSynthParams;

static inline uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static inline uint64_t read_stream(uint64_t seed, int64_t read_index) {
    uint64_t s = seed ^ (0xD6E8FEB86659FD93ULL * (uint64_t)(read_index + 1));
    splitmix64(&s);
    return s;
}

static inline double u01(uint64_t *s) { return (double)(splitmix64(s) >> 11) * (1.0 / 9007199254740992.0); }

static const char BASES[4] = {'A', 'C', 'G', 'T'};

static inline int base_index(char c) {
    switch (c) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        default: return 0;
    }
}

/* Per-locus plan of one allele: what the error-free read looks like against this reference. */
typedef struct {
    uint8_t *op;      /* 0 match, 1 substitution, 2 deletion */
    uint8_t *alt;     /* substitution alt base                */
    uint8_t *lower;   /* 1 inside an inversion block          */
    int32_t *ins_len; /* designed insertion before this locus */
} AllelePlan;

static void build_plan(AllelePlan *pl, int L, const char *ref, const DjSynthFeature *f, int nf) {
    pl->op = (uint8_t *)calloc(L, 1);
    pl->alt = (uint8_t *)calloc(L, 1);
    pl->lower = (uint8_t *)calloc(L, 1);
    pl->ins_len = (int32_t *)calloc(L, sizeof(int32_t));
    for (int k = 0; k < nf; k++) {
        int p = f[k].pos, n = f[k].len;
        if (p < 0 || p >= L) continue;
        switch (f[k].kind) {
            case 1:
                for (int i = p; i < p + n && i < L; i++) {
                    pl->op[i] = 1;
                    pl->alt[i] = (uint8_t)BASES[(base_index(ref[i]) + 1) & 3];
                }
                break;
            case 2:
                for (int i = p; i < p + n && i < L; i++) pl->op[i] = 2;
                break;
            case 3:
                pl->ins_len[p] = n;
                break;
            case 4:
                for (int i = p; i < p + n && i < L; i++) pl->lower[i] = 1;
                break;
            default:
                break;
        }
    }
}

static void free_plan(AllelePlan *pl) {
    free(pl->op);
    free(pl->alt);
    free(pl->lower);
    free(pl->ins_len);
}

static inline char lc(char c, int lower) { return lower ? (char)(c | 0x20) : c; }

/* Emit one read.  When out == NULL only the length is computed.  Returns bytes (without separator). */
static int64_t emit_read(uint8_t *out, const DjSynthParams *P, uint64_t seed, int64_t read_index, int L,
                         const char *ref, const uint8_t *homo, const AllelePlan *plans, const double *cdf,
                         int n_alleles, uint8_t *strand_out, uint8_t *allele_out) {
    uint64_t s = read_stream(seed, read_index);
    /* fixed draw order: allele, strand, truncation (shared by every FASTA allele generated from one seed) */
    double ua = u01(&s);
    int a = 0;
    while (a < n_alleles - 1 && ua >= cdf[a]) a++;
    int plus = u01(&s) < P->plus_strand;
    double ut = u01(&s);
    double ufrac = u01(&s);
    double uside = u01(&s);
    int n_lo = 0, n_hi = L; /* tokens outside [n_lo, n_hi) are =N */
    if (ut < P->trunc_prob) {
        int tail = (int)((P->trunc_min + (P->trunc_max - P->trunc_min) * ufrac) * L);
        if (uside < 0.5) n_hi = L - tail; else n_lo = tail;
    }
    if (strand_out) *strand_out = plus ? '+' : '-';
    if (allele_out) *allele_out = (uint8_t)a;
    const AllelePlan *pl = &plans[a];
    /* designed insertions use a sequence that depends only on (seed, locus) so that all reads agree */
    int64_t n = 0;
/* the argument is always evaluated: some call sites draw from the RNG inside it */
#define PUT(c) do { uint8_t put_c_ = (uint8_t)(c); if (out) out[n] = put_c_; n++; } while (0)
    for (int i = 0; i < L; i++) {
        if (i) PUT(',');
        if (i < n_lo || i >= n_hi) { PUT('='); PUT('N'); continue; }
        int lower = pl->lower[i];
        char r = ref[i];
        if (pl->ins_len[i] > 0) {
            uint64_t si = seed ^ (0xA24BAED4963EE407ULL * (uint64_t)(i + 1));
            for (int k = 0; k < pl->ins_len[i]; k++) {
                PUT('+'); PUT(lc(BASES[splitmix64(&si) & 3], lower)); PUT('|');
            }
        }
        if (pl->op[i] == 1) { PUT('*'); PUT(lc(r, lower)); PUT(lc((char)pl->alt[i], lower)); continue; }
        if (pl->op[i] == 2) { PUT('-'); PUT(lc(r, lower)); continue; }
        if (pl->ins_len[i] > 0) { PUT('='); PUT(lc(r, lower)); continue; }
        /* plain match locus: apply sequencing errors */
        double mult = homo[i] ? P->homopolymer_x : 1.0;
        double u = u01(&s);
        double ps = P->sub_rate * mult, pd = P->del_rate * mult, pi = P->ins_rate * mult;
        if (u < ps) {
            int alt = (base_index(r) + 1 + (int)(splitmix64(&s) % 3)) & 3;
            PUT('*'); PUT(lc(r, lower)); PUT(lc(BASES[alt], lower));
        } else if (u < ps + pd) {
            PUT('-'); PUT(lc(r, lower));
        } else if (u < ps + pd + pi) {
            int k = 1;
            while (u01(&s) >= P->ins_geom_p && k < 64) k++;
            for (int j = 0; j < k; j++) { PUT('+'); PUT(lc(BASES[splitmix64(&s) & 3], lower)); PUT('|'); }
            PUT('='); PUT(lc(r, lower));
        } else {
            PUT('='); PUT(lc(r, lower));
        }
    }
#undef PUT
    return n;
}

static uint8_t *homopolymer_mask(const char *ref, int L) {
    uint8_t *m = (uint8_t *)calloc(L, 1);
    int i = 0;
    while (i < L) {
        int j = i;
        while (j < L && ref[j] == ref[i]) j++;
        if (j - i >= 4) memset(m + i, 1, j - i);
        i = j;
    }
    return m;
}

/*
 * Generate reads first_read, first_read+stride, ... (n_reads of them).
 * Pass text == NULL to obtain only the required byte count (rows + one '\n' each) in the return value and
 * row_len[].  With text != NULL, row_off/row_len/strand/allele are filled and rows are written.
 * Returns total bytes, or -1 when cap is too small.
 */
int64_t dj_synth_generate(const DjSynthParams *P, uint64_t seed, int64_t first_read, int64_t stride,
                          int64_t n_reads, int32_t L, const char *ref, const DjSynthFeature *features,
                          const int32_t *allele_first, const double *allele_cdf, int32_t n_alleles,
                          uint8_t *text, int64_t cap, int64_t *row_off, int32_t *row_len, uint8_t *strand,
                          uint8_t *allele) {
    AllelePlan *plans = (AllelePlan *)calloc(n_alleles, sizeof(AllelePlan));
    for (int a = 0; a < n_alleles; a++)
        build_plan(&plans[a], L, ref, features + allele_first[a], allele_first[a + 1] - allele_first[a]);
    uint8_t *homo = homopolymer_mask(ref, L);

#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < n_reads; t++) {
        int64_t ridx = first_read + t * stride;
        row_len[t] = (int32_t)emit_read(NULL, P, seed, ridx, L, ref, homo, plans, allele_cdf, n_alleles,
                                        strand ? &strand[t] : NULL, allele ? &allele[t] : NULL);
    }
    int64_t total = 0;
    for (int64_t t = 0; t < n_reads; t++) {
        if (row_off) row_off[t] = total;
        total += (int64_t)row_len[t] + 1;
    }
    if (text != NULL) {
        if (total > cap) { total = -1; goto done; }
#pragma omp parallel for schedule(static)
        for (int64_t t = 0; t < n_reads; t++) {
            int64_t ridx = first_read + t * stride;
            int64_t n = emit_read(text + row_off[t], P, seed, ridx, L, ref, homo, plans, allele_cdf, n_alleles,
                                  NULL, NULL);
            text[row_off[t] + n] = '\n';
        }
    }
done:
    for (int a = 0; a < n_alleles; a++) free_plan(&plans[a]);
    free(plans);
    free(homo);
    return total;
}

/* uuid4-looking QNAME derived from (seed, read index); 36 bytes, no terminator. */
void dj_synth_qname(uint64_t seed, int64_t read_index, char *out36) {
    static const char HEX[] = "0123456789abcdef";
    uint64_t s = read_stream(seed ^ 0x5851F42D4C957F2DULL, read_index);
    uint64_t a = splitmix64(&s), b = splitmix64(&s);
    int k = 0;
    for (int i = 0; i < 32; i++) {
        if (i == 8 || i == 12 || i == 16 || i == 20) out36[k++] = '-';
        uint64_t w = i < 16 ? a : b;
        out36[k++] = HEX[(w >> (4 * (i & 15))) & 15];
    }
}

void dj_synth_qnames(uint64_t seed, int64_t first_read, int64_t stride, int64_t n_reads, char *out) {
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < n_reads; t++) dj_synth_qname(seed, first_read + t * stride, out + 36 * t);
}
