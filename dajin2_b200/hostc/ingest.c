/* MIDSV JSONL tokenizer: locates, for every line of a JSONL buffer, the byte spans of the "MIDSV" and "QNAME" string
 * values and the "STRAND" character, without copying or re-encoding the text (the encoder kernel reads the MIDSV
 * strings where they lie).  Replaces json.loads + dict access per line of the reference's read_jsonl
 * (src/DAJIN2/utils/fileio.py:49-52) for the hot path's input files, written by generate_midsv
 * (src/DAJIN2/core/preprocess/alignment/midsv_caller.py:266-272): flat objects whose values are strings.
 * Lines that do not fit that shape (escapes inside the values, missing keys, nested values holding the keys) are
 * reported as irregular and the caller falls back to a real JSON parse.  Threads work on runs of whole lines. */
#define _GNU_SOURCE
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static int64_t count_newlines(const uint8_t *p, int64_t n) {
    int64_t c = 0;
    const uint8_t *end = p + n;
    while (p < end) {
        const uint8_t *q = memchr(p, '\n', (size_t)(end - p));
        if (!q) break;
        c++;
        p = q + 1;
    }
    return c;
}

static inline int is_blank_line(const uint8_t *b, const uint8_t *e) {
    for (; b < e; b++)
        if (*b != ' ' && *b != '\t' && *b != '\r') return 0;
    return 1;
}

/* number of non-blank lines (the last one need not end with a newline) */
int64_t dj_jsonl_count_lines(const uint8_t *buf, int64_t n, int n_threads) {
    if (n <= 0) return 0;
    if (n_threads < 1) n_threads = 1;
    int64_t total = 0;
    /* blank lines are rare: count newlines in parallel, then subtract blank lines in a cheap second look */
#pragma omp parallel for num_threads(n_threads) reduction(+ : total) schedule(static)
    for (int t = 0; t < n_threads; t++) {
        int64_t lo = n * t / n_threads, hi = n * (t + 1) / n_threads;
        /* lines that START in [lo, hi) */
        if (t > 0 && lo > 0) { /* lo == 0 when the file is shorter than n_threads bytes: the line starts at 0 */
            const uint8_t *q = memchr(buf + lo - 1, '\n', (size_t)(n - lo + 1));
            lo = q ? (q - buf) + 1 : n;
        }
        int64_t c = 0;
        const uint8_t *p = buf + lo;
        while (p < buf + n && (p - buf) < hi) {
            const uint8_t *q = memchr(p, '\n', (size_t)(buf + n - p));
            const uint8_t *e = q ? q : buf + n;
            if (!is_blank_line(p, e)) c++;
            p = e + 1;
        }
        total += c;
    }
    return total;
}

/* end of the JSON string that starts after the opening quote at s: returns pointer to the closing quote or NULL;
 * *escaped is set when the string holds a backslash */
static inline const uint8_t *string_end(const uint8_t *s, const uint8_t *line_end, int *escaped) {
    const uint8_t *p = s;
    for (;;) {
        const uint8_t *q = memchr(p, '"', (size_t)(line_end - p));
        if (!q) return 0;
        /* a quote preceded by an odd number of backslashes is escaped */
        const uint8_t *b = q;
        while (b > s && b[-1] == '\\') b--;
        if (((q - b) & 1) == 0) {
            if (memchr(s, '\\', (size_t)(q - s))) *escaped = 1;
            return q;
        }
        *escaped = 1;
        p = q + 1;
    }
}

static int index_line(const uint8_t *buf, const uint8_t *b, const uint8_t *e, int64_t *midsv_off, int32_t *midsv_len,
                      int64_t *qname_off, int32_t *qname_len, uint8_t *strand) {
    int have_midsv = 0, have_qname = 0, depth = 0, irregular = 0;
    *strand = '+';
    const uint8_t *p = b;
    while (p < e) {
        const uint8_t c = *p;
        if (c == '{' || c == '[') { depth++; p++; continue; }
        if (c == '}' || c == ']') { depth--; p++; continue; }
        if (c != '"') { p++; continue; }
        int esc = 0;
        const uint8_t *ks = p + 1, *ke = string_end(ks, e, &esc);
        if (!ke) return 1;
        p = ke + 1;
        while (p < e && (*p == ' ' || *p == '\t')) p++;
        if (p >= e || *p != ':') continue; /* a string value, not a key */
        p++;
        while (p < e && (*p == ' ' || *p == '\t')) p++;
        if (p >= e) return 1;
        if (*p != '"') continue; /* non-string value: scanned by the main loop */
        int vesc = 0;
        const uint8_t *vs = p + 1, *ve = string_end(vs, e, &vesc);
        if (!ve) return 1;
        p = ve + 1;
        if (depth != 1 || esc) continue; /* only top-level keys count */
        const int64_t klen = ke - ks;
        if (klen == 5 && memcmp(ks, "MIDSV", 5) == 0) {
            if (vesc || have_midsv) irregular = 1;
            *midsv_off = vs - buf; *midsv_len = (int32_t)(ve - vs); have_midsv = 1;
            if (ve - vs > 0x7FFFFFFF) irregular = 1;
        } else if (klen == 5 && memcmp(ks, "QNAME", 5) == 0) {
            if (vesc || have_qname) irregular = 1;
            *qname_off = vs - buf; *qname_len = (int32_t)(ve - vs); have_qname = 1;
        } else if (klen == 6 && memcmp(ks, "STRAND", 6) == 0) {
            if (ve - vs == 1 && (vs[0] == '+' || vs[0] == '-')) *strand = vs[0];
            else irregular = 1;
        }
    }
    return irregular || !have_midsv || !have_qname;
}

/* Fill the per-line arrays (n_lines entries, in file order).  Returns the number of irregular lines (0 = the spans
 * can be used as they are), or -1 when the line count does not match. */
int64_t dj_jsonl_index(const uint8_t *buf, int64_t n, int64_t n_lines, int64_t *midsv_off, int32_t *midsv_len,
                       int64_t *qname_off, int32_t *qname_len, uint8_t *strand, int32_t *has_strand, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    int64_t starts[257], first_line[257];
    /* chunk t = lines that start in [n*t/T, n*(t+1)/T) */
    for (int t = 0; t <= n_threads; t++) {
        int64_t lo = n * t / n_threads;
        if (t > 0 && t < n_threads) {
            const uint8_t *q = lo > 0 ? memchr(buf + lo - 1, '\n', (size_t)(n - lo + 1)) : 0;
            lo = q ? (q - buf) + 1 : n;
        }
        starts[t] = t == n_threads ? n : lo;
    }
    int64_t counts[256];
#pragma omp parallel for num_threads(n_threads) schedule(static)
    for (int t = 0; t < n_threads; t++) {
        int64_t c = 0;
        const uint8_t *p = buf + starts[t], *stop = buf + starts[t + 1];
        while (p < stop) {
            const uint8_t *q = memchr(p, '\n', (size_t)(buf + n - p));
            const uint8_t *e = q ? q : buf + n;
            if (!is_blank_line(p, e)) c++;
            p = e + 1;
        }
        counts[t] = c;
    }
    int64_t acc = 0;
    for (int t = 0; t < n_threads; t++) { first_line[t] = acc; acc += counts[t]; }
    if (acc != n_lines) return -1;
    int64_t irregular = 0;
    int32_t strand_seen = 0;
#pragma omp parallel for num_threads(n_threads) schedule(static) reduction(+ : irregular) reduction(| : strand_seen)
    for (int t = 0; t < n_threads; t++) {
        int64_t k = first_line[t];
        const uint8_t *p = buf + starts[t], *stop = buf + starts[t + 1];
        while (p < stop) {
            const uint8_t *q = memchr(p, '\n', (size_t)(buf + n - p));
            const uint8_t *e = q ? q : buf + n;
            if (!is_blank_line(p, e)) {
                midsv_off[k] = 0; midsv_len[k] = 0; qname_off[k] = 0; qname_len[k] = 0;
                irregular += index_line(buf, p, e, midsv_off + k, midsv_len + k, qname_off + k, qname_len + k, strand + k);
                /* a STRAND key anywhere: the caller defaults to '+' only when no line has one */
                if (memmem(p, (size_t)(e - p), "\"STRAND\"", 8)) strand_seen |= 1;
                k++;
            }
            p = e + 1;
        }
    }
    *has_strand = strand_seen;
    return irregular;
}

/* out[k] = buf[off[k] : off[k] + len[k]] zero-padded to `width` bytes (numpy 'S<width>' rows) */
void dj_gather_fixed(const uint8_t *buf, const int64_t *off, const int32_t *len, int64_t n, int32_t width, uint8_t *out,
                     int n_threads) {
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for num_threads(n_threads) schedule(static)
    for (int64_t k = 0; k < n; k++) {
        uint8_t *o = out + k * (int64_t)width;
        int32_t l = len[k] < width ? len[k] : width;
        memcpy(o, buf + off[k], (size_t)l);
        memset(o + l, 0, (size_t)(width - l));
    }
}
