/* Host-side scan of a SAM file for the cs tag -> MIDSV path (dajin2_b200/alignment.py):
 * field spans of every alignment record (QNAME, FLAG, RNAME, POS, the long-form cs:Z: tag) found by threads over
 * whole lines, and the cs tags gathered in read order into the buffer dj_cs_midsv_* take -- what fileio.read_sam
 * (utils/fileio.py) + midsv's record dictionaries do with one Python list per line.  Nothing here interprets a tag
 * except dj_cs_pack's optional deletion-then-insertion fix, the cs-level form of convert_consecutive_indels_to_match
 * (preprocess/alignment/midsv_caller.py:163-202; rule and proof against the reference in alignment.py / the tests). */
#include <stdint.h>
#include <string.h>

static inline int sam_is_record(const uint8_t *b, const uint8_t *e) {
    while (b < e && (e[-1] == '\r' || e[-1] == ' ')) e--;
    return b < e && *b != '@';
}

static void chunk_starts(const uint8_t *buf, int64_t n, int n_threads, int64_t *starts) {
    for (int t = 0; t <= n_threads; t++) {
        int64_t lo = n * t / n_threads;
        if (t > 0 && t < n_threads) {
            const uint8_t *q = lo > 0 ? memchr(buf + lo - 1, '\n', (size_t)(n - lo + 1)) : 0;
            lo = q ? (q - buf) + 1 : n;
        }
        starts[t] = t == n_threads ? n : lo;
    }
}

/* number of alignment records (lines that are neither blank nor headers) */
int64_t dj_sam_count(const uint8_t *buf, int64_t n, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    int64_t starts[257], total = 0;
    chunk_starts(buf, n, n_threads, starts);
#pragma omp parallel for num_threads(n_threads) schedule(static) reduction(+ : total)
    for (int t = 0; t < n_threads; t++) {
        const uint8_t *p = buf + starts[t], *stop = buf + starts[t + 1];
        int64_t c = 0;
        while (p < stop) {
            const uint8_t *q = memchr(p, '\n', (size_t)(buf + n - p));
            const uint8_t *e = q ? q : buf + n;
            c += sam_is_record(p, e);
            p = e + 1;
        }
        total += c;
    }
    return total;
}

static inline int parse_uint(const uint8_t *b, const uint8_t *e, int64_t *out) {
    if (b >= e || e - b > 12) return 1;
    int64_t v = 0;
    for (; b < e; b++) {
        if (*b < '0' || *b > '9') return 1;
        v = v * 10 + (*b - '0');
    }
    *out = v;
    return 0;
}

/* one record: returns 1 when it is malformed (fewer than 11 fields, FLAG / POS not numbers) */
static int index_record(const uint8_t *buf, const uint8_t *b, const uint8_t *e, int64_t *qname_off, int32_t *qname_len,
                        int32_t *flag, int64_t *rname_off, int32_t *rname_len, int32_t *pos, int64_t *cs_off,
                        int32_t *cs_len) {
    while (b < e && e[-1] == '\r') e--;
    const uint8_t *f[12];
    int nf = 0;
    f[nf++] = b;
    for (const uint8_t *p = b; nf < 12;) {
        const uint8_t *q = memchr(p, '\t', (size_t)(e - p));
        if (!q) break;
        f[nf++] = q + 1;
        p = q + 1;
    }
    *cs_off = 0;
    *cs_len = -1;
    if (nf < 11) return 1;
    int64_t v_flag = 0, v_pos = 0;
    if (parse_uint(f[1], f[2] - 1, &v_flag) || parse_uint(f[3], f[4] - 1, &v_pos) || v_flag > 0xFFFF || v_pos > 0x7FFFFFFF)
        return 1;
    *qname_off = f[0] - buf;
    *qname_len = (int32_t)(f[1] - 1 - f[0]);
    *flag = (int32_t)v_flag;
    *rname_off = f[2] - buf;
    *rname_len = (int32_t)(f[3] - 1 - f[2]);
    *pos = (int32_t)v_pos;
    if (nf == 12) { /* optional fields: the first that starts with "cs:Z:" */
        for (const uint8_t *p = f[11]; p < e;) {
            const uint8_t *q = memchr(p, '\t', (size_t)(e - p));
            const uint8_t *fe = q ? q : e;
            if (fe - p >= 5 && memcmp(p, "cs:Z:", 5) == 0) {
                if (fe - p - 5 > 0x7FFFFFF0) return 1;
                *cs_off = p + 5 - buf;
                *cs_len = (int32_t)(fe - p - 5);
                break;
            }
            p = fe + 1;
        }
    }
    return 0;
}

/* Field spans of the n_records alignment records, in file order; cs_len[k] = -1 when record k has no cs tag.
 * Returns the number of malformed records, or -1 when the record count does not match. */
int64_t dj_sam_index(const uint8_t *buf, int64_t n, int64_t n_records, int64_t *qname_off, int32_t *qname_len, int32_t *flag,
                     int64_t *rname_off, int32_t *rname_len, int32_t *pos, int64_t *cs_off, int32_t *cs_len, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    int64_t starts[257], first[257], counts[256];
    chunk_starts(buf, n, n_threads, starts);
#pragma omp parallel for num_threads(n_threads) schedule(static)
    for (int t = 0; t < n_threads; t++) {
        const uint8_t *p = buf + starts[t], *stop = buf + starts[t + 1];
        int64_t c = 0;
        while (p < stop) {
            const uint8_t *q = memchr(p, '\n', (size_t)(buf + n - p));
            const uint8_t *e = q ? q : buf + n;
            c += sam_is_record(p, e);
            p = e + 1;
        }
        counts[t] = c;
    }
    int64_t acc = 0;
    for (int t = 0; t < n_threads; t++) { first[t] = acc; acc += counts[t]; }
    if (acc != n_records) return -1;
    int64_t malformed = 0;
#pragma omp parallel for num_threads(n_threads) schedule(static) reduction(+ : malformed)
    for (int t = 0; t < n_threads; t++) {
        int64_t k = first[t];
        const uint8_t *p = buf + starts[t], *stop = buf + starts[t + 1];
        while (p < stop) {
            const uint8_t *q = memchr(p, '\n', (size_t)(buf + n - p));
            const uint8_t *e = q ? q : buf + n;
            if (sam_is_record(p, e)) {
                qname_off[k] = rname_off[k] = 0;
                qname_len[k] = rname_len[k] = flag[k] = pos[k] = 0;
                malformed += index_record(buf, p, e, qname_off + k, qname_len + k, flag + k, rname_off + k, rname_len + k,
                                          pos + k, cs_off + k, cs_len + k);
                k++;
            }
            p = e + 1;
        }
    }
    return malformed;
}

static inline int is_alpha(uint8_t c) { return (uint8_t)((c | 0x20) - 'a') < 26; }

/* the deletion-then-insertion fix on one tag: returns the new length (<= n); src and dst do not overlap */
static int64_t fix_pairs(const uint8_t *src, int64_t n, uint8_t *dst) {
    int64_t o = 0, i = 0;
    while (i < n) {
        if (src[i] != '-') {
            dst[o++] = src[i++];
            continue;
        }
        int64_t d0 = i + 1, d1 = d0; /* deleted bases [d0, d1) */
        while (d1 < n && is_alpha(src[d1])) d1++;
        int64_t i1 = d1 + 1; /* inserted bases [d1 + 1, i1) when src[d1] == '+' */
        if (d1 < n && src[d1] == '+') while (i1 < n && is_alpha(src[i1])) i1++;
        const int64_t m = d1 - d0, k = (d1 < n && src[d1] == '+') ? i1 - d1 - 1 : 0;
        int convert = 0;
        if (m > 0 && k > 0) {
            int64_t p = i - 1; /* the operator in front of the deletion run, in the ORIGINAL tag */
            while (p >= 0 && is_alpha(src[p])) p--;
            const int64_t usable = m - ((p >= 0 && src[p] == '+') ? 1 : 0);
            convert = k <= usable && memcmp(src + d1 - k, src + d1 + 1, (size_t)k) == 0;
        }
        if (!convert) { /* copy the deletion run as it is; an insertion behind it is handled by the main loop */
            memcpy(dst + o, src + i, (size_t)(d1 - i));
            o += d1 - i;
            i = d1;
            continue;
        }
        if (m > k) {
            dst[o++] = '-';
            memcpy(dst + o, src + d0, (size_t)(m - k));
            o += m - k;
        }
        dst[o++] = '=';
        memcpy(dst + o, src + d1 + 1, (size_t)k);
        o += k;
        i = i1;
    }
    return o;
}

/* out[out_off[j] .. out_off[j + 1]) = the cs tag of record order[j] (out_off = running sum of the tags' lengths, n + 1
 * entries, made by the caller); with fix != 0 the deletion-then-insertion fix is applied and the bytes it saves are
 * NUL at the end of the tag's slot (dj_cs_midsv_* skip NUL bytes). */
void dj_cs_pack(const uint8_t *buf, const int64_t *cs_off, const int32_t *cs_len, const int64_t *order, int64_t n, int fix,
                uint8_t *out, const int64_t *out_off, int n_threads) {
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for num_threads(n_threads) schedule(dynamic, 256)
    for (int64_t j = 0; j < n; j++) {
        const int64_t r = order[j], len = cs_len[r];
        const uint8_t *src = buf + cs_off[r];
        uint8_t *dst = out + out_off[j];
        if (!fix || !memchr(src, '+', (size_t)len) || !memchr(src, '-', (size_t)len)) {
            memcpy(dst, src, (size_t)len);
            continue;
        }
        const int64_t m = fix_pairs(src, len, dst);
        memset(dst + m, 0, (size_t)(len - m));
    }
}
