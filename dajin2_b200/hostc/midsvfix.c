/* The post-fixes the reference applies to freshly transformed MIDSV strings before they become the wire format
 * (preprocess/alignment/midsv_caller.py:88-202, chained at :267-272), as ONE byte-level pass per read:
 *
 *   bit 1  replace_internal_n_to_d   (:108-126, boundaries :88-105)  '=N'-anchored tokens strictly between the
 *                                    leading and the trailing run of such tokens become '-<reference base>'
 *   bit 2  filter_samples_by_n_proportion (:145-155)  keep = n_tokens / tokens * 100 < threshold
 *   bit 4  convert_consecutive_indels_to_match (:163-202)  '-C,-G,+C|+G|=T' -> '=C,=G,=T': an insertion whose
 *                                    inserted elements equal the deletions right before it is an alignment artefact
 *
 * The reference walks the reversed token list and skips over the deletions it has just rewritten; an insertion
 * token can never lie inside another insertion's run of deletions, so every insertion token is decided on the
 * tokens as they were before this step, in any order -- which is what is done here (tests hold it to the real
 * reference on the reference's own vectors and on generated artefacts).  Rows are independent: OpenMP over rows.
 * A fixed row is never longer than its input (the replacements are 2 bytes for tokens of at least 2 bytes, an
 * anchor is part of its token, '-' -> '=' keeps the length), so the output uses the input's row offsets. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { F_N = 1, F_REPL = 2, F_ANCHOR = 4, F_EQ = 8 };

/* utils/midsv_handler.py:8-14 */
static inline int is_n_token(const uint8_t *t, int32_t n) {
    if (n == 2) return t[0] == '=' && (t[1] == 'N' || t[1] == 'n');
    if (n > 2 && t[0] == '+') {
        int32_t a = n;
        while (a > 0 && t[a - 1] != '|') a--; /* a = start of the last '|'-separated piece (0: no '|', whole token) */
        return n - a == 2 && t[a] == '=' && (t[a + 1] == 'N' || t[a + 1] == 'n');
    }
    return 0;
}

static void fix_row(const uint8_t *s, int32_t len, const uint8_t *seq, int64_t seq_len, int mode, double threshold,
                    uint8_t *o, int32_t *out_len, uint8_t *keep, int32_t *st, uint8_t *fl) {
    int32_t T = 0;
    st[0] = 0;
    for (int32_t i = 0; i < len; i++)
        if (s[i] == ',') st[++T] = i + 1;
    T++;
    st[T] = len + 1; /* token t = s[st[t] .. st[t + 1] - 1) ; "".split(",") is one empty token */

    int64_t n_tokens_n = 0;
    for (int32_t t = 0; t < T; t++) {
        fl[t] = is_n_token(s + st[t], st[t + 1] - 1 - st[t]) ? F_N : 0;
        n_tokens_n += fl[t];
    }
    if (mode & 1) { /* :88-105: left = index of the last leading N token, right = index of the first trailing one */
        int32_t left = 0, right = T - 1;
        while (left < T && (fl[left] & F_N)) left++;
        while (right >= 0 && (fl[right] & F_N)) right--;
        left -= 1;
        right += 1;
        const int64_t lim = seq_len < T ? seq_len : T; /* zip(midsv_tags, sequence) */
        for (int64_t j = left + 1; j < right && j < lim; j++)
            if (fl[j] & F_N) {
                fl[j] = F_REPL;
                n_tokens_n--;
            }
    }
    *keep = (mode & 2) ? ((double)n_tokens_n / (double)T * 100.0 < threshold) : 1;

    if (mode & 4)
        for (int32_t p = 0; p < T; p++) {
            if ((fl[p] & F_REPL) || st[p + 1] - 1 == st[p] || s[st[p]] != '+') continue;
            const uint8_t *tok = s + st[p];
            const int32_t n = st[p + 1] - 1 - st[p];
            int32_t k = 0;
            for (int32_t i = 0; i < n; i++) k += tok[i] == '|';
            if (k == 0 || p - k < 0) continue; /* nothing inserted / fewer tokens before it than inserted elements */
            int ok = 1;
            int32_t a = 0;
            for (int32_t e = 0; e < k && ok; e++) { /* element e (forward order) against token p - k + e */
                int32_t b = a;
                while (tok[b] != '|') b++;
                int32_t ea = a;
                while (ea < b && tok[ea] == '+') ea++; /* lstrip("+") */
                const int32_t q = p - k + e;
                uint8_t two[2];
                const uint8_t *d;
                int32_t dn;
                if (fl[q] & F_REPL) {
                    two[0] = '-';
                    two[1] = seq[q];
                    d = two;
                    dn = 2;
                } else {
                    d = s + st[q];
                    dn = st[q + 1] - 1 - st[q];
                }
                if (dn == 0 || d[0] != '-') ok = 0;
                else {
                    int32_t da = 0;
                    while (da < dn && d[da] == '-') da++; /* lstrip("-") */
                    ok = (dn - da == b - ea) && memcmp(d + da, tok + ea, (size_t)(b - ea)) == 0;
                }
                a = b + 1;
            }
            if (!ok) continue;
            fl[p] |= F_ANCHOR;
            for (int32_t q = p - k; q < p; q++) fl[q] |= F_EQ;
        }

    int32_t w = 0;
    for (int32_t t = 0; t < T; t++) {
        if (t) o[w++] = ',';
        const uint8_t *tok = s + st[t];
        int32_t n = st[t + 1] - 1 - st[t];
        uint8_t two[2];
        if (fl[t] & F_REPL) {
            two[0] = '-';
            two[1] = seq[t];
            tok = two;
            n = 2;
        } else if (fl[t] & F_ANCHOR) {
            int32_t a = n;
            while (tok[a - 1] != '|') a--;
            tok += a;
            n -= a;
        }
        if (fl[t] & F_EQ)
            for (int32_t i = 0; i < n; i++) o[w++] = tok[i] == '-' ? '=' : tok[i]; /* str.replace("-", "=") */
        else {
            memcpy(o + w, tok, (size_t)n);
            w += n;
        }
    }
    *out_len = w;
}

/* rows r = text[off[r] .. off[r] + len[r]); out[off[r] .. off[r] + out_len[r]) receives the fixed row.
 * Returns 0, or -1 when the scratch memory could not be allocated. */
int64_t dj_midsv_postfix(const uint8_t *text, const int64_t *off, const int32_t *len, int64_t n, const uint8_t *seq,
                         int64_t seq_len, int32_t mode, double threshold, uint8_t *out, int32_t *out_len,
                         uint8_t *keep, int threads) {
    int32_t longest = 0;
    for (int64_t r = 0; r < n; r++)
        if (len[r] > longest) longest = len[r];
    if (threads < 1) threads = 1;
    int64_t failed = 0;
#pragma omp parallel num_threads(threads)
    {
        int32_t *st = (int32_t *)malloc(((size_t)longest + 3) * sizeof(int32_t));
        uint8_t *fl = (uint8_t *)malloc((size_t)longest + 3);
        if (!st || !fl) {
#pragma omp atomic write
            failed = 1;
        }
#pragma omp barrier
        if (!failed) {
#pragma omp for schedule(dynamic, 64)
            for (int64_t r = 0; r < n; r++)
                fix_row(text + off[r], len[r], seq, seq_len, mode, threshold, out + off[r], out_len + r, keep + r, st, fl);
        }
        free(st);
        free(fl);
    }
    return failed ? -1 : 0;
}
