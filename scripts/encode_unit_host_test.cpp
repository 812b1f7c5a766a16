// Host check of dajin2_b200/csrc/encode_unit.cuh (the byte-parallel arithmetic of the unit encoder) against a plain
// tokenizer:  g++ -O2 -std=c++17 -o /tmp/encode_unit_host_test scripts/encode_unit_host_test.cpp && /tmp/encode_unit_host_test
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../dajin2_b200/csrc/encode_unit.cuh"

static int base_id(char c) {
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2;
                 case 'T': case 't': return 3; case 'N': case 'n': return 4; }
    return -1;
}

// code byte of one token (include/dajin2_b200.h), 0x7F when outside the grammar
static int token_code(const std::string &t) {
    size_t bar = t.rfind('|');
    std::string a = bar == std::string::npos ? t : t.substr(bar + 1);
    int ins = bar == std::string::npos ? 0 : 0x80;
    if (a.size() == 2 && (a[0] == '=' || a[0] == '-') && base_id(a[1]) >= 0)
        return ins | ((a[0] == '-' ? 10 : 0) + 5 * (a[1] >= 'a') + base_id(a[1]));
    if (a.size() == 3 && a[0] == '*' && base_id(a[1]) >= 0 && base_id(a[2]) >= 0 && ((a[1] >= 'a') == (a[2] >= 'a')))
        return ins | (20 + 25 * (a[1] >= 'a') + 5 * base_id(a[1]) + base_id(a[2]));
    return 0x7F;
}

int main() {
    std::mt19937 rng(12345);
    const char *vocab[] = {"=A", "=C", "=G", "=T", "=N", "-A", "-C", "-G", "-T", "-N", "*AG", "*TC", "=a", "-g", "+A|=C",
                           "+T|+G|*CA", "*ag", "+G|+G|+G|+G|+G|+G|+G|+G|+G|+G|+G|+G|+G|+G|+G|+G|+G|=T", "=R", "==", "A"};
    const int n_vocab = sizeof(vocab) / sizeof(vocab[0]);
    long units = 0, plain_units = 0, missed = 0, head_units = 0;
    for (int iter = 0; iter < 4000; ++iter) {
        const int rare = 1 + (int)(rng() % 60);  // one token in `rare` is not a plain match
        std::string row;
        std::vector<int> codes;
        const int n_tok = 200 + (int)(rng() % 400);
        for (int i = 0; i < n_tok; ++i) {
            std::string t = (rng() % rare) ? vocab[rng() % 4] : vocab[rng() % n_vocab];
            if (iter % 7 == 3 && (rng() % 3) == 0) t = vocab[5 + rng() % 5];  // deletion-rich rows
            codes.push_back(token_code(t));
            row += t;
            row += ',';
        }
        const int head = (int)(rng() % 16);
        std::string buf(head + 4, ' ');  // 4 bytes of filler in front so that prev words exist
        buf += row;
        buf += std::string(64, ' ');
        // comma rank of every byte
        std::vector<int> rank(buf.size(), 0);
        int r = 0;
        for (size_t p = 0; p < buf.size(); ++p) { rank[p] = r; if (buf[p] == ',') ++r; }
        for (size_t start = 4 + ((4 - (head & 3)) & 3); start + 48 <= head + 4 + row.size(); start += 4) {
            // only word-aligned (relative to buf + 0) unit starts are meaningful: make `start` a multiple of 4
            if (start % 4) continue;
            uint32_t w[12], prev;
            memcpy(w, buf.data() + start, 48);
            memcpy(&prev, buf.data() + start - 4, 4);
            int n_commas = 0;
            for (int p = 0; p < 48; ++p) n_commas += buf[start + p] == ',';
            const dju::Phase ph = dju::phase_entry(dju::phase_index(w[0]));
            const dju::UnitOut u = dju::decode_unit(prev, w, ph);
            const bool plain = u.bad == 0;
            // truth: the tokens closed by the unit's commas
            const int first = rank[start];
            bool truth_plain = n_commas == 16;
            int n_del = 0;
            if (truth_plain) {
                int c0 = -1;
                for (int p = 0; p < 3; ++p) if (buf[start + p] == ',') { c0 = p; break; }
                truth_plain = c0 >= 0 && buf[start + c0 - 3] == ',';
                for (int j = 0; truth_plain && j < 16; ++j) {
                    const int p = c0 + 3 * j;
                    if (buf[start + p] != ',') truth_plain = false;
                    const char op = buf[start + p - 2], x = buf[start + p - 1];
                    if (!((op == '=' || op == '-') && (x == 'A' || x == 'C' || x == 'G' || x == 'T' || x == 'N')))
                        truth_plain = false;
                    if (op == '-') ++n_del;
                }
            }
            ++units;
            if (plain) {
                ++plain_units;
                if (!truth_plain) { printf("false plain at iter %d start %zu\n", iter, start); return 1; }
                const uint8_t *cb = (const uint8_t *)u.c;
                for (int j = 0; j < 16; ++j)
                    if (cb[j] != codes[first + j]) { printf("code mismatch iter %d start %zu tok %d: %d vs %d\n", iter, start, j, cb[j], codes[first + j]); return 1; }
                if ((int)u.n_del != n_del) { printf("n_del mismatch\n"); return 1; }
                for (uint32_t a = 0; a < 4; ++a) {
                    uint32_t o5[5];
                    uint8_t flat[20];
                    dju::rotate_codes(u.c, a, o5);
                    memcpy(flat, o5, 20);
                    for (uint32_t j = 0; j < 20; ++j) {
                        const int want = (j >= a && j < a + 16) ? cb[j - a] : 0;
                        if (flat[j] != want) { printf("rotate mismatch\n"); return 1; }
                    }
                }
            } else if (truth_plain) {
                ++missed;
            }
            if (u.bad_head == 0) {
                // the tokens the first 16-byte chunk closes are plain and the first eight codes are theirs (and more)
                int n0 = 0, nd = 0;
                for (int p = 0; p < 16; ++p) n0 += buf[start + p] == ',';
                if (n0 != (int)ph.n0) { printf("n0 mismatch %d vs %u\n", n0, ph.n0); return 1; }
                const uint8_t *cb = (const uint8_t *)u.c;
                for (int j = 0; j < 8; ++j)
                    if (cb[j] != codes[first + j]) { printf("head code mismatch\n"); return 1; }
                for (int j = 0; j < n0; ++j) nd += codes[first + j] >= 10 && codes[first + j] < 20;
                if (nd != (int)u.n_del_head) { printf("n_del_head mismatch %d vs %u\n", nd, u.n_del_head); return 1; }
                ++head_units;
            }
        }
    }
    printf("units %ld, plain %ld (%.1f %%), head-plain %ld, truth-plain units not taken %ld\n", units, plain_units,
           100.0 * plain_units / units, head_units, missed);
    if (missed) { printf("FAIL: a plain unit was not recognised\n"); return 1; }
    printf("OK\n");
    return 0;
}
