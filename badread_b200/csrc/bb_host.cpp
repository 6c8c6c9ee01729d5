// bb_host.cpp — host-side table builder helpers of libbadread_b200 (no GPU involved).
//
// Badread builds its error-model table at load time by aligning every alternative k-mer to its k-mer with
// edlib (badread/error_model.py:111-133 -> align_kmers :179-229, ~425k tiny alignments per shipped model).
// That is load-time table preparation, not the per-read hot path; it runs here on the host and its output
// (encoded slot strings) is what bb_upload_error_model ships to HBM.
//
// The alignment rule is edlib's (third-party, not vendored by the reference): NW edit distance, path chosen by
// traceback from the bottom-right cell preferring UP ('I', consumes a query char) over LEFT ('D') over the
// diagonal ('=' / 'X').  Inputs here are tiny (k-mers), so edlib's traceback branch always applies.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/badread_b200.h"

namespace {

// Expanded CIGAR for small inputs; returns the edit distance. Full (n+1)x(m+1) matrix.
int small_nw_path(const uint8_t *q, int n, const uint8_t *t, int m, std::vector<uint8_t> &ops) {
    const int w = m + 1;
    std::vector<int32_t> D(static_cast<size_t>(n + 1) * w);
    for (int j = 0; j <= m; j++) D[j] = j;
    for (int i = 1; i <= n; i++) {
        D[static_cast<size_t>(i) * w] = i;
        for (int j = 1; j <= m; j++) {
            int32_t best = D[static_cast<size_t>(i - 1) * w + j - 1] + (q[i - 1] != t[j - 1] ? 1 : 0);
            best = std::min(best, D[static_cast<size_t>(i - 1) * w + j] + 1);
            best = std::min(best, D[static_cast<size_t>(i) * w + j - 1] + 1);
            D[static_cast<size_t>(i) * w + j] = best;
        }
    }
    ops.clear();
    int i = n, j = m;
    while (i > 0 && j > 0) {
        const int32_t cur = D[static_cast<size_t>(i) * w + j];
        if (D[static_cast<size_t>(i - 1) * w + j] + 1 == cur) { ops.push_back('I'); i--; }
        else if (D[static_cast<size_t>(i) * w + j - 1] + 1 == cur) { ops.push_back('D'); j--; }
        else { ops.push_back(D[static_cast<size_t>(i - 1) * w + j - 1] == cur ? '=' : 'X'); i--; j--; }
    }
    for (; i > 0; i--) ops.push_back('I');
    for (; j > 0; j--) ops.push_back('D');
    std::reverse(ops.begin(), ops.end());
    return D[static_cast<size_t>(n) * w + m];
}

bool edlib_uses_traceback(int64_t n, int64_t m) {
    return 20 * ((n + 63) / 64) * m + 8 * m < 1024 * 1024;
}

}  // namespace

extern "C" int bb_host_align_path(const uint8_t *query, int32_t q_len, const uint8_t *target, int32_t t_len,
                                  uint8_t *ops_out, int64_t ops_cap, int64_t *n_ops, int32_t *distance) {
    if (q_len <= 0 || t_len <= 0 || !query || !target) return BB_ERR_ARG;
    if (static_cast<int64_t>(q_len) * t_len > (1 << 22) || !edlib_uses_traceback(q_len, t_len)) return BB_ERR_ARG;
    std::vector<uint8_t> ops;
    const int d = small_nw_path(query, q_len, target, t_len, ops);
    if (distance) *distance = d;
    if (n_ops) *n_ops = static_cast<int64_t>(ops.size());
    if (static_cast<int64_t>(ops.size()) > ops_cap) return BB_ERR_CAPACITY;
    if (ops_out && !ops.empty()) std::memcpy(ops_out, ops.data(), ops.size());
    return BB_OK;
}

// error_model.align_kmers (error_model.py:179-229) for every (kmer, alt) pair of a model file.
extern "C" int bb_host_align_kmers(int k, int32_t n_alts, const uint8_t *kmers, const uint8_t *alts,
                                   const int32_t *alt_off, uint32_t *slots_out, uint8_t *flags_out, uint8_t *pool,
                                   int64_t pool_cap, int64_t *pool_len) {
    if (k < 3 || k > 32 || n_alts < 0 || !kmers || !alts || !alt_off || !slots_out || !flags_out || !pool_len)
        return BB_ERR_ARG;
    std::vector<uint8_t> ops;
    std::vector<std::vector<uint8_t>> result(static_cast<size_t>(k));
    for (int32_t a = 0; a < n_alts; a++) {
        const uint8_t *kmer = kmers + static_cast<int64_t>(a) * k;
        const uint8_t *alt = alts + alt_off[a];
        const int alt_len = alt_off[a + 1] - alt_off[a];
        if (alt_len < 2 || kmer[0] != alt[0] || kmer[k - 1] != alt[alt_len - 1]) return BB_ERR_ARG;  // :188-195
        for (auto &r : result) r.clear();
        result[0].push_back(kmer[0]);
        result[static_cast<size_t>(k) - 1].push_back(kmer[k - 1]);
        const uint8_t *ik = kmer + 1;
        const int ik_len = k - 2;
        const uint8_t *ia = alt + 1;
        const int ia_len = alt_len - 2;
        if (ia_len == 0) ops.assign(static_cast<size_t>(ik_len), 'D');  // :198-200
        else small_nw_path(ia, ia_len, ik, ik_len, ops);                 // edlib.align(alt, kmer) :202
        int kmer_pos = 0, alt_pos = 0;
        for (uint8_t op : ops) {
            if (op == '=' || op == 'X') {
                result[static_cast<size_t>(kmer_pos) + 1].assign(1, ia[alt_pos]);
                alt_pos++; kmer_pos++;
            } else if (op == 'D') {
                result[static_cast<size_t>(kmer_pos) + 1].clear();
                kmer_pos++;
            } else {  // 'I': attaches to the slot before (:220-222)
                result[static_cast<size_t>(kmer_pos)].push_back(ia[alt_pos]);
                alt_pos++;
            }
        }
        if (result[0].size() == 2) {  // :225-228
            const uint8_t inserted = result[0][1];
            result[0].resize(1);
            result[1].insert(result[1].begin(), inserted);
        }
        bool same = true;
        int joined = 0;
        for (int j = 0; j < k; j++) {
            const auto &s = result[static_cast<size_t>(j)];
            for (uint8_t c : s) {
                if (joined >= k || kmer[joined] != c) same = false;
                joined++;
            }
            if (s.size() > 255) return BB_ERR_ARG;
            uint32_t enc = static_cast<uint32_t>(s.size());
            if (s.size() <= 3) {
                for (size_t c = 0; c < s.size(); c++) enc |= static_cast<uint32_t>(s[c]) << (8 * (c + 1));
            } else {
                if (*pool_len + static_cast<int64_t>(s.size()) > pool_cap || *pool_len >= (1 << 24)) return BB_ERR_CAPACITY;
                enc |= static_cast<uint32_t>(*pool_len) << 8;
                std::memcpy(pool + *pool_len, s.data(), s.size());
                *pool_len += static_cast<int64_t>(s.size());
            }
            slots_out[static_cast<int64_t>(a) * k + j] = enc;
        }
        if (joined != k) same = false;
        flags_out[a] = same ? 1 : 0;
    }
    return BB_OK;
}
