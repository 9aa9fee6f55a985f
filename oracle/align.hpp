// ORACLE — TEST INFRASTRUCTURE ONLY (not part of the product path).
//
// Restatement of src/alignment.rs:9-41.  The reference delegates to the
// un-vendored crate `bio 0.33`:
//   * edit_distance -> bio::alignment::distance::levenshtein      (alignment.rs:9-11)
//   * align         -> bio::alignment::pairwise::Aligner::with_capacity(.., 0, 0,
//                      |a,b| (a==b) as i32).global(query, target)  (alignment.rs:13-22)
// With gap_open = gap_extend = 0, match = 1, mismatch = 0 the affine aligner
// degenerates to a max-matches (LCS) alignment with a fixed traceback
// tie-break.  Pinned by src/alignment.rs:47-86 and every CIGAR
// golden in src/simulate/error.rs.
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

namespace orc {

typedef std::vector<uint8_t> Bytes;

inline uint64_t edit_distance(const uint8_t* a, size_t n, const uint8_t* b, size_t m) {
  std::vector<uint64_t> prev(m + 1), cur(m + 1);
  for (size_t j = 0; j <= m; ++j) prev[j] = j;
  for (size_t i = 1; i <= n; ++i) {
    cur[0] = i;
    for (size_t j = 1; j <= m; ++j) {
      uint64_t sub = prev[j - 1] + (a[i - 1] != b[j - 1]);
      cur[j] = std::min(sub, std::min(prev[j], cur[j - 1]) + 1);
    }
    std::swap(prev, cur);
  }
  return prev[m];
}

// align(query, target) -> (edits, cigar over "=XID")
inline uint64_t align(const uint8_t* q, size_t m, const uint8_t* t, size_t n, Bytes& cigar) {
  // like the reference, buffers are allocated per call (src/alignment.rs:14-20)
  std::vector<int32_t> S((m + 1) * (n + 1), 0);
  const size_t W = n + 1;
  for (size_t i = 1; i <= m; ++i)
    for (size_t j = 1; j <= n; ++j) {
      int32_t diag = S[(i - 1) * W + (j - 1)] + (q[i - 1] == t[j - 1] ? 1 : 0);
      int32_t up = S[(i - 1) * W + j];
      int32_t left = S[i * W + (j - 1)];
      S[i * W + j] = std::max(diag, std::max(up, left));
    }
  cigar.clear();
  cigar.reserve(std::max(m, n));
  size_t i = m, j = n;
  while (i > 0 || j > 0) {
    if (i == 0) { cigar.push_back('D'); --j; continue; }
    if (j == 0) { cigar.push_back('I'); --i; continue; }
    bool eq = q[i - 1] == t[j - 1];
    int32_t diag = S[(i - 1) * W + (j - 1)] + (eq ? 1 : 0);
    int32_t up = S[(i - 1) * W + j];
    int32_t left = S[i * W + (j - 1)];
    if (diag >= left && diag >= up) { cigar.push_back(eq ? '=' : 'X'); --i; --j; }
    else if (up >= left) { cigar.push_back('I'); --i; }
    else { cigar.push_back('D'); --j; }
  }
  std::reverse(cigar.begin(), cigar.end());
  uint64_t edits = 0;
  for (uint8_t c : cigar) edits += (c != '=');
  return edits;
}

}  // namespace orc
