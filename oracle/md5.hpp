// ORACLE — TEST INFRASTRUCTURE ONLY (not part of the product path).
//
// MD5 (RFC 1321) and UUID version 3 in the X.500 namespace, as used for read
// names at src/simulate/mod.rs:199-203 via the un-vendored `uuid 0.8` crate
// (Uuid::new_v3(&NAMESPACE_X500, name).to_hyphenated()).  The reference holds
// no golden vector for a read name: parity of this piece is UNPINNED
// (SURVEY.md §8c); the MD5 itself is checked against RFC 1321's test suite.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>

namespace orc {

inline void md5(const uint8_t* msg, size_t len, uint8_t out[16]) {
  static const uint32_t K[64] = {
      0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
      0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
      0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
      0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
      0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
      0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
      0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
      0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
  static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9,  14, 20, 5, 9,
                            14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                            4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
  uint32_t a0 = 0x67452301, b0 = 0xefcdab89, c0 = 0x98badcfe, d0 = 0x10325476;
  size_t padded = ((len + 8) / 64 + 1) * 64;
  std::string buf((const char*)msg, len);
  buf.resize(padded, '\0');
  buf[len] = (char)0x80;
  uint64_t bits = (uint64_t)len * 8;
  for (int i = 0; i < 8; ++i) buf[padded - 8 + i] = (char)(bits >> (8 * i));
  for (size_t off = 0; off < padded; off += 64) {
    uint32_t M[16];
    for (int i = 0; i < 16; ++i) {
      const uint8_t* p = (const uint8_t*)buf.data() + off + 4 * i;
      M[i] = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    }
    uint32_t A = a0, B = b0, C = c0, D = d0;
    for (int i = 0; i < 64; ++i) {
      uint32_t F;
      int g;
      if (i < 16) { F = (B & C) | (~B & D); g = i; }
      else if (i < 32) { F = (D & B) | (~D & C); g = (5 * i + 1) % 16; }
      else if (i < 48) { F = B ^ C ^ D; g = (3 * i + 5) % 16; }
      else { F = C ^ (B | ~D); g = (7 * i) % 16; }
      F = F + A + K[i] + M[g];
      A = D; D = C; C = B;
      B = B + ((F << S[i]) | (F >> (32 - S[i])));
    }
    a0 += A; b0 += B; c0 += C; d0 += D;
  }
  uint32_t h[4] = {a0, b0, c0, d0};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) out[4 * i + j] = (uint8_t)(h[i] >> (8 * j));
}

// Uuid::new_v3(&Uuid::NAMESPACE_X500, name).to_hyphenated() -> lower-case text
inline std::string uuid_v3_x500(const uint8_t* name, size_t len) {
  static const uint8_t NS[16] = {0x6b, 0xa7, 0xb8, 0x14, 0x9d, 0xad, 0x11, 0xd1,
                                 0x80, 0xb4, 0x00, 0xc0, 0x4f, 0xd4, 0x30, 0xc8};
  std::string msg((const char*)NS, 16);
  msg.append((const char*)name, len);
  uint8_t d[16];
  md5((const uint8_t*)msg.data(), msg.size(), d);
  d[6] = (d[6] & 0x0f) | 0x30;  // version 3
  d[8] = (d[8] & 0x3f) | 0x80;  // RFC 4122 variant
  static const char* hex = "0123456789abcdef";
  std::string s;
  for (int i = 0; i < 16; ++i) {
    if (i == 4 || i == 6 || i == 8 || i == 10) s += '-';
    s += hex[d[i] >> 4];
    s += hex[d[i] & 15];
  }
  return s;
}

}  // namespace orc
