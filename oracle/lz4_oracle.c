/*
 * lz4_oracle.c -- scalar CPU inflater for LZ4 frames.  TEST INFRASTRUCTURE ONLY
 * (same rules as sbmc_oracle.c: only tests/, smoke() and bench.py's CPU legs may
 * use it; nothing under sbmc_b200/ does).
 *
 * What it restates: `lz4.frame.decompress(buf)` as called by the reference's tile
 * reader (sbmc/datasets.py:570-579).  The algorithm lives in a third-party
 * dependency that is NOT under /root/reference: lz4 (python package `lz4`,
 * unpinned in setup.py:104; C library liblz4-dev in dockerfiles/cpu-sbmc.dockerfile:16
 * and cuda-sbmc.dockerfile:19; the writer is LZ4F_compressFrame with default
 * preferences, pbrt_patches/sbmc_pbrt.diff:6140-6158).  This file restates the
 * published LZ4 frame format (v1.6.x) and block format: sequential, one byte at
 * a time, verifying the xxHash32 header, block and content checksums.
 *
 * PARITY PIN: tests/test_tiles.py checks it against frames produced by the real
 * liblz4 (LZ4F_compressFrame through ctypes on /usr/lib/x86_64-linux-gnu/liblz4.so.1,
 * the library the reference links) committed under tests/golden/tiles/, and
 * against pyarrow's bundled lz4 frame codec.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define LZ4O_OK 0
#define LZ4O_BAD_MAGIC 1
#define LZ4O_BAD_HEADER 2
#define LZ4O_TRUNCATED 3
#define LZ4O_OVERFLOW 4
#define LZ4O_BAD_OFFSET 5
#define LZ4O_BLOCK_TOO_LARGE 7
#define LZ4O_BAD_CHECKSUM 8

static uint32_t rd32(const uint8_t *p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

/* xxHash32 (public specification), seed 0 is all the frame format uses. */
static uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
#define P1 2654435761u
#define P2 2246822519u
#define P3 3266489917u
#define P4 668265263u
#define P5 374761393u

uint32_t sbmc_oracle_xxh32(const uint8_t *p, int64_t len, uint32_t seed) {
  const uint8_t *end = p + len;
  uint32_t h;
  if (len >= 16) {
    uint32_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
    const uint8_t *limit = end - 16;
    do {
      v1 = rotl(v1 + rd32(p) * P2, 13) * P1;
      v2 = rotl(v2 + rd32(p + 4) * P2, 13) * P1;
      v3 = rotl(v3 + rd32(p + 8) * P2, 13) * P1;
      v4 = rotl(v4 + rd32(p + 12) * P2, 13) * P1;
      p += 16;
    } while (p <= limit);
    h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
  } else {
    h = seed + P5;
  }
  h += (uint32_t)len;
  while (p + 4 <= end) {
    h = rotl(h + rd32(p) * P3, 17) * P4;
    p += 4;
  }
  while (p < end) {
    h = rotl(h + (*p) * P5, 11) * P1;
    ++p;
  }
  h ^= h >> 15;
  h *= P2;
  h ^= h >> 13;
  h *= P3;
  h ^= h >> 16;
  return h;
}

static int block(const uint8_t *ip, const uint8_t *iend, uint8_t *dst, int64_t *op_io, int64_t cap,
                 int64_t window) {
  int64_t op = *op_io;
  while (1) {
    if (ip >= iend) return LZ4O_TRUNCATED;
    unsigned token = *ip++;
    int64_t lit = token >> 4;
    if (lit == 15) {
      unsigned b;
      do {
        if (ip >= iend) return LZ4O_TRUNCATED;
        b = *ip++;
        lit += b;
      } while (b == 255);
    }
    if (lit > iend - ip) return LZ4O_TRUNCATED;
    if (lit > cap - op) return LZ4O_OVERFLOW;
    memcpy(dst + op, ip, (size_t)lit);
    ip += lit;
    op += lit;
    if (ip == iend) break;
    if (iend - ip < 2) return LZ4O_TRUNCATED;
    int64_t off = ip[0] | (ip[1] << 8);
    ip += 2;
    int64_t ml = token & 15;
    if (ml == 15) {
      unsigned b;
      do {
        if (ip >= iend) return LZ4O_TRUNCATED;
        b = *ip++;
        ml += b;
      } while (b == 255);
    }
    ml += 4;
    if (off == 0 || off > op - window) return LZ4O_BAD_OFFSET;
    if (ml > cap - op) return LZ4O_OVERFLOW;
    for (int64_t i = 0; i < ml; ++i) dst[op + i] = dst[op + i - off]; /* overlap = run-length */
    op += ml;
  }
  *op_io = op;
  return LZ4O_OK;
}

/* Inflates one or more concatenated frames.  dst may be NULL with cap = 0 to
 * fail with LZ4O_OVERFLOW as soon as a byte is produced (not a sizing API). */
int sbmc_oracle_lz4_frame_decompress(const uint8_t *src, int64_t n, uint8_t *dst, int64_t cap,
                                     int64_t *out_len) {
  const uint8_t *ip = src, *end = src + n;
  int64_t op = 0;
  int frames = 0;
  *out_len = 0;
  while (ip < end) {
    if (end - ip < 4) return LZ4O_TRUNCATED;
    uint32_t magic = rd32(ip);
    ip += 4;
    if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {
      if (end - ip < 4) return LZ4O_TRUNCATED;
      int64_t skip = rd32(ip);
      ip += 4;
      if (skip > end - ip) return LZ4O_TRUNCATED;
      ip += skip;
      continue;
    }
    if (magic != 0x184D2204u) return frames ? LZ4O_OK : LZ4O_BAD_MAGIC;
    const uint8_t *desc = ip;
    if (end - ip < 3) return LZ4O_TRUNCATED;
    unsigned flg = ip[0], bd = ip[1];
    ip += 2;
    if ((flg >> 6) != 1 || (flg & 2) || (bd & 0x8F)) return LZ4O_BAD_HEADER;
    int indep = flg & 0x20, bsum = flg & 0x10, csize = flg & 8, csum = flg & 4, dict = flg & 1;
    int id = (bd >> 4) & 7;
    if (id < 4) return LZ4O_BAD_HEADER;
    int64_t bmax = (int64_t)1 << (8 + 2 * id);
    int64_t opt = (csize ? 8 : 0) + (dict ? 4 : 0);
    if (opt + 1 > end - ip) return LZ4O_TRUNCATED;
    ip += opt;
    if (((sbmc_oracle_xxh32(desc, ip - desc, 0) >> 8) & 0xFF) != *ip) return LZ4O_BAD_CHECKSUM;
    ++ip;
    int64_t frame_start = op;
    while (1) {
      if (end - ip < 4) return LZ4O_TRUNCATED;
      uint32_t word = rd32(ip);
      ip += 4;
      if (word == 0) break;
      int64_t bs = word & 0x7FFFFFFFu;
      if (bs > bmax) return LZ4O_BLOCK_TOO_LARGE;
      if (bs > end - ip) return LZ4O_TRUNCATED;
      if (word & 0x80000000u) {
        if (bs > cap - op) return LZ4O_OVERFLOW;
        memcpy(dst + op, ip, (size_t)bs);
        op += bs;
      } else {
        int rc = block(ip, ip + bs, dst, &op, cap, indep ? op : frame_start);
        if (rc) return rc;
      }
      if (bsum) {
        if (end - ip - bs < 4) return LZ4O_TRUNCATED;
        if (rd32(ip + bs) != sbmc_oracle_xxh32(ip, bs, 0)) return LZ4O_BAD_CHECKSUM;
        ip += 4;
      }
      ip += bs;
    }
    if (csum) {
      if (end - ip < 4) return LZ4O_TRUNCATED;
      if (rd32(ip) != sbmc_oracle_xxh32(dst + frame_start, op - frame_start, 0))
        return LZ4O_BAD_CHECKSUM;
      ip += 4;
    }
    ++frames;
    *out_len = op;
  }
  return frames ? LZ4O_OK : LZ4O_BAD_MAGIC;
}
