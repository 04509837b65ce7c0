// lz4_warp.cuh -- LZ4 *frame* inflater, one warp per frame.
//
// The reference stores every tile as a header plus (1 + sample_count) LZ4
// frames (writer: LZ4F_compressFrame with default preferences,
// pbrt_patches/sbmc_pbrt.diff:6140-6158; reader: lz4.frame.decompress,
// sbmc/datasets.py:570-579).  lz4 itself is a third-party dependency that is
// not vendored (setup.py:104 `lz4`, liblz4-dev in dockerfiles/*.dockerfile);
// this file restates its published formats:
//   frame  = magic 0x184D2204, FLG, BD, [content size 8B], [dict id 4B], HC,
//            blocks { u32 size (bit 31 = stored raw), data, [u32 checksum] },
//            end mark 0, [u32 content checksum]
//   block  = sequences { token, literal-length bytes, literals,
//            u16 offset, match-length bytes }, the last sequence stops after
//            its literals.  Blocks of one frame may reference up to 64 KiB of
//            earlier output of the same frame ("linked" blocks, the default).
//
// Parallel form: the control state (read / write cursors, lengths) is
// warp-uniform -- every lane parses the same token bytes (one broadcast load)
// -- and only the two copy loops are split over the lanes.  A match that
// overlaps its own output (offset < length) is periodic with period `offset`,
// so lane i reads dst[op - offset + i % offset]: every read lies below `op`,
// i.e. in bytes finished before this sequence, and no intra-sequence ordering
// is needed.  One __syncwarp() per sequence publishes the bytes to the lanes
// that read them next.
//
// The same source compiles for the host (tests/native/tiles_emul.cpp) with the lane
// loops run sequentially, so the cursor arithmetic is checked on CPU against
// the reference library's own output.  Header, block and content checksums
// (xxHash32) are verified like `lz4.frame.decompress` does.
#pragma once
#include <stdint.h>

#if defined(__CUDA_ARCH__)
#define SBMC_LZ4_FN __device__ __forceinline__
#define SBMC_LZ4_LANES(lane) for (int lane = (int)(threadIdx.x & 31u), once__ = 1; once__; once__ = 0)
#define SBMC_LZ4_PUBLISH() __syncwarp()
#else
#define SBMC_LZ4_FN static inline
#if defined(SBMC_LZ4_REVERSE_LANES)  // host tests: lane order must not matter
#define SBMC_LZ4_LANES(lane) for (int lane = 31; lane >= 0; --lane)
#else
#define SBMC_LZ4_LANES(lane) for (int lane = 0; lane < 32; ++lane)
#endif
#define SBMC_LZ4_PUBLISH() ((void)0)
#endif

namespace sbmc {
namespace lz4 {

enum Status : int {
  kOk = 0,
  kBadMagic = 1,      // not an LZ4 frame
  kBadHeader = 2,     // unsupported version / reserved bits / block size id
  kTruncated = 3,     // input ends inside a header, block or sequence
  kOverflow = 4,      // output would exceed the caller's capacity
  kBadOffset = 5,     // match offset 0 or beyond the produced output
  kSizeMismatch = 6,  // decoded size differs from the size the tile header implies
  kBlockTooLarge = 7, // block larger than the frame's declared maximum
  kBadChecksum = 8,   // header / block / content xxHash32 mismatch
};

SBMC_LZ4_FN uint32_t load_u32(const uint8_t *p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// xxHash32 (seed 0) of p[0..len): the checksum of the frame format.  Computed
// by every lane alike (uniform loads broadcast), so the result is warp-uniform
// without a shuffle; only frames that carry checksums pay for it (the
// reference's writer emits none besides the 1-byte header check).
SBMC_LZ4_FN uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

SBMC_LZ4_FN uint32_t xxh32(const uint8_t *p, int64_t len) {
  const uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u,
                 P5 = 374761393u;
  const uint8_t *const end = p + len;
  uint32_t h;
  if (len >= 16) {
    uint32_t v1 = P1 + P2, v2 = P2, v3 = 0, v4 = 0u - P1;
    const uint8_t *const limit = end - 16;
    do {
      v1 = rotl32(v1 + load_u32(p) * P2, 13) * P1;
      v2 = rotl32(v2 + load_u32(p + 4) * P2, 13) * P1;
      v3 = rotl32(v3 + load_u32(p + 8) * P2, 13) * P1;
      v4 = rotl32(v4 + load_u32(p + 12) * P2, 13) * P1;
      p += 16;
    } while (p <= limit);
    h = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
  } else {
    h = P5;
  }
  h += (uint32_t)len;
  while (p + 4 <= end) {
    h = rotl32(h + load_u32(p) * P3, 17) * P4;
    p += 4;
  }
  while (p < end) {
    h = rotl32(h + (uint32_t)(*p) * P5, 11) * P1;
    ++p;
  }
  h ^= h >> 15;
  h *= P2;
  h ^= h >> 13;
  h *= P3;
  h ^= h >> 16;
  return h;
}

// Copies n bytes src -> dst, lanes interleaved byte-wise (32 consecutive bytes
// per warp access = one sector), four accesses in flight per lane.
SBMC_LZ4_FN void copy_bytes(uint8_t *dst, const uint8_t *src, int64_t n) {
  SBMC_LZ4_LANES(lane) {
    int64_t i = lane;
    for (; i + 96 < n; i += 128) {
      uint8_t a = src[i], b = src[i + 32], c = src[i + 64], d = src[i + 96];
      dst[i] = a;
      dst[i + 32] = b;
      dst[i + 64] = c;
      dst[i + 96] = d;
    }
    for (; i < n; i += 32) dst[i] = src[i];
  }
}

#if defined(SBMC_LZ4_WIDE_COPY)
// Experimental (build with -DSBMC_LZ4_WIDE_COPY, off by default): long runs move
// as 16-byte stores.  The destination is brought to 16-byte alignment byte-wise;
// the source keeps an arbitrary alignment, so each 16-byte chunk is assembled
// from five 4-byte-aligned words with a funnel shift.  The word window of a
// chunk starts at most 3 bytes before its first source byte and ends at or
// before src + n (the last chunks go byte-wise), so nothing outside
// [src - 3, src + n) is read -- for a match (src = dst - offset, offset >= n)
// that stays below dst.
SBMC_LZ4_FN uint32_t load_aligned_u32(const uint8_t *p) {
#if defined(__CUDA_ARCH__)
  return *reinterpret_cast<const uint32_t *>(p);
#else
  uint32_t v;
  __builtin_memcpy(&v, p, 4);
  return v;
#endif
}

SBMC_LZ4_FN uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t shift) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, shift);
#else
  return shift ? (uint32_t)((((uint64_t)hi << 32) | lo) >> shift) : lo;
#endif
}

SBMC_LZ4_FN void copy_lanes(uint8_t *dst, const uint8_t *src, int64_t n) {
  if (n < 256) {
    copy_bytes(dst, src, n);
    return;
  }
  const int64_t head = (int64_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15);
  const int64_t chunks = (n - head - 3) / 16;     // chunks whose word window ends inside src
  copy_bytes(dst, src, head);
  SBMC_LZ4_LANES(lane) {
    for (int64_t c = lane; c < chunks; c += 32) {
      const uint8_t *s = src + head + 16 * c;
      const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(s) & 3) * 8;
      const uint8_t *a = s - (shift >> 3);
      const uint32_t w0 = load_aligned_u32(a), w1 = load_aligned_u32(a + 4),
                     w2 = load_aligned_u32(a + 8), w3 = load_aligned_u32(a + 12);
      const uint32_t w4 = shift ? load_aligned_u32(a + 16) : 0u;
      uint32_t out[4] = {funnel_r(w0, w1, shift), funnel_r(w1, w2, shift),
                         funnel_r(w2, w3, shift), funnel_r(w3, w4, shift)};
#if defined(__CUDA_ARCH__)
      *reinterpret_cast<uint4 *>(dst + head + 16 * c) = make_uint4(out[0], out[1], out[2], out[3]);
#else
      __builtin_memcpy(dst + head + 16 * c, out, 16);
#endif
    }
  }
  const int64_t done = head + 16 * chunks;
  copy_bytes(dst + done, src + done, n - done);
}
#else
SBMC_LZ4_FN void copy_lanes(uint8_t *dst, const uint8_t *src, int64_t n) { copy_bytes(dst, src, n); }
#endif

// dst[0..n) = the `offset` bytes before dst, repeated.  Reads stay below dst.
SBMC_LZ4_FN void match_lanes(uint8_t *dst, int64_t offset, int64_t n) {
  const uint8_t *from = dst - offset;
  if (offset >= n) {
    copy_lanes(dst, from, n);
    return;
  }
  SBMC_LZ4_LANES(lane) {
    // i % offset without a division per byte: advance a running remainder
    // (offsets are 16-bit, so 32-bit arithmetic).
    const uint32_t period = (uint32_t)offset;
    uint32_t r = (uint32_t)lane % period;
    const uint32_t step = 32u % period;
    for (int64_t i = lane; i < n; i += 32) {
      dst[i] = from[r];
      r += step;
      if (r >= period) r -= period;
    }
  }
}

// One LZ4 block [ip, ip_end) appended at dst + op; `window` = first output byte
// a match may reference.  Returns a Status; *op_io advances by the block's size.
//
// A frame is a serial chain of sequences and ONE warp walks it, so the cost of a frame is
// (instructions per sequence) x (sequences): ncu of the first version showed the kernel
// issue-bound per warp, not memory-bound (profiles/r3p_lz4_ncu.md).  Short sequences --
// at most 32 literals and a match of at most 32 bytes, the common case on sample data,
// where most matches are one repeated float -- therefore take a loop-free path: one
// predicated byte per lane for the literals, one for the match, 32-bit lengths, cursors as
// pointers.  One __syncwarp() per sequence, in front of the match (it publishes the
// previous match and this sequence's literals to the lanes that may read them).  Sequences
// without extension bytes run in a loop whose truncation / overflow checks are hoisted.
SBMC_LZ4_FN int decode_block(const uint8_t *ip, const uint8_t *ip_end, uint8_t *dst,
                             int64_t *op_io, int64_t dst_cap, int64_t window) {
  uint8_t *out = dst + *op_io;
  uint8_t *const out_end = dst + dst_cap;
  const uint8_t *const win = dst + window;
  // bytes a match may reach back from `out`, saturated (offsets are at most 65535)
  uint32_t have = (out - win) > 0x100000 ? 0x100000u : (uint32_t)(out - win);
  for (;;) {
    // Fast loop: sequences without length-extension bytes (at most 14 literals, a match of
    // at most 18 bytes), while 32 input bytes and 32 output bytes are left -- no truncation
    // / overflow checks per sequence, and such a sequence cannot be the block's last one.
    while (ip_end - ip >= 32 && out_end - out >= 32) {
      const uint32_t token = *ip;
      const uint32_t lit = token >> 4, mcode = token & 15;
      if (lit == 15 || mcode == 15) break;
      SBMC_LZ4_LANES(lane) {
        if ((uint32_t)lane < lit) out[lane] = ip[1 + lane];
      }
      const uint8_t *q = ip + 1 + lit;
      const uint32_t offset = (uint32_t)q[0] | ((uint32_t)q[1] << 8);
      const uint32_t mlen = mcode + 4;
      out += lit;
      have += lit;
      if (offset == 0 || offset > have) return kBadOffset;
      SBMC_LZ4_PUBLISH();  // earlier output (incl. the literals just written) may be the source
      const uint8_t *from = out - offset;
      if (offset >= mlen) {
        SBMC_LZ4_LANES(lane) {
          if ((uint32_t)lane < mlen) out[lane] = from[lane];
        }
      } else {  // overlapping: periodic with period `offset`, every read lies below `out`
        SBMC_LZ4_LANES(lane) {
          if ((uint32_t)lane < mlen) out[lane] = from[(uint32_t)lane % offset];
        }
      }
      out += mlen;
      have = have > 0x100000 ? have : have + mlen;
      ip = q + 2;
    }
    // General form: one sequence with every check.
    if (ip >= ip_end) return kTruncated;
    const uint32_t token = *ip++;
    uint32_t lit = token >> 4;
    if (lit == 15) {
      uint32_t b;
      do {
        if (ip >= ip_end) return kTruncated;
        b = *ip++;
        lit += b;
      } while (b == 255);
    }
    if ((int64_t)lit > ip_end - ip) return kTruncated;
    if ((int64_t)lit > out_end - out) return kOverflow;
    if (lit <= 32) {
      SBMC_LZ4_LANES(lane) {
        if ((uint32_t)lane < lit) out[lane] = ip[lane];
      }
    } else {
      copy_lanes(out, ip, lit);
    }
    ip += lit;
    out += lit;
    have = have > 0x100000 ? have : have + (lit > 0x100000 ? 0x100000u : lit);
    if (ip == ip_end) break;  // the last sequence carries literals only
    if (ip_end - ip < 2) return kTruncated;
    const uint32_t offset = (uint32_t)ip[0] | ((uint32_t)ip[1] << 8);
    ip += 2;
    uint32_t mlen = token & 15;
    if (mlen == 15) {
      uint32_t b;
      do {
        if (ip >= ip_end) return kTruncated;
        b = *ip++;
        mlen += b;
      } while (b == 255);
    }
    mlen += 4;
    if (offset == 0 || offset > have) return kBadOffset;
    if ((int64_t)mlen > out_end - out) return kOverflow;
    SBMC_LZ4_PUBLISH();
    if (mlen <= 32) {
      const uint8_t *from = out - offset;
      if (offset >= mlen) {
        SBMC_LZ4_LANES(lane) {
          if ((uint32_t)lane < mlen) out[lane] = from[lane];
        }
      } else {
        SBMC_LZ4_LANES(lane) {
          if ((uint32_t)lane < mlen) out[lane] = from[(uint32_t)lane % offset];
        }
      }
    } else {
      match_lanes(out, offset, mlen);
    }
    out += mlen;
    have = have > 0x100000 ? have : have + (mlen > 0x100000 ? 0x100000u : mlen);
  }
  SBMC_LZ4_PUBLISH();
  *op_io = out - dst;
  return kOk;
}

// Inflates src[0..src_len) (one or more concatenated frames, skippable frames
// ignored) into dst[0..dst_cap).  *out_len = bytes produced.  Warp-uniform.
SBMC_LZ4_FN int decode_frames(const uint8_t *src, int64_t src_len, uint8_t *dst, int64_t dst_cap,
                              int64_t *out_len) {
  const uint8_t *ip = src;
  const uint8_t *const end = src + src_len;
  int64_t op = 0;
  int frames = 0;
  *out_len = 0;
  while (ip < end) {
    if (end - ip < 4) return kTruncated;
    const uint32_t magic = load_u32(ip);
    ip += 4;
    if ((magic & 0xFFFFFFF0u) == 0x184D2A50u) {  // skippable frame
      if (end - ip < 4) return kTruncated;
      const int64_t skip = load_u32(ip);
      ip += 4;
      if (skip > end - ip) return kTruncated;
      ip += skip;
      continue;
    }
    if (magic != 0x184D2204u) return frames ? kOk : kBadMagic;  // trailing bytes: ignored
    if (end - ip < 3) return kTruncated;
    const uint32_t flg = ip[0], bd = ip[1];
    ip += 2;
    if ((flg >> 6) != 1 || (flg & 2) || (bd & 0x8F)) return kBadHeader;
    const bool independent = flg & 0x20, block_sum = flg & 0x10, has_size = flg & 8,
               content_sum = flg & 4, has_dict = flg & 1;
    const int size_id = (bd >> 4) & 7;
    if (size_id < 4) return kBadHeader;
    const int64_t block_max = (int64_t)1 << (8 + 2 * size_id);
    const int64_t skip = (has_size ? 8 : 0) + (has_dict ? 4 : 0);
    if (skip + 1 > end - ip) return kTruncated;
    ip += skip;
    if (((xxh32(ip - skip - 2, skip + 2) >> 8) & 0xFFu) != *ip) return kBadChecksum;
    ++ip;  // header checksum byte
    const int64_t frame_start = op;
    for (;;) {
      if (end - ip < 4) return kTruncated;
      const uint32_t word = load_u32(ip);
      ip += 4;
      if (word == 0) break;  // end mark
      const int64_t bsize = word & 0x7FFFFFFFu;
      if (bsize > block_max) return kBlockTooLarge;
      if (bsize > end - ip) return kTruncated;
      if (word & 0x80000000u) {  // stored block
        if (bsize > dst_cap - op) return kOverflow;
        copy_lanes(dst + op, ip, bsize);
        op += bsize;
        SBMC_LZ4_PUBLISH();
      } else {
        const int rc = decode_block(ip, ip + bsize, dst, &op, dst_cap, independent ? op : frame_start);
        if (rc != kOk) return rc;
      }
      if (block_sum) {  // xxHash32 of the block as stored
        if (end - ip - bsize < 4) return kTruncated;
        if (load_u32(ip + bsize) != xxh32(ip, bsize)) return kBadChecksum;
        ip += 4;
      }
      ip += bsize;
    }
    if (content_sum) {  // xxHash32 of the frame's inflated bytes
      if (end - ip < 4) return kTruncated;
      if (load_u32(ip) != xxh32(dst + frame_start, op - frame_start)) return kBadChecksum;
      ip += 4;
    }
    ++frames;
    *out_len = op;
  }
  *out_len = op;
  return frames ? kOk : kBadMagic;
}

}  // namespace lz4
}  // namespace sbmc
