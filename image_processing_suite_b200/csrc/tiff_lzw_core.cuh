// TIFF LZW strip coder, shared between the CUDA kernels of tiff.cu and the host harness of
// tests/lzw_host_harness.cpp (which compiles this header with g++, a 1-lane "warp", to check
// the state machines against Pillow/libtiff on the CPU-only authoring box).
//
// The stream format is TIFF 6.0 section 13 as libtiff writes it (the codec behind
// img.save(..., compression='tiff_lzw'), Image_re-binning.py:19-21): MSB-first codes of 9..12
// bits with the "early change", ClearCode first in every strip, table reset once code 4093 has
// been assigned, and a compression-ratio check every 10000 input bytes that may reset early.
// Following that policy to the letter makes the strips byte-identical to libtiff's, so the
// files can be compared with cmp, not only decoded.
//
// Execution model: one warp per strip.  LZW is a serial state machine, so every lane runs the
// same instruction stream on the same values (shared-memory reads broadcast, identical writes
// collapse) and only lane 0 stores to global memory; the lanes split the work that does
// parallelise -- clearing the 24 KB hash table and flushing decoded bytes in 16-byte vectors.
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define LZW_HD __device__ __forceinline__
#else
#define LZW_HD inline
#endif

namespace ips_lzw {

enum : int { BITS_MIN = 9, BITS_MAX = 12, CODE_CLEAR = 256, CODE_EOI = 257, CODE_FIRST = 258, CODE_MAX = 4095 };
enum : uint32_t { CHECK_GAP = 10000 };
enum : uint32_t { ENC_SLOTS = 6144, ENC_EMPTY = 0xFFFFFFFFu };      // open addressing, load <= 0.63
enum : uint32_t { DEC_CODES = 4096, DEC_OBUF = 4096 };
enum : uint32_t { OVERFLOW = 0xFFFFFFFFu };
enum : int { ST_OK = 0, ST_TRUNCATED = 1, ST_CORRUPT = 2, ST_OLD_STYLE = 3 };

#ifdef __CUDACC__
struct Warp {
  int lane;
  __device__ __forceinline__ Warp() : lane(threadIdx.x & 31) {}
  static constexpr int n = 32;
  __device__ __forceinline__ void sync() const { __syncwarp(); }
};
LZW_HD uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }
LZW_HD uint32_t load_u32(const uint8_t* p) { return *reinterpret_cast<const uint32_t*>(p); }
struct alignas(16) Vec16 { uint32_t a, b, c, d; };
LZW_HD void copy16(uint8_t* dst, const uint8_t* src) {
  *reinterpret_cast<Vec16*>(dst) = *reinterpret_cast<const Vec16*>(src);
}
#else
struct Warp {
  int lane = 0;
  static constexpr int n = 1;
  void sync() const {}
};
inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }
inline uint32_t load_u32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline void copy16(uint8_t* dst, const uint8_t* src) { memcpy(dst, src, 16); }
#endif

// Worst case of a strip of n bytes: every byte becomes a 12-bit code, plus a ClearCode per
// table generation, the first ClearCode, EOI and the word-granular flush.
static inline size_t encode_bound(size_t n) { return (n + n / 2 + n / 1024 + 64 + 15) / 16 * 16; }

// ------------------------------------------------------------------------------------------
// encoder
// ------------------------------------------------------------------------------------------
// One byte costs one dependent chain  key -> hash -> shared-memory probe -> compare  (the warp
// has nothing else to overlap it with), so the bookkeeping libtiff does per byte is folded into
// quantities that only change on a table miss: incount = consumed - in_base, outcount =
// bits emitted - bits_base, and the "code width grows / table full" tests share one compare.
#ifdef __CUDACC__
LZW_HD uint32_t slot_of(uint32_t key) { return __umulhi(key * 0x9E3779B1u, (uint32_t)ENC_SLOTS); }
#else
inline uint32_t slot_of(uint32_t key) { return (uint32_t)(((uint64_t)(key * 0x9E3779B1u) * ENC_SLOTS) >> 32); }
#endif

struct Encoder {
  uint32_t* tab;     // [ENC_SLOTS] (key << 12) | code, key = (prefix code << 8) | byte
  uint8_t* out;      // 4-byte aligned
  uint32_t cap;      // multiple of 4
  uint32_t op;       // bytes written (multiple of 4 until finish())
  uint32_t acc;      // low nacc (< 32) bits are pending output; bits above them are stale
  int nacc;
  int nbits;
  uint32_t limit;    // free_ent at which the next width change (512, 1024, 2048) or the reset (4094) happens
  uint32_t free_ent, ent;
  uint32_t in_base, bits_base, checkpoint, ratio;
  bool overflow;
  int lane;

  LZW_HD void put(uint32_t code) {
    const int t = nacc + nbits;
    if (t >= 32) {
      const int r = t - 32;                                    // low bits of `code` that stay pending
      const uint32_t word = (acc << (nbits - r)) | (code >> r);
      acc = code;
      nacc = r;
      if (op + 4 <= cap) {
        if (lane == 0) *reinterpret_cast<uint32_t*>(out + op) = bswap32(word);
        op += 4;
      } else {
        overflow = true;
      }
    } else {
      acc = (acc << nbits) | code;
      nacc = t;
    }
  }

  template <class W>
  LZW_HD void clear_table(const W& w) {
    w.sync();
    uint64_t* t64 = reinterpret_cast<uint64_t*>(tab);
#pragma unroll 4
    for (uint32_t i = w.lane; i < ENC_SLOTS / 2; i += W::n) t64[i] = ~0ull;
    w.sync();
  }

  // libtiff: cl_hash, ratio = incount = outcount = 0, CODE_CLEAR at the old width, 9-bit codes
  template <class W>
  LZW_HD void reset(uint32_t consumed, const W& w) {
    clear_table(w);
    ratio = 0;
    in_base = consumed;
    bits_base = op * 8u + (uint32_t)nacc;
    free_ent = CODE_FIRST;
    put(CODE_CLEAR);
    nbits = BITS_MIN;
    limit = 1u << BITS_MIN;
  }

  template <class W>
  LZW_HD void begin(uint32_t* table, uint8_t* dst, uint32_t capacity, const W& w) {
    tab = table;
    out = dst;
    cap = capacity & ~3u;
    op = 0;
    acc = 0;
    nacc = 0;
    nbits = BITS_MIN;
    limit = 1u << BITS_MIN;
    free_ent = CODE_FIRST;
    ent = 0;
    in_base = 0;
    bits_base = 0;
    checkpoint = CHECK_GAP;
    ratio = 0;
    overflow = false;
    lane = w.lane;
    clear_table(w);
  }

  // c is the consumed-th byte of the strip (1-based), not the first
  template <class W>
  LZW_HD void byte(uint32_t c, uint32_t consumed, const W& w) {
    const uint32_t key = (ent << 8) | c;
    uint32_t h = slot_of(key);
    for (;;) {
      const uint32_t s = tab[h];
      if ((s >> 12) == key) {
        ent = s & 0xFFFu;
        return;
      }
      if (s == ENC_EMPTY) break;
      h = h + 1 == ENC_SLOTS ? 0u : h + 1;
    }
    put(ent);
    ent = c;
    tab[h] = (key << 12) | free_ent;
    free_ent++;
    if (free_ent == limit) {
      if (limit == (uint32_t)CODE_MAX - 1) {
        reset(consumed, w);
      } else {
        nbits++;
        limit = limit == 2048u ? (uint32_t)CODE_MAX - 1 : limit << 1;
      }
    } else if (consumed - in_base >= checkpoint) {
      const uint32_t incount = consumed - in_base;
      const uint32_t outcount = op * 8u + (uint32_t)nacc - bits_base;
      checkpoint = incount + CHECK_GAP;
      uint32_t rat;
      if (incount > 0x007fffffu) {
        rat = outcount >> 8;
        rat = rat == 0 ? 0x7fffffffu : incount / rat;
      } else {
        rat = (incount << 8) / outcount;
      }
      if (rat <= ratio)
        reset(consumed, w);
      else
        ratio = rat;
    }
  }

  // LZWPostEncode; returns the strip's byte count or OVERFLOW
  LZW_HD uint32_t finish(bool any) {
    if (any) {
      put(ent);
      free_ent++;
      if (free_ent == (uint32_t)CODE_MAX - 1) {
        put(CODE_CLEAR);
        nbits = BITS_MIN;
      } else if (free_ent == limit) {
        nbits++;
      }
    }
    put(CODE_EOI);
    while (nacc > 0) {
      const uint32_t b = nacc >= 8 ? (acc >> (nacc - 8)) & 0xFFu : (acc << (8 - nacc)) & 0xFFu;
      nacc -= 8;
      if (op < cap) {
        if (lane == 0) out[op] = (uint8_t)b;
        op++;
      } else {
        overflow = true;
      }
    }
    return overflow ? (uint32_t)OVERFLOW : op;
  }
};

// in: n < 2^28 bytes (any alignment); out: 4-byte aligned, cap bytes; table: ENC_SLOTS words
template <class W>
LZW_HD uint32_t encode_strip(const uint8_t* in, uint32_t n, uint8_t* out, uint32_t cap, uint32_t* table, const W& w) {
  Encoder e;
  e.begin(table, out, cap, w);
  if (n == 0) return e.finish(false);
  e.put(CODE_CLEAR);
  e.ent = in[0];
  uint32_t next = n > 1 ? in[1] : 0u;          // one byte of look-ahead keeps the load off the chain
#pragma unroll 1
  for (uint32_t i = 1; i < n; ++i) {
    const uint32_t c = next;
    if (i + 1 < n) next = in[i + 1];
    e.byte(c, i + 1, w);
  }
  return e.finish(true);
}

// ------------------------------------------------------------------------------------------
// decoder
// ------------------------------------------------------------------------------------------
// tab[k] = prefix << 20 | length << 8 | last byte, for k >= 258; firstc[k] = first byte of the string
struct Decoder {
  const uint8_t* in;
  uint32_t n_in, ip;
  uint64_t acc;      // top nb bits valid
  int nb;
  uint32_t nextw;    // next 32 input bits, loaded one refill ahead

  LZW_HD uint32_t fetch_word() {      // 32 bits at ip (4-byte aligned address), zero-padded past the end
    uint32_t v;
    if (ip + 4 <= n_in) {
      v = bswap32(load_u32(in + ip));
    } else {
      v = 0;
      for (uint32_t k = 0; k < 4; ++k)
        if (ip + k < n_in) v |= (uint32_t)in[ip + k] << (24 - 8 * k);
    }
    ip += 4;
    return v;
  }
  LZW_HD void begin(const uint8_t* src, uint32_t n) {
    in = src;
    n_in = n;
    ip = 0;
    acc = 0;
    nb = 0;
    while (ip < n_in && ((reinterpret_cast<uintptr_t>(in + ip)) & 3u)) {
      acc |= (uint64_t)in[ip++] << (56 - nb);
      nb += 8;
    }
    nextw = fetch_word();
  }
  LZW_HD void refill() {
    if (nb <= 32) {
      acc |= (uint64_t)nextw << (32 - nb);
      nb += 32;
      nextw = fetch_word();
    }
  }
};

// Decodes one strip to out[0..n_out).  obuf: DEC_OBUF bytes of 16-byte aligned shared memory;
// tab: DEC_CODES words; firstc: DEC_CODES bytes.  Returns a status; on any failure the
// undecoded remainder is zero-filled so the output is deterministic.
template <class W>
LZW_HD int decode_strip(const uint8_t* in, uint32_t n_in, uint8_t* out, uint32_t n_out, uint32_t* tab, uint8_t* firstc,
                        uint8_t* obuf, const W& w) {
  int status = ST_OK;
  if (n_in >= 2 && in[0] == 0 && (in[1] & 1)) status = ST_OLD_STYLE;   // pre-6.0 LSB-first streams
  Decoder d;
  d.begin(in, n_in);
  // bits that really exist (n_in < 2^28), to tell padding zeros from data
  int32_t avail = (int32_t)(n_in * 8u);
  uint32_t gpos = 0;                                              // bytes of the strip already in global memory
  uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(out) & 15u);   // obuf index of byte gpos
  uint32_t q = 0;                                                 // bytes buffered
  uint32_t pos = 0;                                               // gpos + q
  int nbits = BITS_MIN, free_ent = CODE_FIRST;
  int old = -1, old_len = 0;
  uint32_t old_first = 0;

  auto flush = [&](bool final) {
    w.sync();
    uint8_t* g = out + gpos;                    // global address of obuf[shift]
    const uint32_t n = q;
    uint32_t head = (16u - shift) & 15u;
    if (head > n) head = n;
    uint32_t body = (n - head) & ~15u;
    uint32_t tail = n - head - body;
    for (uint32_t i = w.lane; i < head; i += W::n) g[i] = obuf[shift + i];
    for (uint32_t i = w.lane * 16u; i < body; i += W::n * 16u) copy16(g + head + i, obuf + shift + head + i);
    if (final) {
      for (uint32_t i = w.lane; i < tail; i += W::n) g[head + body + i] = obuf[shift + head + body + i];
      gpos += n;
      q = 0;
      shift = (shift + n) & 15u;
    } else {
      w.sync();
      const uint32_t src = shift + head + body;       // a multiple of 16 whenever tail > 0
      if (tail && w.lane == 0)
        for (uint32_t i = 0; i < tail; ++i) obuf[i] = obuf[src + i];
      gpos += head + body;
      q = tail;
      shift = tail ? 0u : ((shift + head + body) & 15u);
    }
    w.sync();
  };

  while (status == ST_OK && pos < n_out) {
    d.refill();
    if (avail < nbits) { status = ST_TRUNCATED; break; }
    const int code = (int)(d.acc >> (64 - nbits));
    d.acc <<= nbits;
    d.nb -= nbits;
    avail -= nbits;
    if (code == CODE_EOI) { status = ST_TRUNCATED; break; }
    if (code == CODE_CLEAR) {
      free_ent = CODE_FIRST;
      nbits = BITS_MIN;
      old = -1;
      continue;
    }
    if (old < 0) {                         // first code after a clear is a literal
      if (code >= 256) { status = ST_CORRUPT; break; }
      if (shift + q + 1 > DEC_OBUF) flush(false);
      obuf[shift + q] = (uint8_t)code;            // every lane stores the same byte
      q++;
      pos++;
      old = code;
      old_len = 1;
      old_first = (uint32_t)code;
      continue;
    }
    if (code > free_ent || (code == free_ent && free_ent >= (int)DEC_CODES)) { status = ST_CORRUPT; break; }
    uint32_t first;
    if (code < 256) first = (uint32_t)code;
    else if (code == free_ent) first = old_first;
    else first = firstc[code];
    if (free_ent < (int)DEC_CODES) {       // entry = previous string + first byte of this one
      tab[free_ent] = ((uint32_t)old << 20) | ((uint32_t)(old_len + 1) << 8) | first;
      firstc[free_ent] = (uint8_t)old_first;
    }
    uint32_t len = code < 256 ? 1u : ((tab[code] >> 8) & 0xFFFu);
    const uint32_t room = n_out - pos;
    const uint32_t keep = len < room ? len : room;
    if (shift + q + keep > DEC_OBUF) flush(false);
    {
      int k = code;
      uint32_t i = len;
      uint8_t* dst = obuf + shift + q;
      while (i > 1) {
        const uint32_t e = tab[k];
        --i;
        if (i < keep) dst[i] = (uint8_t)(e & 0xFFu);
        k = (int)(e >> 20);
      }
      dst[0] = (uint8_t)k;
    }
    q += keep;
    pos += keep;
    old = code;
    old_len = (int)len;
    old_first = first;
    if (free_ent < (int)DEC_CODES) {
      free_ent++;
      if (free_ent >= (1 << nbits) - 1 && nbits < BITS_MAX) nbits++;
    }
  }
  flush(true);
  if (pos < n_out) {
    if (status == ST_OK) status = ST_TRUNCATED;
    for (uint32_t i = pos + w.lane; i < n_out; i += W::n) out[i] = 0;
  }
  return status;
}

}  // namespace ips_lzw
