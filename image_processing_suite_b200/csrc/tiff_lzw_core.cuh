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
// parallelise -- clearing the 32 KB hash table and flushing decoded bytes in 16-byte vectors.
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
enum : uint32_t { ENC_SLOTS = 8192, ENC_EMPTY = 0xFFFFFFFFu };      // open addressing, load <= 0.47
enum : uint32_t { DEC_CODES = 4096, DEC_OBUF = 8192 };
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
struct Encoder {
  uint32_t* tab;     // [ENC_SLOTS] (key << 12) | code, key = (prefix code << 8) | byte
  uint8_t* out;      // 4-byte aligned
  uint32_t cap;      // multiple of 4
  uint32_t op;       // bytes written (multiple of 4 until finish())
  uint64_t acc;      // low nacc bits are pending output
  int nacc;
  int nbits, maxcode, free_ent, ent;
  uint64_t checkpoint, ratio, incount, outcount;
  bool overflow;
  int lane;

  LZW_HD void put(int code) {
    acc = (acc << nbits) | (uint32_t)code;
    nacc += nbits;
    outcount += (uint64_t)nbits;
    if (nacc >= 32) {
      const uint32_t word = (uint32_t)(acc >> (nacc - 32));
      nacc -= 32;
      if (op + 4 <= cap) {
        if (lane == 0) *reinterpret_cast<uint32_t*>(out + op) = bswap32(word);
        op += 4;
      } else {
        overflow = true;
      }
    }
  }

  template <class W>
  LZW_HD void clear_table(const W& w) {
    w.sync();
    uint64_t* t64 = reinterpret_cast<uint64_t*>(tab);
    for (uint32_t i = w.lane; i < ENC_SLOTS / 2; i += W::n) t64[i] = ~0ull;
    w.sync();
  }

  template <class W>
  LZW_HD void reset_after_clear(const W& w) {   // libtiff: cl_hash + CODE_CLEAR + 9-bit codes
    clear_table(w);
    ratio = 0;
    incount = 0;
    outcount = 0;
    free_ent = CODE_FIRST;
    put(CODE_CLEAR);
    nbits = BITS_MIN;
    maxcode = (1 << BITS_MIN) - 1;
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
    maxcode = (1 << BITS_MIN) - 1;
    free_ent = CODE_FIRST;
    ent = -1;
    checkpoint = CHECK_GAP;
    ratio = 0;
    incount = 0;
    outcount = 0;
    overflow = false;
    lane = w.lane;
    clear_table(w);
  }

  LZW_HD void first(uint32_t c) {      // first byte of the strip
    put(CODE_CLEAR);
    ent = (int)c;
    incount++;
  }

  template <class W>
  LZW_HD void byte(uint32_t c, const W& w) {
    incount++;
    const uint32_t key = ((uint32_t)ent << 8) | c;
    uint32_t h = (key * 0x9E3779B1u) >> 19;
    for (;;) {
      const uint32_t s = tab[h];
      if ((s >> 12) == key) {
        ent = (int)(s & 0xFFFu);
        return;
      }
      if (s == ENC_EMPTY) break;
      h = (h + 1) & (ENC_SLOTS - 1);
    }
    put(ent);
    ent = (int)c;
    tab[h] = (key << 12) | (uint32_t)free_ent;
    free_ent++;
    if (free_ent == CODE_MAX - 1) {
      reset_after_clear(w);
    } else if (free_ent > maxcode) {
      nbits++;
      maxcode = (1 << nbits) - 1;
    } else if (incount >= checkpoint) {
      checkpoint = incount + CHECK_GAP;
      uint64_t rat;
      if (incount > 0x007fffffull) {
        rat = outcount >> 8;
        rat = rat == 0 ? 0x7fffffffull : incount / rat;
      } else {
        rat = (incount << 8) / outcount;
      }
      if (rat <= ratio)
        reset_after_clear(w);
      else
        ratio = rat;
    }
  }

  // LZWPostEncode; returns the strip's byte count or OVERFLOW
  LZW_HD uint32_t finish() {
    if (ent >= 0) {
      put(ent);
      free_ent++;
      if (free_ent == CODE_MAX - 1) {
        outcount = 0;
        put(CODE_CLEAR);
        nbits = BITS_MIN;
      } else if (free_ent > maxcode) {
        nbits++;
      }
    }
    put(CODE_EOI);
    while (nacc > 0) {
      const uint32_t b = nacc >= 8 ? (uint32_t)(acc >> (nacc - 8)) & 0xFFu : (uint32_t)(acc << (8 - nacc)) & 0xFFu;
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

// in: n bytes (any alignment); out: 4-byte aligned, cap bytes; table: ENC_SLOTS words
template <class W>
LZW_HD uint32_t encode_strip(const uint8_t* in, uint32_t n, uint8_t* out, uint32_t cap, uint32_t* table, const W& w) {
  Encoder e;
  e.begin(table, out, cap, w);
  if (n == 0) return e.finish();
  e.first(in[0]);
  uint32_t i = 1;
  // bytes up to the first 4-byte boundary of the input
  while (i < n && ((reinterpret_cast<uintptr_t>(in + i)) & 3u)) e.byte(in[i++], w);
  if (i + 4 <= n) {
    uint32_t next = load_u32(in + i);                      // one word of look-ahead hides the load
    for (; i + 4 <= n; i += 4) {
      const uint32_t cur = next;
      if (i + 8 <= n) next = load_u32(in + i + 4);
      e.byte(cur & 0xFFu, w);
      e.byte((cur >> 8) & 0xFFu, w);
      e.byte((cur >> 16) & 0xFFu, w);
      e.byte(cur >> 24, w);
    }
  }
  while (i < n) e.byte(in[i++], w);
  return e.finish();
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
  bool have_next;

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
  // total bits that really exist, to tell padding zeros from data
  int64_t avail = (int64_t)n_in * 8;
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
      if (w.lane == 0) obuf[shift + q] = (uint8_t)code;
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
        if (i < keep && w.lane == 0) dst[i] = (uint8_t)(e & 0xFFu);
        k = (int)(e >> 20);
      }
      if (w.lane == 0) dst[0] = (uint8_t)k;
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
