// TIFF LZW strip coder, shared between the CUDA kernels of tiff.cu and the host harness of
// tests/native/lzw_host_harness.cpp (which compiles this header with g++, 32 threads standing in
// for the lanes and a barrier for the warp collectives, to check the very code the kernels run
// against Pillow/libtiff on the CPU-only authoring box).
//
// The stream format is TIFF 6.0 section 13 as libtiff writes it (the codec behind
// img.save(..., compression='tiff_lzw'), Image_re-binning.py:19-21): MSB-first codes of 9..12
// bits with the "early change", ClearCode first in every strip, table reset once code 4093 has
// been assigned, and a compression-ratio check every 10000 input bytes that may reset early.
// Following that policy to the letter makes the strips byte-identical to libtiff's, so the
// files can be compared with cmp, not only decoded.
//
// Execution model: one warp per strip, the 32 lanes working on 32 consecutive bytes (encoder) or
// 32 consecutive codes (decoder) at a time; see the two sections below.
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
enum : uint32_t { ENC_SLOTS = 5504, ENC_EMPTY = 0xFFFFFFFFu };      // open addressing, load <= 0.70; 21.5 KB: 10 strips per SM
enum : uint32_t { OVERFLOW = 0xFFFFFFFFu };
enum : int { ST_OK = 0, ST_TRUNCATED = 1, ST_CORRUPT = 2, ST_OLD_STYLE = 3 };

#ifdef __CUDACC__
struct Warp {
  int lane;
  __device__ __forceinline__ Warp() : lane(threadIdx.x & 31) {}
  static constexpr int n = 32;
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  __device__ __forceinline__ uint32_t ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
  __device__ __forceinline__ uint32_t shfl(uint32_t v, uint32_t src) const { return __shfl_sync(0xffffffffu, v, (int)src); }
  __device__ __forceinline__ uint32_t reduce_or(uint32_t v) const { return __reduce_or_sync(0xffffffffu, v); }
  __device__ __forceinline__ uint32_t match_any(uint32_t v) const { return __match_any_sync(0xffffffffu, v); }
  __device__ __forceinline__ void prefetch(const void* p) const { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
  __device__ __forceinline__ void atomic_or(uint32_t* p, uint32_t v) const { atomicOr(p, v); }
  __device__ __forceinline__ uint32_t atomic_cas(uint32_t* p, uint32_t cmp, uint32_t v) const { return atomicCAS(p, cmp, v); }
};
LZW_HD uint32_t ctz32(uint32_t v) { return (uint32_t)__ffs((int)v) - 1u; }
LZW_HD uint32_t popc32(uint32_t v) { return (uint32_t)__popc(v); }
LZW_HD uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }
LZW_HD uint32_t load_u32(const uint8_t* p) { return *reinterpret_cast<const uint32_t*>(p); }
struct alignas(16) Vec16 { uint32_t a, b, c, d; };
LZW_HD void copy16(uint8_t* dst, const uint8_t* src) {
  *reinterpret_cast<Vec16*>(dst) = *reinterpret_cast<const Vec16*>(src);
}
#else
// the host build supplies its own warp type (32 threads, tests/native/lzw_host_harness.cpp)
inline uint32_t ctz32(uint32_t v) { return (uint32_t)__builtin_ctz(v); }
inline uint32_t popc32(uint32_t v) { return (uint32_t)__builtin_popcount(v); }
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
// Dictionary: open addressing in shared memory, slot = (key << 12) | code with key = (prefix code
// << 8) | byte, linear probing, all-ones = empty.
#ifdef __CUDACC__
LZW_HD uint32_t slot_of(uint32_t key) { return __umulhi(key * 0x9E3779B1u, (uint32_t)ENC_SLOTS); }
#else
inline uint32_t slot_of(uint32_t key) { return (uint32_t)(((uint64_t)(key * 0x9E3779B1u) * ENC_SLOTS) >> 32); }
#endif

// Greedy LZW parsing is serial only through the phrase boundaries.  Each lane therefore walks the
// dictionary from its own byte of a 32-byte window as if a phrase started there; the boundaries
// are then found by pointer doubling over "start + length" (five shuffle rounds), and every lane
// that really starts a phrase emits its code (bit offsets by a warp scan, atomicOr into a staging
// word buffer) and inserts its new entry (atomicCAS along the probe path) at once.  The walks
// see the dictionary as of the window start; the only way an entry made inside the window can
// change a later phrase is by extending it at its failing (node, byte) pair, i.e. when two
// phrases of the window fail on the same pair -- the window is cut before the second one and
// the next window starts there.  Windows are also cut right after a phrase at which libtiff
// resets the table or evaluates the compression ratio, so those decisions are taken by uniform
// scalar code between windows with exactly libtiff's counters.
enum : uint32_t { PE_STAGE = 16 };   // staging words: 31 carried bits + 32 codes x 12 bits < 16 words

template <class W>
struct ParEncoder {
  const W& w;
  uint32_t lane;
  uint32_t* tab;
  uint32_t* stage;
  uint8_t* out;
  uint32_t cap, op, carry, obits;
  bool overflow;

  LZW_HD ParEncoder(const W& w_) : w(w_) {}

  // this lane's `wd` bits of `code` at bit offset `b` of the staged stream (wd == 0: nothing)
  LZW_HD void place(uint32_t code, uint32_t wd, uint32_t b) {
    if (wd) {
      const uint32_t word = b >> 5, o = b & 31u;
      if (o + wd <= 32u) {
        w.atomic_or(stage + word, code << (32u - o - wd));
      } else {
        w.atomic_or(stage + word, code >> (o + wd - 32u));
        w.atomic_or(stage + word + 1u, code << (64u - o - wd));
      }
    }
  }
  // `total` bits have been placed behind the carried ones: whole words go to global memory
  LZW_HD void flush(uint32_t total) {
    w.sync();
    const uint32_t pending = carry + total, full = pending >> 5;
    const uint32_t v = lane < PE_STAGE ? stage[lane] : 0u;
    const uint32_t part = w.shfl(v, full);
    if (op + 4u * full <= cap) {
      if (lane < full) *reinterpret_cast<uint32_t*>(out + op + 4u * lane) = bswap32(v);
    } else {
      overflow = true;
    }
    w.sync();
    if (lane < PE_STAGE) stage[lane] = lane == 0u ? part : 0u;
    w.sync();
    op += 4u * full;
    carry = pending & 31u;
    obits += total;
  }
  // bits of the first j codes of a window that starts with free_ent == F: 9 each, one more for
  // every code emitted while free_ent >= 512, 1024, 2048
  static LZW_HD uint32_t bits_before(uint32_t j, uint32_t F) {
    uint32_t b = 9u * j;
    for (uint32_t T = 512u; T <= 2048u; T <<= 1) {
      const int over = (int)(F + j) - (int)T;          // codes i < j with F + i >= T: clamp(F + j - T, 0, j)
      b += over <= 0 ? 0u : ((uint32_t)over < j ? (uint32_t)over : j);
    }
    return b;
  }
  // the codes of the m accepted phrases of a window (this lane: phrase j if acc)
  LZW_HD void emit_phrases(uint32_t code, bool acc, uint32_t j, uint32_t m, uint32_t F) {
    place(code, acc ? (uint32_t)(9 + (F + j >= 512u) + (F + j >= 1024u) + (F + j >= 2048u)) : 0u, carry + bits_before(j, F));
    flush(bits_before(m, F));
  }
  LZW_HD void emit1(uint32_t code, int width) {       // one code, uniform arguments
    place(code, lane == 0u ? (uint32_t)width : 0u, carry);
    flush((uint32_t)width);
  }

  LZW_HD void clear_table() {
    w.sync();
    uint64_t* t64 = reinterpret_cast<uint64_t*>(tab);
    for (uint32_t i = lane; i < ENC_SLOTS / 2; i += 32u) t64[i] = ~0ull;
    w.sync();
  }
};

LZW_HD int width_for(uint32_t free_ent) {      // code width while free_ent entries + specials exist
  return 9 + (free_ent >= 512u) + (free_ent >= 1024u) + (free_ent >= 2048u);
}

// in: n < 2^28 bytes; out: 4-byte aligned, cap bytes; table: ENC_SLOTS words; stage: PE_STAGE words.
// Needs a full warp (W::n == 32).  Returns the strip's byte count or OVERFLOW.
template <class W>
LZW_HD uint32_t encode_strip(const uint8_t* in, uint32_t n, uint8_t* out, uint32_t cap, uint32_t* table,
                                 uint32_t* stage, const W& w) {
  ParEncoder<W> e(w);
  const uint32_t lane = (uint32_t)w.lane;
  e.lane = lane;
  e.tab = table;
  e.stage = stage;
  e.out = out;
  e.cap = cap & ~3u;
  e.op = 0;
  e.carry = 0;
  e.obits = 0;
  e.overflow = false;
  uint32_t F = CODE_FIRST;          // free_ent
  int nbits = BITS_MIN;
  uint32_t in_base = 0, bits_base = 0, checkpoint = CHECK_GAP, ratio = 0;
  if (lane < PE_STAGE) stage[lane] = 0;
  e.clear_table();
  const uint32_t lt = (1u << lane) - 1u;

  auto reset = [&](uint32_t consumed) {       // libtiff: cl_hash, counters to zero, ClearCode at the old width
    e.clear_table();
    ratio = 0;
    in_base = consumed;
    bits_base = e.obits;
    F = CODE_FIRST;
    e.emit1(CODE_CLEAR, nbits);
    nbits = BITS_MIN;
  };

  uint32_t ent = 0;
  bool any = false;
  if (n > 0) {
    any = true;
    e.emit1(CODE_CLEAR, nbits);
    uint32_t x0 = 0;                // start of the next phrase
    for (;;) {
      // ---- every lane walks the dictionary from its own byte ---------------------------------
      // st: 0 walking, 1 the phrase of d bytes fails on `key` (h is the free slot), 2 ran into the
      // end of the strip (the last phrase) or lies beyond it.  One probe per iteration; the next
      // input byte is loaded one step ahead of its use.
      const uint32_t p = x0 + lane;
      const bool active = p < n;
      if (lane == 0u && x0 + 384u < n) w.prefetch(in + x0 + 384u);     // the input line three windows ahead
      uint32_t node = 0, d = 1, key = 0, h = 0, ahead = 0, st = 2;
      if (active) {
        node = in[p];
        if (p + 1u < n) {
          key = (node << 8) | in[p + 1u];
          h = slot_of(key);
          st = 0;
          if (p + 2u < n) ahead = in[p + 2u];
        }
      }
      // one probe per call, written without branches: the three outcomes (hit -> next byte, free slot
      // -> the phrase ends, other key -> next slot) would otherwise serialise the lanes
      auto probe = [&]() {
        const bool on = st == 0u;
        const uint32_t s = table[h];
        const bool hit = on && (s >> 12) == key;
        const bool stop = on && !hit && s == (uint32_t)ENC_EMPTY;
        if (hit) {
          node = s & 0xFFFu;
          ++d;
        }
        const bool end = hit && p + d >= n;
        const uint32_t nkey = (node << 8) | ahead;
        const uint32_t nslot = slot_of(nkey);
        const uint32_t step = h + 1u == ENC_SLOTS ? 0u : h + 1u;
        key = hit ? nkey : key;
        h = hit ? nslot : ((on && !stop) ? step : h);
        st = stop ? 1u : (end ? 2u : st);
        if (hit && p + d + 1u < n) ahead = in[p + d + 1u];
      };
      while (w.ballot(st == 0u)) {
        probe();
        probe();
        probe();
        probe();
      }
      const bool open = active && st == 2u;
      // lanes failing on the same (node, byte) pair; issued here so that its latency hides behind the
      // doubling rounds below
      const uint32_t same = w.match_any(st == 1u ? key : (0x80000000u | lane));
      // ---- phrase starts reachable from lane 0: pointer doubling -------------------------------
      const uint32_t nxt = lane + d;                                   // window-relative end of this lane's phrase
      uint32_t jump = (!active || open || nxt > 31u) ? 32u : nxt;
      uint32_t reach = 1u;
      for (int r = 0; r < 5; ++r) {
        reach |= w.reduce_or((((reach >> lane) & 1u) && jump < 32u) ? (1u << jump) : 0u);
        const uint32_t jj = w.shfl(jump, jump & 31u);
        jump = jump < 32u ? jj : 32u;
      }
      const bool start = (reach >> lane) & 1u;
      const uint32_t j = popc32(reach & lt);      // phrase index within the window
      uint32_t m = popc32(reach);                // phrases to accept
      // ---- where to cut ---------------------------------------------------------------------------
      const bool later_dup = start && st == 1u && (same & reach & lt) != 0u;   // an earlier phrase fails on the same pair
      const uint32_t f1 = F + j + 1u;                                  // free_ent after this phrase's entry
      const bool limit_ev = f1 == 512u || f1 == 1024u || f1 == 2048u || f1 == (uint32_t)CODE_MAX - 1u;
      const bool reset_ev = f1 == (uint32_t)CODE_MAX - 1u;
      const uint32_t consumed = p + d + 1u;                            // bytes consumed when the phrase fails
      const bool ratio_ev = !limit_ev && consumed - in_base >= checkpoint;
      const uint32_t before = w.ballot(start && (open || later_dup));
      const uint32_t after = w.ballot(start && !open && !later_dup && (reset_ev || ratio_ev));
      if (before) {
        const uint32_t jb = popc32(reach & ((1u << ctz32(before)) - 1u));
        if (jb < m) m = jb;
      }
      if (after) {
        const uint32_t ja = popc32(reach & ((1u << ctz32(after)) - 1u)) + 1u;
        if (ja < m) m = ja;
      }
      if (m == 0u) {                    // lane 0's phrase runs to the end of the strip
        ent = w.shfl(node, 0u);
        break;
      }
      // ---- accepted phrases: emit codes, insert entries ------------------------------------------------
      const bool acc = start && j < m;
      e.emit_phrases(node, acc, j, m, F);
      if (acc) {
        const uint32_t val = (key << 12) | (F + j);
        while (w.atomic_cas(table + h, (uint32_t)ENC_EMPTY, val) != (uint32_t)ENC_EMPTY) h = h + 1u == ENC_SLOTS ? 0u : h + 1u;
      }
      w.sync();
      const uint32_t last = ctz32(w.ballot(acc && j == m - 1u));
      const uint32_t tail = w.shfl(nxt | (reset_ev ? 0x10000u : 0u) | (ratio_ev ? 0x20000u : 0u), last);
      x0 += tail & 0xFFFFu;                                            // nxt <= 31 + 3839
      const uint32_t cons = x0 + 1u;                                   // = consumed of the last accepted phrase
      const uint32_t ev = tail >> 16;
      F += m;
      nbits = width_for(F);
      if (ev & 1u) {
        reset(cons);
      } else if (ev & 2u) {
        const uint32_t incount = cons - in_base, outcount = e.obits - bits_base;
        checkpoint = incount + CHECK_GAP;
        uint32_t rat;
        if (incount > 0x007fffffu) {
          rat = outcount >> 8;
          rat = rat == 0 ? 0x7fffffffu : incount / rat;
        } else {
          rat = (incount << 8) / outcount;
        }
        if (rat <= ratio)
          reset(cons);
        else
          ratio = rat;
      }
    }
  }
  // ---- LZWPostEncode ------------------------------------------------------------------------------
  if (any) {
    e.emit1(ent, nbits);
    F++;
    if (F == (uint32_t)CODE_MAX - 1u) {
      e.emit1(CODE_CLEAR, nbits);
      nbits = BITS_MIN;
    } else if (F == 512u || F == 1024u || F == 2048u) {
      nbits++;
    }
  }
  e.emit1(CODE_EOI, nbits);
  const uint32_t tail = (e.carry + 7u) >> 3;
  if (e.op + tail <= cap) {
    if (lane < tail) out[e.op + lane] = (uint8_t)(stage[0] >> (24u - 8u * lane));
  } else {
    e.overflow = true;
  }
  w.sync();
  return e.overflow ? (uint32_t)OVERFLOW : e.op + tail;
}

// ------------------------------------------------------------------------------------------
// decoder
// ------------------------------------------------------------------------------------------
// A serial LZW decoder spends ~600 cycles per code on one warp.  This one parses 32 codes per
// step, one per lane, using three facts about a table generation (the codes between two
// ClearCodes):
//   1. the width of the k-th code after a ClearCode depends on k alone (9 bits for k < 254,
//      10 for k < 766, 11 for k < 1790, then 12), so every lane knows where its code starts;
//   2. the string of code 258 + t is the string of the t-th code of the generation followed by
//      the first byte of the (t+1)-th: those bytes are adjacent in the output, so the string is
//      out[O_t .. O_t + L_t + 1), where O_t is the output offset of the t-th code.  One table
//      of offsets per generation (O_t; 8.4 KB of shared memory as 16-bit offsets within groups
//      of 16 codes) replaces the prefix/suffix dictionary; lengths are differences of neighbours;
//   3. lengths depend on earlier lengths only (L_j = L_t + 1), offsets are their prefix sum:
//      a few shuffle rounds and a warp scan per 32 codes.
// Bytes go through an 8 KB circular window in shared memory (recent output, which is where
// most references point) and leave for global memory in 16-byte vectors; references that fell
// out of the window are read back from global memory.  Chunks whose strings are short are
// copied one lane per code, in passes ordered by the in-chunk dependencies; chunks with long
// strings are copied code by code with the lanes striding over the bytes.
enum : uint32_t { PD_WIN = 8192, PD_TAB = 3840, PD_PAR_MAX = 512 };

LZW_HD uint32_t code_bitpos(uint32_t k) {     // bit offset of the k-th code after a ClearCode
  if (k <= 254u) return 9u * k;
  if (k <= 766u) return 9u * 254u + 10u * (k - 254u);
  if (k <= 1790u) return 9u * 254u + 10u * 512u + 11u * (k - 766u);
  return 9u * 254u + 10u * 512u + 11u * 1024u + 12u * (k - 1790u);
}
LZW_HD int code_width(uint32_t k) { return k < 254u ? 9 : k < 766u ? 10 : k < 1790u ? 11 : 12; }

struct BitReader {
  const uint8_t* base;   // 4-byte aligned address at or before the first byte
  uint32_t skew;         // bits between base and the first byte
  uint32_t bytes;        // bytes from base to the end of the strip
  LZW_HD void begin(const uint8_t* in, uint32_t n_in) {
    const uint32_t a = (uint32_t)(reinterpret_cast<uintptr_t>(in) & 3u);
    base = in - a;
    skew = 8u * a;
    bytes = n_in + a;
  }
  // word idx as loaded (little-endian register image of the four stream bytes), zero past the end;
  // the byte swap is left to code() so that a prefetched word is not waited for before its use
  LZW_HD uint32_t word(uint32_t idx) const {
    if (idx * 4u + 4u <= bytes) return load_u32(base + idx * 4u);
    uint32_t v = 0;
    for (uint32_t k = 0; k < 4u; ++k)
      if (idx * 4u + k < bytes) v |= (uint32_t)base[idx * 4u + k] << (8u * k);
    return v;
  }
  // the two words that hold the code at bitpos (counted from the first byte) ...
  LZW_HD void fetch(uint32_t bitpos, uint32_t& hi, uint32_t& lo) const {
    const uint32_t idx = (bitpos + skew) >> 5;
    hi = word(idx);
    lo = word(idx + 1u);
  }
  // ... and the code itself
  LZW_HD uint32_t code(uint32_t bitpos, int width, uint32_t hi, uint32_t lo) const {
    const uint64_t both = ((uint64_t)bswap32(hi) << 32) | bswap32(lo);
    return (uint32_t)((both << ((bitpos + skew) & 31u)) >> (64 - width));
  }
};

// Decodes one strip to out[0..n_out) with a full warp (W::n == 32).  orel: PD_TAB halfwords, obase: PD_TAB / 16 words, win:
// PD_WIN bytes, 16-byte aligned.  n_in < 2^28.  Returns a status; on any failure the undecoded
// remainder is zero-filled so the output is deterministic.
template <class W>
LZW_HD int decode_strip(const uint8_t* in, uint32_t n_in, uint8_t* out, uint32_t n_out, uint16_t* orel, uint32_t* obase, uint8_t* win,
                        const W& w) {
  const uint32_t lane = (uint32_t)w.lane;
  int status = ST_OK;
  if (n_in >= 2 && in[0] == 0 && (in[1] & 1)) status = ST_OLD_STYLE;   // pre-6.0 LSB-first streams
  if (n_in >= (1u << 29)) status = ST_CORRUPT;                          // bit positions are 32-bit
  BitReader br;
  br.begin(in, n_in);
  const uint32_t total_bits = n_in * 8u;
  const uint32_t A = (uint32_t)(reinterpret_cast<uintptr_t>(out) & 15u);   // window index = (pos + A) mod PD_WIN
  uint8_t* const gA = out - A;
  uint32_t pos = 0;          // bytes decoded (all of them written to the window or beyond)
  uint32_t flushed = 0;      // bytes [0, flushed) are in global memory
  uint32_t gen_bit0 = 0;     // bit position of code 0 of the current generation
  uint32_t kbase = 0;        // generation index of lane 0's code
  uint32_t pre_hi = 0, pre_lo = 0, pre_bit0 = 0xFFFFFFFFu, pre_kbase = 0;   // prefetched input words

  // window bytes [flushed, frontier) -> global memory, whole 16-byte vectors unless final
  auto flush = [&](uint32_t frontier, bool final) {
    w.sync();
    const uint32_t v0 = flushed + A;
    uint32_t v1 = frontier + A;
    if (!final) v1 &= ~15u;
    if (v1 > v0) {
      uint32_t head_end = (v0 + 15u) & ~15u;
      if (head_end > v1) head_end = v1;
      for (uint32_t v = v0 + lane; v < head_end; v += 32u) gA[v] = win[v & (PD_WIN - 1u)];
      const uint32_t ve = v1 & ~15u;
      for (uint32_t v = head_end + 16u * lane; v < ve; v += 512u) copy16(gA + v, win + (v & (PD_WIN - 1u)));
      const uint32_t tail = ve > head_end ? ve : head_end;
      for (uint32_t v = tail + lane; v < v1; v += 32u) gA[v] = win[v & (PD_WIN - 1u)];
      flushed = v1 - A;
    }
    w.sync();
  };

  while (status == ST_OK && pos < n_out) {
    // ---- this lane's code ----------------------------------------------------------------
    const uint32_t k = kbase + lane;
    const int width = code_width(k);
    const uint32_t bp = gen_bit0 + code_bitpos(k);
    const bool over = bp + (uint32_t)width > total_bits;
    uint32_t hi = pre_hi, lo = pre_lo;
    if (pre_bit0 != gen_bit0 || pre_kbase != kbase) br.fetch(bp, hi, lo);   // a ClearCode moved the generation
    const uint32_t c = over ? (uint32_t)CODE_EOI : br.code(bp, width, hi, lo);
    // the words of this lane's code in the next chunk, assuming no ClearCode in this one: in flight
    // while the lengths, offsets and copies below run
    pre_bit0 = gen_bit0;
    pre_kbase = kbase + 32u;
    br.fetch(gen_bit0 + code_bitpos(k + 32u), pre_hi, pre_lo);
    const uint32_t endmask = w.ballot(over || c == (uint32_t)CODE_CLEAR || c == (uint32_t)CODE_EOI);
    uint32_t nvalid = endmask ? ctz32(endmask) : 32u;
    const bool lit = c < 256u;
    const int t = (int)c - CODE_FIRST;                     // generation index of the code that made this entry
    const uint32_t badmask = w.ballot(lane < nvalid && !lit && t > (int)k - 1);
    bool corrupt = false;
    if (badmask && ctz32(badmask) < nvalid) {
      nvalid = ctz32(badmask);
      corrupt = true;
    }
    const bool valid = lane < nvalid;
    const uint32_t rel = (uint32_t)(t - (int)kbase) & 31u;   // source lane when the entry was made in this chunk
    const bool far = t < (int)kbase;                         // entry made in an earlier chunk

    // ---- string lengths: L = L_t + 1 --------------------------------------------------------
    uint32_t len = (valid && lit) ? 1u : 0u;
    bool have = !valid || lit;
    if (valid && !lit && far) {
      len = (obase[(t + 1) >> 4] + orel[t + 1]) - (obase[t >> 4] + orel[t]) + 1u;
      have = true;
    }
    while (w.ballot(!have)) {
      const uint32_t ls = w.shfl(len, rel);
      const uint32_t hs = w.shfl(have ? 1u : 0u, rel);
      if (!have && hs) {
        len = ls + 1u;
        have = true;
      }
    }
    // ---- output offsets: exclusive scan ---------------------------------------------------------
    uint32_t incl = len;
    for (uint32_t d = 1; d < 32u; d <<= 1) {
      const uint32_t up = w.shfl(incl, (lane - d) & 31u);
      if (lane >= d) incl += up;
    }
    const uint32_t total = w.shfl(incl, 31u);
    const uint32_t off = pos + incl - len;                    // O_k
    // O_k = obase[k / 16] + orel[k]: 16 consecutive strings span at most 16 * 3839 bytes
    const uint32_t group0 = w.shfl(off, lane & 16u);
    if (valid && k < PD_TAB) {
      orel[k] = (uint16_t)(off - group0);
      if ((lane & 15u) == 0u) obase[k >> 4] = off;
    }
    if (lane == 0 && nvalid == 32u && kbase + 32u < PD_TAB) {   // the next chunk's first offset closes this one's last length
      orel[kbase + 32u] = 0;
      obase[(kbase + 32u) >> 4] = pos + total;
    }
    const uint32_t room = n_out - pos;
    const uint32_t chunk = total < room ? total : room;       // bytes this chunk writes
    uint32_t cp = 0;                                          // bytes this lane's code writes
    if (valid && off < n_out) cp = len < n_out - off ? len : n_out - off;
    w.sync();
    // source offset O_t of the string this code copies
    const uint32_t src_far = (valid && !lit && far) ? obase[t >> 4] + orel[t] : 0u;
    const uint32_t src_near = w.shfl(off, rel);
    const uint32_t src = far ? src_far : src_near;

    // ---- copy -----------------------------------------------------------------------------------
    if (chunk <= PD_PAR_MAX) {
      const uint32_t wr_end = pos + chunk;
      if (wr_end - flushed > PD_WIN - 16u) flush(pos, false);
      bool done = cp == 0u;
      if (!done && lit) {
        win[(off + A) & (PD_WIN - 1u)] = (uint8_t)c;
        done = true;
      }
      for (;;) {
        w.sync();
        const uint32_t dm = w.ballot(done);
        if (dm == 0xFFFFFFFFu) break;
        if (!done) {
          const bool dep1 = far || ((dm >> rel) & 1u);
          const uint32_t u = (uint32_t)t + 1u;                // the code that supplies the last byte
          const bool dep2 = u == k || u < kbase || ((dm >> ((u - kbase) & 31u)) & 1u);
          if (dep1 && dep2) {
            const uint32_t period = off - src;
            for (uint32_t i = 0; i < cp; ++i) {
              uint32_t sp = src + i;
              if (sp >= off) sp -= period;                    // the entry being defined right now (KwKwK)
              const uint8_t b = sp + PD_WIN > wr_end ? win[(sp + A) & (PD_WIN - 1u)] : out[sp];
              win[(off + i + A) & (PD_WIN - 1u)] = b;
            }
            done = true;
          }
        }
      }
    } else {
      for (uint32_t q = 0; q < nvalid; ++q) {
        const uint32_t cq = w.shfl(c, q), oq = w.shfl(off, q), nq = w.shfl(cp, q), sq = w.shfl(src, q);
        if (nq == 0u) break;
        const uint32_t wr_end = oq + nq;
        if (wr_end - flushed > PD_WIN - 16u) flush(oq, false);
        if (cq < 256u) {
          if (lane == 0) win[(oq + A) & (PD_WIN - 1u)] = (uint8_t)cq;
        } else {
          const uint32_t period = oq - sq;
          for (uint32_t i = lane; i < nq; i += 32u) {
            uint32_t sp = sq + i;
            if (sp >= oq) sp -= period;
            const uint8_t b = sp + PD_WIN > wr_end ? win[(sp + A) & (PD_WIN - 1u)] : out[sp];
            win[(oq + i + A) & (PD_WIN - 1u)] = b;
          }
        }
        w.sync();
      }
    }
    pos += chunk;
    if (pos >= n_out) break;
    if (corrupt) {
      status = ST_CORRUPT;
      break;
    }
    if (nvalid < 32u) {                                       // a terminator sits in lane nvalid
      const uint32_t ce = w.shfl(c, nvalid);
      const uint32_t oe = w.shfl(over ? 1u : 0u, nvalid);
      if (oe || ce == (uint32_t)CODE_EOI) {
        status = ST_TRUNCATED;
        break;
      }
      gen_bit0 = w.shfl(bp, nvalid) + (uint32_t)w.shfl((uint32_t)width, nvalid);
      kbase = 0;
    } else {
      kbase += 32u;
    }
  }
  flush(pos, true);
  if (pos < n_out) {
    if (status == ST_OK) status = ST_TRUNCATED;
    for (uint32_t i = pos + lane; i < n_out; i += 32u) out[i] = 0;
  }
  return status;
}

}  // namespace ips_lzw
