// Host build of the LZW strip coder of image_processing_suite_b200/csrc/tiff_lzw_core.cuh with a
// 1-lane "warp": lets the CPU-only suite check the exact state machines the CUDA kernels run
// against Pillow/libtiff.  Test infrastructure only; libips.so does not contain this code path.
#include <stdint.h>
#include <stdlib.h>
#include "../../image_processing_suite_b200/csrc/tiff_lzw_core.cuh"

extern "C" {
uint64_t harness_bound(uint64_t n) { return ips_lzw::encode_bound(n); }
uint32_t harness_encode(const uint8_t* in, uint32_t n, uint8_t* out, uint32_t cap) {
  uint32_t* table = (uint32_t*)aligned_alloc(16, ips_lzw::ENC_SLOTS * 4);
  ips_lzw::Warp w;
  const uint32_t r = ips_lzw::encode_strip(in, n, out, cap, table, w);
  free(table);
  return r;
}
int harness_decode(const uint8_t* in, uint32_t n_in, uint8_t* out, uint32_t n_out) {
  uint32_t* tab = (uint32_t*)aligned_alloc(16, ips_lzw::DEC_CODES * 4);
  uint8_t* firstc = (uint8_t*)malloc(ips_lzw::DEC_CODES);
  uint8_t* obuf = (uint8_t*)aligned_alloc(16, ips_lzw::DEC_OBUF);
  ips_lzw::Warp w;
  const int st = ips_lzw::decode_strip(in, n_in, out, n_out, tab, firstc, obuf, w);
  free(tab); free(firstc); free(obuf);
  return st;
}
}
