// Host build of the LZW strip coder of image_processing_suite_b200/csrc/tiff_lzw_core.cuh with 32
// threads standing in for the 32 lanes (ballot / shuffle / match / reduce / syncwarp through a
// barrier, shared-memory atomics through __atomic builtins).  Lets the CPU-only suite check the code the
// CUDA kernels run against Pillow/libtiff.  Test infrastructure only; libips.so does not
// contain this code path.
#include <stdint.h>
#include <stdlib.h>
#include <barrier>
#include <thread>
#include <vector>
#include "../../image_processing_suite_b200/csrc/tiff_lzw_core.cuh"

namespace {
struct Shared {
  std::barrier<> bar{32};
  uint32_t slot[32];
};
struct Warp32 {
  int lane;
  Shared* sh;
  static constexpr int n = 32;
  void sync() const { sh->bar.arrive_and_wait(); }
  uint32_t shfl(uint32_t v, uint32_t src) const {
    sh->slot[lane] = v;
    sh->bar.arrive_and_wait();
    const uint32_t r = sh->slot[src & 31u];
    sh->bar.arrive_and_wait();
    return r;
  }
  uint32_t reduce_or(uint32_t v) const {
    sh->slot[lane] = v;
    sh->bar.arrive_and_wait();
    uint32_t r = 0;
    for (int i = 0; i < 32; ++i) r |= sh->slot[i];
    sh->bar.arrive_and_wait();
    return r;
  }
  uint32_t match_any(uint32_t v) const {
    sh->slot[lane] = v;
    sh->bar.arrive_and_wait();
    uint32_t r = 0;
    for (int i = 0; i < 32; ++i) r |= (uint32_t)(sh->slot[i] == v) << i;
    sh->bar.arrive_and_wait();
    return r;
  }
  void prefetch(const void*) const {}
  void atomic_or(uint32_t* p, uint32_t v) const { __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
  uint32_t atomic_cas(uint32_t* p, uint32_t cmp, uint32_t v) const {
    __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
  }
  uint32_t ballot(bool p) const {
    sh->slot[lane] = p ? 1u : 0u;
    sh->bar.arrive_and_wait();
    uint32_t r = 0;
    for (int i = 0; i < 32; ++i) r |= sh->slot[i] << i;
    sh->bar.arrive_and_wait();
    return r;
  }
};
}  // namespace

extern "C" {
uint64_t harness_bound(uint64_t n) { return ips_lzw::encode_bound(n); }
uint32_t harness_encode(const uint8_t* in, uint32_t n, uint8_t* out, uint32_t cap) {
  uint32_t* table = (uint32_t*)aligned_alloc(16, ips_lzw::ENC_SLOTS * 4);
  uint32_t* stage = (uint32_t*)aligned_alloc(16, ips_lzw::PE_STAGE * 4);
  Shared sh;
  uint32_t ret[32];
  std::vector<std::thread> lanes;
  for (int l = 0; l < 32; ++l)
    lanes.emplace_back([&, l] {
      Warp32 w{l, &sh};
      ret[l] = ips_lzw::encode_strip(in, n, out, cap, table, stage, w);
    });
  for (auto& t : lanes) t.join();
  free(table); free(stage);
  for (int l = 1; l < 32; ++l)
    if (ret[l] != ret[0]) return 0xFFFFFFFEu;   // the lanes must agree
  return ret[0];
}
int harness_decode(const uint8_t* in, uint32_t n_in, uint8_t* out, uint32_t n_out) {
  uint16_t* orel = (uint16_t*)aligned_alloc(16, ips_lzw::PD_TAB * 2);
  uint32_t* obase = (uint32_t*)aligned_alloc(16, ips_lzw::PD_TAB / 16 * 4);
  uint8_t* win = (uint8_t*)aligned_alloc(16, ips_lzw::PD_WIN);
  Shared sh;
  int status[32];
  std::vector<std::thread> lanes;
  for (int l = 0; l < 32; ++l)
    lanes.emplace_back([&, l] {
      Warp32 w{l, &sh};
      status[l] = ips_lzw::decode_strip(in, n_in, out, n_out, orel, obase, win, w);
    });
  for (auto& t : lanes) t.join();
  for (int l = 1; l < 32; ++l)
    if (status[l] != status[0]) return -100 - l;   // the lanes must agree
  free(orel); free(obase); free(win);
  return status[0];
}
}
