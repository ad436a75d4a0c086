// Small host-only helpers of the C ABI that have nothing to do with the GPU.
#include <cstddef>
#include <cstdint>

#include "../../include/n2nmn_b200.h"

namespace {
struct Crc32cTable {
  uint32_t t[256];
  Crc32cTable() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      t[i] = c;
    }
  }
};
}  // namespace

extern "C" uint32_t n2nmn_crc32c(const void* data, size_t n, uint32_t crc) {
  static const Crc32cTable tab;
  const uint8_t* p = static_cast<const uint8_t*>(data);
  uint32_t c = ~crc;
  for (size_t i = 0; i < n; ++i) c = tab.t[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  return ~c;
}
