// Host-side helpers of the C ABI (no device code).
//
// exp_crc32c: CRC-32C (Castagnoli, reflected polynomial 0x82F63B78), the checksum TensorFlow's
// checkpoint bundles carry per table block and per tensor (tensor_bundle.cc; read / written by
// exposure_b200/tf_bundle.py for the reference's tf.train.Saver files, net.py:271,383-387,405-407).
// Slicing-by-8, ~1 GB/s on one core: the 34 MB of a full checkpoint take tens of milliseconds
// instead of the tens of seconds of a Python byte loop.
#include <cstddef>
#include <cstdint>

namespace {

struct Crc32cTables {
  uint32_t t[8][256];
  Crc32cTables() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xFF];
  }
};

}  // namespace

extern "C" uint32_t exp_crc32c(uint32_t crc, const void* data, size_t n) {
  static const Crc32cTables tab;
  const unsigned char* p = static_cast<const unsigned char*>(data);
  uint32_t c = crc ^ 0xFFFFFFFFu;
  while (n >= 8) {
    const uint32_t lo = c ^ ((uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24);
    c = tab.t[7][lo & 0xFF] ^ tab.t[6][(lo >> 8) & 0xFF] ^ tab.t[5][(lo >> 16) & 0xFF] ^ tab.t[4][lo >> 24] ^
        tab.t[3][p[4]] ^ tab.t[2][p[5]] ^ tab.t[1][p[6]] ^ tab.t[0][p[7]];
    p += 8;
    n -= 8;
  }
  while (n--) c = tab.t[0][(c ^ *p++) & 0xFF] ^ (c >> 8);
  return c ^ 0xFFFFFFFFu;
}
