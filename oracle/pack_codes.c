/* Plain-C restatement of nlsh/utils.pyx:6-32 (binarr_to_int looped by hash_codes) — TEST
 * INFRASTRUCTURE (oracle), not product code.  bits: C-contiguous int [n][s][hs] of 0/1;
 * out: int16 [n][s].  The accumulator is a 32-bit int and the value is returned through an
 * int16, exactly as `cdef np.int16_t binarr_to_int(int[:] binarr)` does. */
#include <stdint.h>

void oracle_pack_codes(const int *bits, long n, long s, long hs, int16_t *out) {
  for (long i = 0; i < n; ++i) {
    for (long j = 0; j < s; ++j) {
      const int *row = bits + (i * s + j) * hs;
      uint32_t acc = 0; /* same bit pattern as the reference's np.int32_t, without signed-shift UB */
      for (long b = 0; b < hs; ++b) acc = (acc << 1) | (uint32_t)row[b];
      out[i * s + j] = (int16_t)(uint16_t)(acc & 0xffffu);
    }
  }
}
