// Host-side packers for the compact upload formats of the C ABI (include/phylo_engine.h): data-format
// helpers for the caller's loader, no arithmetic of the scoring path. An OCaml loader calls them once
// when it builds the alignment Bigarrays; the per-evaluation uploads then move half the bytes.
#include <cstring>

#include "phylo_engine.h"

// one byte per cell -> two 4-bit masks per byte (pattern 2j in the low nibble of byte j)
extern "C" int phylo_pack_nibbles(const uint8_t *masks, int T, int64_t N, uint8_t *packed) {
  if (!masks || !packed || T < 1 || N < 1) return PHYLO_ERR_ARG;
  const int64_t nb = (N + 1) / 2;
  for (int t = 0; t < T; ++t) {
    const uint8_t *in = masks + (size_t)t * N;
    uint8_t *out = packed + (size_t)t * nb;
    for (int64_t j = 0; j < N / 2; ++j) out[j] = (uint8_t)((in[2 * j] & 15) | ((in[2 * j + 1] & 15) << 4));
    if (N & 1) out[nb - 1] = (uint8_t)((in[N - 1] & 15) | 0xf0);
  }
  return PHYLO_OK;
}

extern "C" int phylo_fitch_plane_count(int n_states) {
  static const int sizes[] = {1, 2, 3, 4, 5, 6, 8, 12, 16, 24, 32, 64};
  for (int s : sizes)
    if (n_states <= s) return s;
  return n_states <= 64 ? 64 : -1;
}

// reference layout (one character per element of elt_bytes, lib/bitvector/bv.h:29-55) -> the bit-sliced
// device layout: per taxon ceil(N/32) words x NP planes of uint32, planes[w * NP + s] bit c <=> state s
// is in the set of character 32 w + c. NP = phylo_fitch_plane_count(n_states).
extern "C" int phylo_fitch_pack_planes(const void *codes, int elt_bytes, int n_states, int T, int64_t N, uint32_t *planes) {
  const int NP = phylo_fitch_plane_count(n_states);
  if (!codes || !planes || T < 1 || N < 1 || NP < 0 || !(elt_bytes == 1 || elt_bytes == 2 || elt_bytes == 4 || elt_bytes == 8) ||
      n_states < 1 || n_states > elt_bytes * 8)
    return PHYLO_ERR_ARG;
  const int64_t words = (N + 31) / 32;
  std::memset(planes, 0, sizeof(uint32_t) * (size_t)T * words * NP);
  for (int t = 0; t < T; ++t) {
    const unsigned char *row = (const unsigned char *)codes + (size_t)t * N * elt_bytes;
    uint32_t *out = planes + (size_t)t * words * NP;
    for (int64_t i = 0; i < N; ++i) {
      uint64_t v = 0;
      std::memcpy(&v, row + (size_t)i * elt_bytes, elt_bytes);  // little endian, like the Bigarray element
      uint32_t *w = out + (i >> 5) * NP;
      const uint32_t bit = 1u << (i & 31);
      for (int s = 0; s < n_states; ++s)
        if ((v >> s) & 1) w[s] |= bit;
    }
  }
  return PHYLO_OK;
}
