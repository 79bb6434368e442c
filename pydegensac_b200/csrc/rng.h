// rng.h -- counter-based sampling stream (Philox4x32-10).
//
// Replaces the reference's libc rand()/random() re-seeded every iteration
// (exp_ranF.c:1277,1331-1342; exp_ranH.c:510,539-552; rtools.c:12-39).  Contract:
//   value31(seed,k,j) = philox4x32_10(ctr=(j>>2, k, 0, 0), key=(seed_lo, seed_hi))[j&3] >> 1
//   iteration k>=1, draws j=0..m-1 : minimal sample, stateless partial Fisher-Yates over a fresh
//                                    identity pool with s_i = value31 % (N-i)  (cf. rtools.c:12-23)
//   draw j=m                        : the reference's `seed = rand()` slot (unused)
//   draws j>m                       : LO / DEGENSAC draws (randsubset, rFtH, dual_sample), raw 31-bit values
// Any hypothesis k can therefore be generated independently by any thread.
#pragma once
#include "common.h"

namespace dg {

DG_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                         uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

DG_HD uint32_t value31(uint64_t seed, uint32_t k, uint32_t j) {
  uint32_t o[4];
  philox4x32_10(j >> 2, k, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), o);
  return o[j & 3] >> 1;
}

// Minimal sample of iteration k: M distinct indices in DRAW order.
// (The reference's samidx[] is the reverse: samidx[t] = sel[M-1-t], because the sample lives in the
//  last M pool slots: exp_ranF.c:1302, rtools.c:12-23.)
template <int M>
DG_HD void minimal_sample(uint64_t seed, uint32_t k, int N, int* sel) {
  uint32_t r[8];
  philox4x32_10(0u, k, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
  if (M > 4) philox4x32_10(1u, k, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r + 4);
  // Virtual pool: slots touched by earlier swaps are logged as (position, value) pairs, two per draw at FIXED
  // log indices (2i, 2i+1), later entries overriding earlier ones -- every index below is a compile-time
  // constant after unrolling, so the log lives in registers.
  int tp[2 * M], tv[2 * M];
#pragma unroll
  for (int i = 0; i < M; ++i) {
    const int s = (int)((r[i] >> 1) % (uint32_t)(N - i));
    const int top = N - i - 1;
    int vs = s, vt = top;
#pragma unroll
    for (int t = 0; t < 2 * i; ++t) {
      if (tp[t] == s) vs = tv[t];
      if (tp[t] == top) vt = tv[t];
    }
    tp[2 * i] = s;       tv[2 * i] = vt;        // pool[s]   <- value that sat at the top slot
    tp[2 * i + 1] = top; tv[2 * i + 1] = vs;    // pool[top] <- the drawn value
    sel[i] = vs;
  }
}

// Sequential draw cursor for the LO / DEGENSAC draws of one iteration.
struct DrawCursor {
  uint64_t seed;
  uint32_t k;
  uint32_t j;
};
DG_HD uint32_t next_draw(DrawCursor& c) {
  const uint32_t v = value31(c.seed, c.k, c.j);
  ++c.j;
  return v;
}

}  // namespace dg
