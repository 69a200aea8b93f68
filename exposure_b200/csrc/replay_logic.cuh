// Selection logic of the replay memory (replay_memory.py:64-75, 187-196, 230-273) as plain integer code that
// compiles for the device (one thread of a one-block kernel: the pool has 128 records) AND for the host
// (tests/host_math/replay_harness.cpp drives exactly these functions on the CPU).
//
// Data model.  One buffer of records, three regions addressed by a flat index:
//     [0, P)                  the pool                                   (P = cfg.replay_memory_size)
//     [P, P + B)              the outputs of the last generator step    (B = cfg.batch_size)
//     [P + B, P + B + F)      fresh RAW records of this iteration       (F = P + B, states == 0)
// Every operation of the reference becomes an index list into that buffer; the images / states move by gathers.
//
//   draw_generator   get_next_fake_batch (230-246): shuffle the pool; pop records from the front until B
//                    non-terminated ones are found, dropping the terminated ones met on the way; if the pool runs
//                    dry, fill_pool (64-75) replaces it by fresh records and the walk continues there.
//   replace          replace_memory (187-196): append the generator's outputs whose step is below
//                    maximum_trajectory_length (others with probability over_length_keep_prob), fill_pool up to P with
//                    fresh records, truncate to P.  (The reference's two further shuffles only randomise an order that
//                    every later draw re-randomises.)
//   draw_critic      replay_fake_batch (249-273): shuffle; the first B TERMINATED records in that order, cycling
//                    through them when there are fewer than B (the reference's while / for construct); none at all
//                    raises an assertion there and sets an error flag here.
//
// Randomness: Philox-4x32-10 keyed by (seed, call counter) -- reproducible for a seed, identical on host and device,
// and the counter lives in device memory so that a captured CUDA graph draws new numbers at every replay.  (The
// reference's own sequence comes from Python's Mersenne Twister; exposure_b200.replay.ReplayMemory keeps reproducing
// THAT draw for draw in host mode -- tests/test_replay_reference.py.)
#pragma once
#include <stdint.h>

#ifndef EXPO_RL_HD
#ifdef __CUDACC__
#define EXPO_RL_HD __host__ __device__ __forceinline__
#else
#define EXPO_RL_HD inline
#endif
#endif

namespace expo {
namespace rl {

constexpr int kStateStopped = 1;     // util.py:13-16
constexpr int kStateStep = 2;

struct Philox {
  uint32_t key[2];
  uint32_t ctr[4];
  uint32_t out[4];
  int have;
};
EXPO_RL_HD void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0], n1 = (uint32_t)p1;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1], n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}
EXPO_RL_HD void philox_init(Philox& g, uint64_t seed, uint64_t call, uint32_t stream) {
  g.key[0] = (uint32_t)seed; g.key[1] = (uint32_t)(seed >> 32);
  g.ctr[0] = 0; g.ctr[1] = stream; g.ctr[2] = (uint32_t)call; g.ctr[3] = (uint32_t)(call >> 32);
  g.have = 0;
}
EXPO_RL_HD uint32_t philox_next(Philox& g) {
  if (g.have == 0) {
    uint32_t c[4] = {g.ctr[0], g.ctr[1], g.ctr[2], g.ctr[3]};
    uint32_t k[2] = {g.key[0], g.key[1]};
    for (int r = 0; r < 10; ++r) philox_round(c, k);
    g.out[0] = c[0]; g.out[1] = c[1]; g.out[2] = c[2]; g.out[3] = c[3];
    g.ctr[0] += 1;
    g.have = 4;
  }
  return g.out[--g.have];
}
// uniform integer in [0, n) (multiply-shift; the bias of 2^-32 n is irrelevant for n <= 1024)
EXPO_RL_HD uint32_t philox_below(Philox& g, uint32_t n) { return (uint32_t)(((uint64_t)philox_next(g) * n) >> 32); }
EXPO_RL_HD float philox_uniform(Philox& g) { return (float)(philox_next(g) >> 8) * (1.0f / 16777216.0f); }   // [0, 1)

// Random permutation by sorting random keys: record i draws key_i = Philox(seed, call, stream; counter i) and lands
// at position rank(i) = #{j : key_j < key_i, ties by index}.  Every key and every rank is independent of the others,
// so the device computes them with one thread per record (shuffle_key / shuffle_rank below, ~1 us for 128 records
// instead of ~20 us for a serial Fisher-Yates on one thread); the host build loops over i.  Same permutation either way.
EXPO_RL_HD uint32_t shuffle_key(uint64_t seed, uint64_t call, uint32_t stream, int i) {
  Philox g;
  philox_init(g, seed, call, stream);
  g.ctr[0] = (uint32_t)i;
  g.ctr[1] ^= 0x80000000u;                    // a counter space of its own, away from the sequential draws of the call
  return philox_next(g);
}
EXPO_RL_HD int shuffle_rank(const uint32_t* keys, int n, int i) {
  const uint32_t k = keys[i];
  int r = 0;
  for (int j = 0; j < n; ++j) r += (keys[j] < k || (keys[j] == k && j < i)) ? 1 : 0;
  return r;
}
// serial form (host build, and any caller without a thread per record)
EXPO_RL_HD void shuffle(int* perm, uint32_t* keys, int n, uint64_t seed, uint64_t call, uint32_t stream) {
  for (int i = 0; i < n; ++i) keys[i] = shuffle_key(seed, call, stream, i);
  for (int i = 0; i < n; ++i) perm[shuffle_rank(keys, n, i)] = i;
}

// get_next_fake_batch.  stopped[P]: the STOPPED column of the pool's states (the kernels stage it in shared memory --
// the walk below is serial, and a dependent global load per record would cost more than everything else); perm: a
// random order of the pool (shuffle).  Outputs: batch_src[B] (flat indices), rest_src[<= P] and *n_rest (the records
// that stay in the pool, flat indices), *fresh_used (fresh records consumed so far this iteration).
EXPO_RL_HD void draw_generator(const float* stopped, int P, int B, const int* perm, long long* batch_src,
                               int* rest_src, int* n_rest, int* fresh_used) {
  int taken = 0, nr = 0;
  for (int i = 0; i < P; ++i) {
    const int r = perm[i];
    if (taken < B) {
      if (stopped[r] > 0.f) continue;                                         // finished records are dropped here
      batch_src[taken++] = r;
    } else {
      rest_src[nr++] = r;
    }
  }
  int used = 0;
  if (taken < B) {
    // the pool ran dry: fill_pool() rebuilds it from fresh records (replay_memory.py:237-238, 64-75) and the walk
    // goes on there; fresh records are i.i.d. and never terminated, so "shuffle and pop" is "take the next ones"
    const int fresh0 = P + B;
    while (taken < B) batch_src[taken++] = fresh0 + used++;
    nr = 0;
    while (used < P) rest_src[nr++] = fresh0 + used++;
  }
  *n_rest = nr;
  *fresh_used = used;
}

// replace_memory + fill_pool.  new_step[B]: the STEP column of the generator's output states.  new_pool_src[P] = flat
// indices of the records that form the pool from now on.
EXPO_RL_HD void replace(const float* new_step, int P, int B, int max_traj_len, float keep_prob, Philox& g,
                        const int* rest_src, int n_rest, int fresh_used, long long* new_pool_src) {
  int n = 0;
  for (int i = 0; i < n_rest && n < P; ++i) new_pool_src[n++] = rest_src[i];
  for (int j = 0; j < B; ++j) {
    const float step = new_step[j];
    // replay_memory.py:190-192: always draws the random number only when the step test fails (short-circuit `or`)
    const bool keep = step < (float)max_traj_len || philox_uniform(g) < keep_prob;
    if (keep && n < P) new_pool_src[n++] = P + j;           // beyond P the reference truncates (image_pool[:target])
  }
  const int fresh0 = P + B;
  while (n < P) new_pool_src[n++] = fresh0 + fresh_used++;  // fill_pool: top up with fresh RAW records
}

// replay_fake_batch.  perm: a random order of the pool.  Returns the number of terminated records (0 = the reference's
// assertion would fire).
EXPO_RL_HD int draw_critic(const float* stopped, int P, int B, const int* perm, int* term, long long* batch_src) {
  int nt = 0;
  for (int i = 0; i < P; ++i)
    if (stopped[perm[i]] > 0.f) term[nt++] = perm[i];
  for (int i = 0; i < B; ++i) batch_src[i] = nt > 0 ? term[i % nt] : perm[i % P];
  return nt;
}

}  // namespace rl
}  // namespace expo
