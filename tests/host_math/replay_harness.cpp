// TEST INFRASTRUCTURE ONLY -- compiles the product's replay selection logic (exposure_b200/csrc/replay_logic.cuh, plain
// integer code shared by the device kernels) for the host, so that tests/test_replay_logic.py can drive thousands of
// iterations of it on the CPU and compare its behaviour with the reference-faithful host ReplayMemory.
#include <stdint.h>
#include <string.h>

#include "../../exposure_b200/csrc/replay_logic.cuh"

using namespace expo::rl;

extern "C" {
// one generator draw: returns n_rest; fresh_used through the pointer
int rl_draw_generator(const float* pool_states, int S, int P, int B, unsigned long long seed, unsigned long long call,
                      long long* batch_src, int* rest_src, int* fresh_used) {
  static int perm[1024];
  static uint32_t keys[1024];
  static float stopped[1024];
  for (int i = 0; i < P; ++i) stopped[i] = pool_states[(long long)i * S + kStateStopped];
  shuffle(perm, keys, P, seed, call, 1u);
  int n_rest = 0;
  draw_generator(stopped, P, B, perm, batch_src, rest_src, &n_rest, fresh_used);
  return n_rest;
}
void rl_replace(const float* new_states, int S, int P, int B, int max_len, float keep, unsigned long long seed,
                unsigned long long call, const int* rest_src, int n_rest, int fresh_used, long long* new_pool_src) {
  static float step[1024];
  for (int j = 0; j < B; ++j) step[j] = new_states[(long long)j * S + kStateStep];
  Philox g;
  philox_init(g, seed, call, 2u);
  replace(step, P, B, max_len, keep, g, rest_src, n_rest, fresh_used, new_pool_src);
}
int rl_draw_critic(const float* pool_states, int S, int P, int B, unsigned long long seed, unsigned long long call,
                   long long* batch_src) {
  static int perm[1024], term[1024];
  static uint32_t keys[1024];
  static float stopped[1024];
  for (int i = 0; i < P; ++i) stopped[i] = pool_states[(long long)i * S + kStateStopped];
  shuffle(perm, keys, P, seed, call, 3u);
  return draw_critic(stopped, P, B, perm, term, batch_src);
}
void rl_uniforms(unsigned long long seed, unsigned long long call, float* out, int n) {
  Philox g;
  philox_init(g, seed, call, 9u);
  for (int i = 0; i < n; ++i) out[i] = philox_uniform(g);
}
}
