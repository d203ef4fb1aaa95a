#define RB_HOST_SIM 1
#include "../../rabe_b200/csrc/pairing.cuh"
#include <cstdio>
using namespace rb;
#define DT_FN static
#include "devtest_body.h"
int main() {
  static uint8_t out[384 * N_SLOTS];
  run_tests(out);
  for (int s = 0; s < N_SLOTS; ++s) { printf("%2d ", s); for (int i = 0; i < 16; ++i) printf("%02x", out[384 * s + i]); printf("..."); for (int i = 368; i < 384; ++i) printf("%02x", out[384 * s + i]); printf("\n"); }
}
