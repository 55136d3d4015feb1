/* Plain-C consumer of include/bnn_b200.h: proves the boundary is a C ABI (no C++ / torch types in the signatures) and
 * exercises the host-only entry points (no GPU needed).  Built and run by tests/test_abi_cpu.py::test_plain_c_consumer. */
#include <stdio.h>
#include <string.h>

#include "bnn_b200.h"

int main(void) {
  BnnMlpDesc d;
  int32_t masked[BNN_MAX_LINEAR];
  int i;
  memset(&d, 0, sizeof d);
  d.n_linear = 3;
  d.dims[0] = 16; d.dims[1] = 200; d.dims[2] = 200; d.dims[3] = 1;
  d.act = BNN_ACT_RELU;
  if (bnn_abi_version() != BNN_ABI_VERSION) { printf("abi %d\n", (int)bnn_abi_version()); return 1; }
  /* cfg 4: rows of roundup4(in + 1) floats: 200 x 20 + 200 x 204 + 1 x 204 = 45004 */
  if (bnn_param_stride(&d) != 45004) { printf("stride %lld\n", (long long)bnn_param_stride(&d)); return 2; }
  if (bnn_layer_ld(&d, 0) != 20 || bnn_layer_ld(&d, 1) != 204 || bnn_layer_ld(&d, 2) != 204) return 3;
  if (bnn_layer_offset(&d, 0) != 0 || bnn_layer_offset(&d, 1) != 4000 || bnn_layer_offset(&d, 2) != 44800) return 4;
  for (i = 0; i < BNN_MAX_LINEAR; ++i) masked[i] = i < 3;
  if (bnn_mask_len(&d, masked) != 16 + 200 + 200) return 5;
  /* error convention: negative status + message */
  d.n_linear = 0;
  if (bnn_param_stride(&d) >= 0) return 6;
  if (strlen(bnn_last_error()) == 0) return 7;
  if (bnn_forward(NULL, NULL) >= 0) return 8;
  printf("ok\n");
  return 0;
}
