/* Compiled as C (not C++) and linked against libamb200.so by tests/test_abi.py: the header is valid
 * C99, every call below resolves, and the host-only entry points (size queries, options, error text,
 * argument checks) behave without a GPU.  No compute call. */
#include <stdio.h>
#include <string.h>

#include "amb200.h"

int main(void) {
  int failures = 0;
  /* size queries are pure host arithmetic */
  if (amb_packed_bytes(1000, 512) == 0) { printf("amb_packed_bytes\n"); ++failures; }
  if (amb_knn_ws_bytes(1000, 1000, 512, 5) == 0) { printf("amb_knn_ws_bytes\n"); ++failures; }
  if (amb_prdc_ws_bytes(1000, 900) == 0) { printf("amb_prdc_ws_bytes\n"); ++failures; }
  if (amb_frechet_ws_bytes(1, 512) == 0) { printf("amb_frechet_ws_bytes\n"); ++failures; }
  if (amb_prdc_list_cap(1000, 900) < 1) { printf("amb_prdc_list_cap\n"); ++failures; }
  /* options: unknown names are argument errors with a message */
  if (amb_set_option("no_such_option", 1) != AMB_ERR_ARG) { printf("amb_set_option\n"); ++failures; }
  if (strlen(amb_last_error()) == 0) { printf("amb_last_error\n"); ++failures; }
  if (amb_get_option("fad_method") != 0) { printf("amb_get_option\n"); ++failures; }
  /* argument checks come before any CUDA call */
  {
    float r[4];
    if (amb_host_knn_radii(0, NULL, AMB_F32, 4, 8, 1, r) != AMB_ERR_ARG) { printf("amb_host_knn_radii\n"); ++failures; }
  }
  {
    amb_comm_t* comm = NULL;
    int devs[2] = {0, 0};
    if (amb_comm_init(devs, 2, &comm) == AMB_OK) { printf("amb_comm_init accepted a duplicate device\n"); ++failures; }
    if (amb_comm_size(NULL) != 0 || amb_comm_destroy(NULL) != AMB_OK) { printf("amb_comm_size / destroy\n"); ++failures; }
  }
  printf(failures ? "FAILED %d\n" : "ok\n", failures);
  return failures;
}
