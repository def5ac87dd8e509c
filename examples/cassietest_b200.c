/* The reference's example/cassietest.c (its lines 23-45, minus the window) against this library's header: the same verbs, one environment,
 * then the batched entry point.  Build:  gcc -I include examples/cassietest_b200.c -L cassie-mujoco-sim_b200 -lcassie_b200 -Wl,-rpath,'$ORIGIN/../cassie-mujoco-sim_b200'
 * Needs a CUDA device at run time (there is no CPU backend: without one it reports the error and exits 1). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "cassie_b200.h"

int main(int argc, char **argv) {
  const char *model = argc > 1 ? argv[1] : "cassie-mujoco-sim_b200/models/cassie.cmodel";
  cassie_sim_t *c = cassie_sim_init(model, false);
  if (!c) { fprintf(stderr, "cassie_sim_init failed: %s\n", cassie_b200_last_error()); return 1; }
  pd_in_t u; memset(&u, 0, sizeof u);
  state_out_t y;
  for (int i = 0; i < 1000; i++) cassie_sim_step_pd(c, &y, &u);          /* BASELINE config 1: zero pd_in_t, 1000 ticks */
  double cfrc[12]; cassie_sim_foot_forces(c, cfrc);
  printf("t = %.3f s, pelvis z = %.4f m, left foot force z = %.1f N\n", *cassie_sim_time(c), cassie_sim_qpos(c)[2], cfrc[2]);
  cassie_sim_free(c);

  const int n = 4096;
  cassie_batch_t *b = cassie_batch_init(model, n, 0, CASSIE_B200_FP32);
  if (!b) { fprintf(stderr, "cassie_batch_init failed: %s\n", cassie_b200_last_error()); return 1; }
  pd_in_t *pu = calloc(n, sizeof *pu); state_out_t *py = calloc(n, sizeof *py);
  for (int i = 0; i < 100; i++) cassie_sim_step_pd_batch(b, pu, py);     /* == n x cassie_sim_step_pd per call */
  printf("batch of %d: env 17 motor 3 position %.4f rad, left foot at z = %.4f m in the pelvis frame\n", n, py[17].motor.position[3], py[17].leftFoot.position[2]);
  free(pu); free(py); cassie_batch_free(b);
  return 0;
}
