/* Differential fuzz of the oracle's Agility-block twins (o_pd_input_step, o_core_sim_step in cassie_oracle.c)
 * against the reference's closed archive src/libagilitycassie.a.  TEST INFRASTRUCTURE.  Built by
 * `make -C oracle fuzz` into oracle/_ref/fuzz_agility (needs /root/reference).  Usage: fuzz_agility [n] [mode] */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "../include/cassie_bus.h"
typedef struct CassieCoreSim cassie_core_sim_t; typedef struct PdInput pd_input_t;
cassie_core_sim_t *cassie_core_sim_alloc(void); void cassie_core_sim_setup(cassie_core_sim_t *);
void cassie_core_sim_step(cassie_core_sim_t *, const cassie_user_in_t *, const cassie_out_t *, cassie_in_t *);
pd_input_t *pd_input_alloc(void); void pd_input_setup(pd_input_t *);
void pd_input_step(pd_input_t *, const pd_in_t *, const cassie_out_t *, cassie_user_in_t *);
void o_pd_input_step(const pd_in_t *u, const cassie_out_t *o, double torque[10]);
void o_core_sim_step(const double u[10], const cassie_out_t *o, double out[10]);
static double urand(double a, double b) { return a + (b - a) * (rand() / (double)RAND_MAX); }
static elmo_out_t *drv(cassie_out_t *o, int i) { cassie_leg_out_t *l = i < 5 ? &o->leftLeg : &o->rightLeg; elmo_out_t *t[5] = {&l->hipRollDrive, &l->hipYawDrive, &l->hipPitchDrive, &l->kneeDrive, &l->footDrive}; return t[i % 5]; }
static const elmo_in_t *din(const cassie_in_t *o, int i) { const cassie_leg_in_t *l = i < 5 ? &o->leftLeg : &o->rightLeg; const elmo_in_t *t[5] = {&l->hipRollDrive, &l->hipYawDrive, &l->hipPitchDrive, &l->kneeDrive, &l->footDrive}; return t[i % 5]; }
int main(int argc, char **argv) {
  int n = argc > 1 ? atoi(argv[1]) : 200000, mode = argc > 2 ? atoi(argv[2]) : 0;
  static const double lo[5] = {-0.3, -0.4, -0.9, -2.9, -2.5}, hi[5] = {0.3, 0.4, 1.4, -0.6, -0.5};
  static const double tl[5] = {140.63, 140.63, 216.16, 216.16, 45.14};
  cassie_core_sim_t *core = cassie_core_sim_alloc(); cassie_core_sim_setup(core);
  pd_input_t *pd = pd_input_alloc(); pd_input_setup(pd);
  double maxe_pd = 0, maxe_core = 0; int worst = -1; srand(12345);
  for (int it = 0; it < n; it++) {
    cassie_out_t o; memset(&o, 0, sizeof o); o.isCalibrated = 1; o.pelvis.radio.channel[8] = (it % 17 == 0) ? urand(-1, 1) : 1;
    o.pelvis.radio.radioReceiverSignalGood = 1; o.pelvis.radio.receiverMedullaSignalGood = 1;
    for (int i = 0; i < 10; i++) {
      elmo_out_t *e = drv(&o, i); int k = i % 5; double mid = 0.5 * (lo[k] + hi[k]), half = 0.5 * (hi[k] - lo[k]);
      /* mode 0: mostly inside limits, a few joints outside; mode 1: everything anywhere */
      double w = (mode == 1 || rand() % 4 == 0) ? 1.0 : 0.55;
      e->position = mid + half * w * urand(-1, 1); if (i >= 5 && k == 0) e->position = -e->position;
      e->velocity = urand(-8, 8); e->torqueLimit = tl[k]; e->gearRatio = 1; e->statusWord = 0x0637;
    }
    { cassie_joint_out_t *jj[6] = {&o.leftLeg.shinJoint, &o.leftLeg.tarsusJoint, &o.leftLeg.footJoint, &o.rightLeg.shinJoint, &o.rightLeg.tarsusJoint, &o.rightLeg.footJoint};
      for (int i = 0; i < 6; i++) { jj[i]->position = (i % 3 == 0) ? urand(-0.2, 0.2) : (i % 3 == 1 ? urand(0.8, 2.6) : urand(-2.4, -0.6)); jj[i]->velocity = urand(-6, 6); } }
    pd_in_t u; memset(&u, 0, sizeof u);
    if (it % 3) for (int sd = 0; sd < 2; sd++) { pd_task_in_t *t = sd ? &u.rightLeg.taskPd : &u.leftLeg.taskPd;   /* taskPd branch: two thirds of the cases */
      for (int k = 0; k < 6; k++) { t->torque[k] = urand(-30, 30); t->pTarget[k] = urand(-1, 1); t->dTarget[k] = urand(-2, 2); t->pGain[k] = (rand() % 4) ? urand(0, 300) : 0; t->dGain[k] = (rand() % 4) ? urand(0, 10) : 0; } }
    for (int i = 0; i < 10; i++) { pd_motor_in_t *p = i < 5 ? &u.leftLeg.motorPd : &u.rightLeg.motorPd; int k = i % 5;
      p->torque[k] = urand(-50, 50); p->pTarget[k] = urand(-2, 2); p->dTarget[k] = urand(-3, 3); p->pGain[k] = urand(0, 200); p->dGain[k] = urand(0, 10); }
    cassie_user_in_t ui; pd_input_step(pd, &u, &o, &ui);
    double tq[10]; o_pd_input_step(&u, &o, tq);
    for (int i = 0; i < 10; i++) { double e = fabs(tq[i] - ui.torque[i]); if (e > maxe_pd) maxe_pd = e; }
    cassie_in_t ci; cassie_core_sim_step(core, &ui, &o, &ci);
    double out[10]; o_core_sim_step(ui.torque, &o, out);
    for (int i = 0; i < 10; i++) { double e = fabs(out[i] - din(&ci, i)->torque); if (e > maxe_core) { maxe_core = e; worst = it;
      if (e > 1e-3 && argc > 3) { printf("it %d motor %d ref %.6f twin %.6f u %.4f sto %.3f\n pos:", it, i, din(&ci, i)->torque, out[i], ui.torque[i], o.pelvis.radio.channel[8]);
        for (int j = 0; j < 10; j++) printf(" %.4f", drv(&o, j)->position); printf("\n vel:"); for (int j = 0; j < 10; j++) printf(" %.3f", drv(&o, j)->velocity); printf("\n"); } } }
  }
  printf("n=%d mode=%d max|pd twin - archive| = %.3e   max|core twin - archive| = %.3e (worst it %d)\n", n, mode, maxe_pd, maxe_core, worst);
  return 0;
}
