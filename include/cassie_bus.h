/* cassie_bus.h -- the five bus structs that cross the cassie_sim_* C-ABI.
 *
 * These are ABI: field order, types and padding must equal the reference's
 *   pd_in_t           /root/reference/include/pd_in_t.h:24-49        (952 B)
 *   state_out_t       /root/reference/include/state_out_t.h:24-78    (992 B)
 *   cassie_out_t      /root/reference/include/cassie_out_t.h:27-109  (1336 B)
 *   cassie_in_t       /root/reference/include/cassie_in_t.h:24-52    (192 B)
 *   cassie_user_in_t  /root/reference/include/cassie_user_in_t.h:24-27 (104 B)
 * so that a caller compiled against the reference headers (or the generated ctypes mirror,
 * example/cassiemujoco_ctypes.py) can pass its buffers unchanged.  Sizes are checked at
 * compile time below and again in tests/test_abi.py against the reference's own ctypes mirror.
 */
#ifndef CASSIE_BUS_H
#define CASSIE_BUS_H
#include <stdbool.h>

/* guards: if the caller already included the reference's own headers, reuse those typedefs */
#ifndef PD_IN_T_H
#define PD_IN_T_H
typedef struct { double torque[5], pTarget[5], dTarget[5], pGain[5], dGain[5]; } pd_motor_in_t;
typedef struct { double torque[6], pTarget[6], dTarget[6], pGain[6], dGain[6]; } pd_task_in_t;
typedef struct { pd_task_in_t taskPd; pd_motor_in_t motorPd; } pd_leg_in_t;
typedef struct { pd_leg_in_t leftLeg, rightLeg; double telemetry[9]; } pd_in_t;
#endif

#ifndef STATE_OUT_T_H
#define STATE_OUT_T_H
typedef struct { double stateOfCharge, current; } state_battery_out_t;
typedef struct {
  double position[3], orientation[4], footRotationalVelocity[3], footTranslationalVelocity[3];
  double toeForce[3], heelForce[3];
} state_foot_out_t;
typedef struct { double position[6], velocity[6]; } state_joint_out_t;
typedef struct { double position[10], velocity[10], torque[10]; } state_motor_out_t;
typedef struct {
  double position[3], orientation[4], rotationalVelocity[3], translationalVelocity[3];
  double translationalAcceleration[3], externalMoment[3], externalForce[3];
} state_pelvis_out_t;
typedef struct { double channel[16]; bool signalGood; } state_radio_out_t;
typedef struct { double height, slope[2]; } state_terrain_out_t;
typedef struct {
  state_pelvis_out_t pelvis;
  state_foot_out_t leftFoot, rightFoot;
  state_terrain_out_t terrain;
  state_motor_out_t motor;
  state_joint_out_t joint;
  state_radio_out_t radio;
  state_battery_out_t battery;
} state_out_t;
#endif

#ifndef CASSIE_OUT_T_H
#define CASSIE_OUT_T_H
typedef short DiagnosticCodes;
typedef struct { bool dataGood; double stateOfCharge, voltage[12], current, temperature[4]; } battery_out_t;
typedef struct { double position, velocity; } cassie_joint_out_t;
typedef struct {
  unsigned short statusWord;
  double position, velocity, torque, driveTemperature, dcLinkVoltage, torqueLimit, gearRatio;
} elmo_out_t;
typedef struct {
  elmo_out_t hipRollDrive, hipYawDrive, hipPitchDrive, kneeDrive, footDrive;
  cassie_joint_out_t shinJoint, tarsusJoint, footJoint;
  unsigned char medullaCounter;
  unsigned short medullaCpuLoad;
  bool reedSwitchState;
} cassie_leg_out_t;
typedef struct { bool radioReceiverSignalGood, receiverMedullaSignalGood; double channel[16]; } radio_out_t;
typedef struct {
  int etherCatStatus[6], etherCatNotifications[21];
  double taskExecutionTime;
  unsigned int overloadCounter;
  double cpuTemperature;
} target_pc_out_t;
typedef struct {
  bool dataGood;
  unsigned short vpeStatus;
  double pressure, temperature, magneticField[3], angularVelocity[3], linearAcceleration[3], orientation[4];
} vectornav_out_t;
typedef struct {
  target_pc_out_t targetPc;
  battery_out_t battery;
  radio_out_t radio;
  vectornav_out_t vectorNav;
  unsigned char medullaCounter;
  unsigned short medullaCpuLoad;
  bool bleederState, leftReedSwitchState, rightReedSwitchState;
  double vtmTemperature;
} cassie_pelvis_out_t;
typedef struct {
  cassie_pelvis_out_t pelvis;
  cassie_leg_out_t leftLeg, rightLeg;
  bool isCalibrated;
  DiagnosticCodes messages[4];
} cassie_out_t;
#endif

#ifndef CASSIE_IN_T_H
#define CASSIE_IN_T_H
typedef struct { unsigned short controlWord; double torque; } elmo_in_t;
typedef struct { elmo_in_t hipRollDrive, hipYawDrive, hipPitchDrive, kneeDrive, footDrive; } cassie_leg_in_t;
typedef struct { short channel[14]; } radio_in_t;
typedef struct { radio_in_t radio; bool sto, piezoState; unsigned char piezoTone; } cassie_pelvis_in_t;
typedef struct { cassie_pelvis_in_t pelvis; cassie_leg_in_t leftLeg, rightLeg; } cassie_in_t;
#endif

#ifndef CASSIE_USER_IN_T_H
#define CASSIE_USER_IN_T_H
typedef struct { double torque[10]; short telemetry[9]; } cassie_user_in_t;
#endif

#if defined(__cplusplus)
static_assert(sizeof(pd_in_t) == 952 && sizeof(state_out_t) == 992 && sizeof(cassie_out_t) == 1336 &&
              sizeof(cassie_in_t) == 192 && sizeof(cassie_user_in_t) == 104, "bus struct ABI mismatch");
#else
_Static_assert(sizeof(pd_in_t) == 952 && sizeof(state_out_t) == 992 && sizeof(cassie_out_t) == 1336 &&
               sizeof(cassie_in_t) == 192 && sizeof(cassie_user_in_t) == 104, "bus struct ABI mismatch");
#endif
#endif /* CASSIE_BUS_H */
