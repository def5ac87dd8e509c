"""Development aid: CPU rehearsal of tests/test_zz_gpu_estimator_filter.py -- the observation row comes from the host emulation of the kernel source
(tests/emu), the leg forces and filters from the product library's host entry points, the reference values from the oracle linked with the real archive.
Not part of the product or of the test suite."""
import sys, os
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in ('tests', 'oracle', ''):
    sys.path.insert(0, os.path.join(REPO, p))
import numpy as np, importlib, ctypes as C
import emu_harness as EH, oracle as O
from conftest import PD_TARGET, PD_PGAIN, PD_DGAIN, GOLDEN
pkg=importlib.import_module('cassie-mujoco-sim_b200')
L=pkg.lib()
dp=C.POINTER(C.c_double)
L.cassie_b200_estimator_leg_force.argtypes=[C.c_int,dp,dp,dp]; L.cassie_b200_estimator_leg_force.restype=None
L.cassie_b200_estimator_filter_new.restype=C.c_void_p; L.cassie_b200_estimator_filter_step.argtypes=[C.c_void_p,C.c_void_p]; L.cassie_b200_estimator_filter_reset.argtypes=[C.c_void_p]
OB_EST_ACC, OB_FOOT, OB_EST_QUAT, OB_QUAT = 56, 60, 86, 42
def unpack(row, y, filt):
    y.pelvis.orientation[:]=list(row[OB_EST_QUAT:OB_EST_QUAT+4]); y.pelvis.translationalAcceleration[:]=list(row[OB_EST_ACC:OB_EST_ACC+3])
    for sd,f in enumerate((y.leftFoot,y.rightFoot)):
        fo=row[OB_FOOT+13*sd:OB_FOOT+13*sd+13]; f.position[:]=list(fo[0:3])
        ang=(C.c_double*7)(row[5*sd],row[5*sd+1],row[5*sd+2],row[5*sd+3],row[30+3*sd],row[31+3*sd],row[5*sd+4]); q=(C.c_double*4)(*row[OB_QUAT:OB_QUAT+4]); out=(C.c_double*3)()
        L.cassie_b200_estimator_leg_force(sd,ang,q,out); f.toeForce[:]=list(out); f.heelForce[:]=list(out)
    L.cassie_b200_estimator_filter_step(filt,C.byref(y))
def filtered(y): return np.concatenate([y.pelvis.position[:],y.pelvis.translationalVelocity[:],y.pelvis.externalForce[:],[y.terrain.height]])
o=O.OracleSim(os.path.join(GOLDEN,'cassie.omodel'),ref=True)
e=EH.EmuSim(pkg.model_path('cassie'))
u=O.make_pd(pTarget=PD_TARGET,pGain=PD_PGAIN,dGain=PD_DGAIN); row=np.concatenate([np.zeros(10),PD_TARGET,np.zeros(10),PD_PGAIN,PD_DGAIN])
y=pkg.state_out_t(); yc=pkg.state_out_t(); filt=L.cassie_b200_estimator_filter_new(); worst=0
for k in range(700):
    o.step_pd(u,y); e.step(row)
    unpack(e.get('obs'),yc,filt)
    a,b=filtered(y),filtered(yc); worst=max(worst,(np.abs(a-b)/(1+np.abs(a))).max())
print('worst rel',worst,'b',b,'toeForce z',yc.leftFoot.toeForce[2])
print('asserts', worst<2e-3, abs(b[2])>0.3, abs(b[8]-31*9.806)>10, b[9]!=0, yc.leftFoot.toeForce[2]<-50)
# full_reset rehearsal: reset filter, next step
L.cassie_b200_estimator_filter_reset(filt)
e.step(row); unpack(e.get('obs'),yc,filt); ag=filtered(yc)
print('after reset', ag, ag[8], np.abs(ag[3:6]).max(), 0.4<ag[2]<1.0)
# ---- test 2 rehearsal: reset one filter at call 150, compare at 300
e=EH.EmuSim(pkg.model_path('cassie')); f1=L.cassie_b200_estimator_filter_new(); f2=L.cassie_b200_estimator_filter_new(); y1=pkg.state_out_t(); y2=pkg.state_out_t()
for k in range(300):
    e.step(row); r=e.get('obs'); unpack(r,y1,f1); unpack(r,y2,f2)
    if k==150: L.cassie_b200_estimator_filter_reset(f2)
w=filtered(y1); g=filtered(y2)
print('reset env differs by', (np.abs(g-w)/(1+np.abs(w))).max(), w, g)
