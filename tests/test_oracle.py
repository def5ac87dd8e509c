"""The oracle itself: physics invariants (no MuJoCo is available to pin it -- SURVEY.md section 8c) and committed golden trajectories."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, PD_DGAIN, PD_PGAIN, PD_TARGET, REPO

OMODEL = os.path.join(GOLDEN, 'cassie.omodel')


def table(path):
    d = {}
    for line in open(path):
        if line[0] == '#':
            continue
        t = line.split()
        if t[1] != 'S':
            d[t[0]] = np.array([float(x) for x in t[3:3 + int(t[2])]])
    return d


def dense_M(o, T):
    nv = 32
    qM, par, adr = o.arr('qM'), T['dof_parentid'].astype(int), T['dof_Madr'].astype(int)
    D = np.zeros((nv, nv))
    for i in range(nv):
        a, j = adr[i], i
        while j >= 0:
            D[i, j] = D[j, i] = qM[a]
            a += 1
            j = par[j]
    return D


def write_table(T, path, **over):
    with open(path, 'w') as f:
        for line in open(OMODEL):
            k = line.split()[0]
            if k in over:
                v = np.asarray(over[k])
                ty = 'I' if v.dtype.kind in 'iu' else 'F'
                f.write('%s %s %d %s\n' % (k, ty, v.size, ' '.join(('%d' if ty == 'I' else '%.17g') % x for x in v.ravel())))
            else:
                f.write(line)


def test_mass_matrix_is_spd_and_total_mass(oracle_mod):
    o = oracle_mod.OracleSim(OMODEL)
    T = table(OMODEL)
    M = dense_M(o, T)
    assert np.linalg.eigvalsh(M).min() > 0
    assert abs(M[0, 0] - T['body_mass'].sum()) < 1e-9 and abs(M[0, 0] - 33.312) < 2e-3
    # sparse L'DL solve == dense solve
    b = o.arr('qfrc_smooth').copy()
    assert np.abs(np.linalg.solve(M, b) - o.arr('qacc_smooth')).max() < 1e-8


def test_gravity_bias_is_potential_gradient(oracle_mod):
    o = oracle_mod.OracleSim(OMODEL)
    T = table(OMODEL)
    q0 = o.arr('qpos').copy()
    o.arr('qvel')[:] = 0
    o.forward()
    bias = o.arr('qfrc_bias').copy()

    def V(q):
        o.arr('qpos')[:] = q
        o.forward()
        return 9.81 * np.dot(T['body_mass'], o.arr('xipos').reshape(-1, 3)[:, 2])
    for j in range(26):
        if int(T['jnt_type'][j]) in (2, 3):
            qa, da, e = int(T['jnt_qposadr'][j]), int(T['jnt_dofadr'][j]), 1e-6
            qp, qm = q0.copy(), q0.copy()
            qp[qa] += e
            qm[qa] -= e
            assert abs((V(qp) - V(qm)) / (2 * e) - bias[da]) < 1e-5


def test_energy_drift_is_first_order_in_h(oracle_mod, tmp_path):
    """unconstrained, undamped tumbling tree: total energy error over a fixed horizon halves when h halves"""
    T = table(OMODEL)

    def drift(h):
        p = str(tmp_path / ('free_%g.omodel' % h))
        write_table(T, p, neq=np.array([0]), dof_damping=np.zeros(32), jnt_limited=np.zeros(26, dtype=int), geom_contype=np.zeros(25, dtype=int),
                    geom_conaffinity=np.zeros(25, dtype=int), opt_timestep=np.array([h]), eq_obj1id=np.zeros(0, dtype=int), eq_obj2id=np.zeros(0, dtype=int))
        o = oracle_mod.OracleSim(p)
        o.arr('qvel')[:] = np.random.default_rng(0).normal(size=32)

        def E():
            o.forward()
            v = o.arr('qvel')
            ke = 0.5 * v @ dense_M(o, T) @ v
            pe = 9.81 * np.dot(T['body_mass'], o.arr('xipos').reshape(-1, 3)[:, 2])
            q = o.arr('qpos')
            for j in range(26):
                if T['jnt_stiffness'][j] > 0:
                    pe += 0.5 * T['jnt_stiffness'][j] * q[int(T['jnt_qposadr'][j])] ** 2
            return ke + pe
        e0 = E()
        for _ in range(int(round(0.2 / h))):
            o.mj_step()
        return E() - e0
    d1, d2 = drift(5e-4), drift(2.5e-4)
    assert abs(d1) < 0.5 and 1.7 < d1 / d2 < 2.3, (d1, d2)


def test_pd_stance_contact_forces_carry_the_weight(oracle_mod):
    o = oracle_mod.OracleSim(OMODEL)
    u = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    forces = []
    for k in range(1000):
        o.step_pd(u)
        if k >= 300 and o.get_int('ncon') > 0:
            n, ne = o.get_int('nefc'), 12 + o.get_int('nl')
            forces.append(o.arr('efc_force')[ne:n].sum())
    assert abs(np.mean(forces) - 33.312 * 9.81) < 0.15 * 33.312 * 9.81     # normal force ~ weight while (quasi) standing
    assert np.abs(o.arr("efc_pos")[:12]).max() < 5e-3                       # loop closures hold
    q = o.arr('qpos')
    for a in (3, 10, 24):
        assert abs(np.linalg.norm(q[a:a + 4]) - 1) < 1e-12
    assert o.get_int('unsupported_pairs') == 0 and o.get_int('dropped_contacts') == 0


def test_left_right_symmetry(oracle_mod):
    """mirror-symmetric model + symmetric input: left and right motor joints move identically (up to the slightly asymmetric qpos_init)"""
    o = oracle_mod.OracleSim(OMODEL)
    q = o.arr('qpos')
    q[21:35] = q[7:21]
    q[21] = -q[7]; q[22] = -q[8]                      # roll / yaw mirror
    q[24] = q[10]; q[25] = -q[11]; q[26] = -q[12]; q[27] = q[13]     # achilles quaternion reflected in the leg's local xy plane: (w, -x, -y, z)
    o.forward()
    u = oracle_mod.make_pd()
    for _ in range(200):
        o.step_pd(u)
    q = o.arr('qpos')
    assert abs(q[9] - q[23]) < 1e-4 and abs(q[14] - q[28]) < 1e-4 and abs(q[20] - q[34]) < 1e-4   # hip pitch, knee, foot
    assert abs(q[1]) < 1e-4                                                                        # no lateral drift


@pytest.mark.parametrize('name', ['traj_zero_pd', 'traj_fixed_pd'])
def test_golden_trajectory(oracle_mod, name):
    """regression pin: committed fp64 oracle trajectories (tests/golden/make_golden.py)"""
    g = np.load(os.path.join(GOLDEN, name + '.npz'))
    u = oracle_mod.make_pd() if name == 'traj_zero_pd' else oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    o = oracle_mod.OracleSim(OMODEL)
    ticks = list(g['ticks'])
    for k in range(1, int(max(ticks)) + 1):
        o.step_pd(u)
        if k in ticks:
            i = ticks.index(k)
            assert np.abs(o.arr('qpos') - g['qpos'][i]).max() < 1e-7, k
            assert np.abs(o.arr('qvel') - g['qvel'][i]).max() < 1e-5, k


def test_capacity_option_of_the_checker(oracle_mod):
    """the oracle can be given the product's capacity limits (12 contacts / 48 rows in the product; tiny ones here) so that overflow situations stay
    comparable: contacts beyond the cap are dropped in contact order and counted, the row count never exceeds the cap"""
    o = oracle_mod.OracleSim(os.path.join(GOLDEN, 'cassie.omodel'))
    o.set_caps(2, 20)
    u = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    most = 0
    for _ in range(600):
        o.step_pd(u)
        assert o.get_int('ncon') <= 2 and o.get_int('nefc') <= 20
        most = max(most, o.get_int('ncon'))
    assert most == 2 and o.get_int('dropped_contacts') > 0 and np.isfinite(o.arr('qpos')).all()


def test_free_flight_conserves_momentum(oracle_mod):
    """no gravity, no contact (model/cassie_no_grav.xml lifted off the floor): springs, joint limits, loop closures, rotor inertias, damping and
    the PD motors are all internal forces, so the centre-of-mass velocity and the angular momentum about it stay put while the joints thrash.
    A physics invariant that needs no reference trajectory; the queries mix one-sub-step-old kinematics with new velocities, hence O(h) noise."""
    import emu_harness as E
    rng = np.random.default_rng(1)
    for pd in (False, True):
        o = oracle_mod.OracleSim(os.path.join(GOLDEN, 'cassie_no_grav.omodel'))
        e = E.EmuSim(os.path.join(REPO, 'cassie-mujoco-sim_b200', 'models', 'cassie_no_grav.cmodel'))
        q, v = o.arr('qpos'), o.arr('qvel')
        q[2] = 3.0
        v[:] = rng.normal(0, 0.7, 32)
        e.set('qpos', np.concatenate([q, np.zeros(len(e.get('qpos')) - 35)]))
        e.set('qvel', v)
        o.forward()
        e.forward()
        u = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN) if pd else oracle_mod.make_pd()
        row = np.concatenate([np.zeros(10), PD_TARGET, np.zeros(10), PD_PGAIN, PD_DGAIN]) if pd else np.zeros(50)
        o.step_pd(u)
        e.step(row)
        cm0, L0, dev, dev_e = o.cm_velocity().copy(), o.angular_momentum().copy(), np.zeros(2), np.zeros(2)
        j0, jpeak = o.arr('qpos')[7:35].copy(), 0.0
        for k in range(400):
            o.step_pd(u)
            e.step(row)
            dev = np.maximum(dev, [np.abs(o.cm_velocity() - cm0).max(), np.abs(o.angular_momentum() - L0).max()])
            jpeak = max(jpeak, np.abs(o.arr('qpos')[7:35] - j0).max())
            if k % 50 == 49:    # the kernel source's derived-quantity row says the same
                e.query()
                a = e.get('aux')
                dev_e = np.maximum(dev_e, [np.abs(a[45:48] - cm0).max(), np.abs(a[48:51] - L0).max()])
        assert len(o.contacts()) == 0 and jpeak > 0.05      # free flight, and the joints did move
        assert np.abs(L0).max() > 0.5 and dev[0] < 2e-4 and dev[1] < 5e-3, (pd, cm0, L0, dev)
        assert dev_e[0] < 2e-4 and dev_e[1] < 5e-3, (pd, dev_e)
        e.close()


def test_half_and_quarter_turn_invariance(oracle_mod):
    """the same start turned about the vertical by 180 or 90 degrees gives the same trajectory turned back (flat floor; the friction pyramid has
    exactly that symmetry, an arbitrary heading does not): kinematics, Jacobians, collision frames, constraint rows and sensors agree on what
    a rotation is.  Sliding touch-down under PD control, oracle and kernel source."""
    import emu_harness as E

    def qmul(a, b):
        return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                         a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])
    u = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    row = np.concatenate([np.zeros(10), PD_TARGET, np.zeros(10), PD_PGAIN, PD_DGAIN])

    def run(yaw, n, emu):
        c, s = np.cos(yaw), np.sin(yaw)
        Rz, qz = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]]), np.array([np.cos(yaw / 2), 0, 0, np.sin(yaw / 2)])
        o = oracle_mod.OracleSim(OMODEL)
        q, v = o.arr('qpos').copy(), o.arr('qvel').copy()
        q[0:3], q[3:7], v[0:3] = Rz @ [0.3, -0.2, q[2]], qmul(qz, q[3:7]), Rz @ [0.4, 0.1, 0.0]     # free joint: linear velocity in the world frame
        if emu:
            e = E.EmuSim(os.path.join(REPO, 'cassie-mujoco-sim_b200', 'models', 'cassie.cmodel'))
            e.set('qpos', q)
            e.set('qvel', v)
            e.forward()
            for k in range(n):
                e.step(row)
            q, v, ncon = e.get('qpos')[:35].copy(), e.get('qvel')[:32].copy(), int(e.get('counters')[1])
            e.close()
        else:
            o.arr('qpos')[:], o.arr('qvel')[:] = q, v
            o.forward()
            for k in range(n):
                o.step_pd(u)
            q, v, ncon = o.arr('qpos').copy(), o.arr('qvel').copy(), len(o.contacts())
        q[0:3], q[3:7], v[0:3] = Rz.T @ q[0:3], qmul(qz * [1, 1, 1, -1], q[3:7]), Rz.T @ v[0:3]
        return q, v, ncon
    for emu in (False, True):
        q0, v0, ncon = run(0.0, 400, emu)
        assert ncon >= 2 and abs(q0[0] - 0.3) > 0.01                      # it landed and slid
        for yaw, tol in ((np.pi, 1e-11), (np.pi / 2, 1e-6)):               # the quarter turn permutes the pyramid's rows: Gauss-Seidel order noise
            q1, v1, _ = run(yaw, 400, emu)
            assert np.abs(q1 - q0).max() < tol and np.abs(v1 - v0).max() < 100 * tol, (emu, yaw, np.abs(q1 - q0).max(), np.abs(v1 - v0).max())


def test_free_fall_is_ballistic(oracle_mod):
    """gravity on, no contact: the centre of mass falls with g (MuJoCo's default 9.81, model/cassie.xml sets none) whatever the joints do, its
    horizontal velocity and the angular momentum about it stay put"""
    rng = np.random.default_rng(2)
    o = oracle_mod.OracleSim(OMODEL)
    q, v = o.arr('qpos'), o.arr('qvel')
    q[2] = 4.0
    v[:] = rng.normal(0, 0.7, 32)
    o.forward()
    u = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    o.step_pd(u)
    cm0, L0, t0 = o.cm_velocity().copy(), o.angular_momentum().copy(), float(o.arr('time')[0])
    worst = np.zeros(3)
    for k in range(300):
        o.step_pd(u)
        t = float(o.arr('time')[0]) - t0
        cm = o.cm_velocity()
        worst = np.maximum(worst, [np.abs(cm[:2] - cm0[:2]).max(), abs(cm[2] - (cm0[2] - 9.81 * t)), np.abs(o.angular_momentum() - L0).max()])
    assert len(o.contacts()) == 0 and t == pytest.approx(0.15)
    assert worst[0] < 2e-4 and worst[1] < 2e-4 and worst[2] < 5e-3, worst


def test_constraint_forces_satisfy_the_kkt_conditions(oracle_mod):
    """the constraint solve minimises 1/2 f'(A + R) f + f'b with f free on the equality rows and f >= 0 on limit and pyramid rows (MuJoCo's dual
    problem).  At the solver's exit the gradient g = (A + R) f + b vanishes on equality rows and active rows and is non-negative on rows at
    zero force -- to the solver's own stopping tolerance, relative to |b|.  Independent of how the solution was reached."""
    o = oracle_mod.OracleSim(OMODEL)
    u = oracle_mod.make_pd(pTarget=PD_TARGET, pGain=PD_PGAIN, dGain=PD_DGAIN)
    seen_contact = seen_inactive = 0
    for k in range(450):
        o.step_pd(u)
        if k % 25 != 24:
            continue
        n, ne = o.get_int('nefc'), o.get_int('ne')
        AR, b, f = o.arr('efc_AR')[:n * n].reshape(n, n), o.arr('efc_b')[:n], o.arr('efc_force')[:n]
        g, scale = AR @ f + b, np.abs(b).max()
        assert np.abs(AR - AR.T).max() < 1e-9 * np.abs(AR).max() and np.linalg.eigvalsh(AR).min() > 0      # A + R symmetric positive definite
        assert np.abs(g[:ne]).max() < 2e-4 * scale, (k, np.abs(g[:ne]).max(), scale)
        if n > ne:
            fi, gi = f[ne:], g[ne:]
            assert fi.min() >= 0 and gi.min() > -2e-4 * scale and np.abs(gi[fi > 0]).max(initial=0.0) < 2e-4 * scale, (k, fi.min(), gi.min(), scale)
            seen_contact += 1
            seen_inactive += int((fi == 0).any())
    assert seen_contact >= 10
